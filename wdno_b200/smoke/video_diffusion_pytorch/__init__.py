"""mirror of the reference package `smoke/video_diffusion_pytorch`: `video_diffusion_pytorch_conv3d` is the engine's;
`video_diffusion_pytorch` (the unused lucidrains Unet3D that smoke/train_2d.py:7 imports) and `text` fall through to the
reference tree configured by wdno_b200.install()."""
from wdno_b200._dropin import extend_path

__path__ = extend_path(__path__, "smoke/video_diffusion_pytorch")
