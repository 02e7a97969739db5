"""Engine snapshots must never outlive the parameters they were packed from (parent.load_state_dict,
EMA copy_, optimiser steps are in-place writes that bypass the child's load_state_dict override)."""
import torch

from wdno_b200.unet2d import Unet2D
from wdno_b200.unet3d import Unet3D_with_Conv3D


class _Parent(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.model = m


def _check(m):
    made = []
    m._make_engine = lambda: made.append(object()) or made[-1]
    e0 = m.engine()
    assert m.engine() is e0 and len(made) == 1, "unchanged parameters must reuse the snapshot"
    # nested load through the parent (the way the reference loads checkpoints: diffusion.load_state_dict(data['model']))
    parent = _Parent(m)
    parent.load_state_dict({k: v.clone() for k, v in parent.state_dict().items()})
    e1 = m.engine()
    assert e1 is not e0, "parent.load_state_dict must invalidate the engine"
    # in-place update (EMA copy_ / optimiser step)
    with torch.no_grad():
        next(m.parameters()).add_(1.0)
    e2 = m.engine()
    assert e2 is not e1, "in-place parameter updates must invalidate the engine"
    assert m.engine() is e2
    m.invalidate()
    assert m.engine() is not e2


def test_unet3d_engine_follows_parameter_updates():
    _check(Unet3D_with_Conv3D(dim=16, dim_mults=(1, 2), channels=4))


def test_unet2d_engine_follows_parameter_updates():
    _check(Unet2D(dim=16, dim_mults=(1, 2), channels=3))
