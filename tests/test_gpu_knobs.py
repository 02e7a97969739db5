"""Round-1 verdict, housekeeping: the non-default WDNO_* knobs that switch kernels (fallback / A-B paths) must stay correct.
Each setting runs tests/helpers/knob_probe.py in a fresh process (the knobs are read once per process)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KNOBS = [{}, {"WDNO_ZSTACK": "0"}, {"WDNO_FOLD": "0"}, {"WDNO_FOLD": "force"}, {"WDNO_CONV1X1": "0"}, {"WDNO_TATTN_WARP": "0"},
         {"WDNO_LA2_WARP": "0"}, {"WDNO_LA1_NPH": "2"}, {"WDNO_DWT3D_STREAM": "0"}, {"WDNO_DWT3D_FUSED": "0"}, {"WDNO_NTILE_BIG": "0"},
         {"WDNO_TIME_UNIFORM": "0"}, {"WDNO_PDL": "1"}, {"WDNO_TATTN_TC": "0"}, {"WDNO_TATTN_TC": "1"}, {"WDNO_LINATTN_TC": "0"},
         {"WDNO_LINATTN_TC": "0", "WDNO_LA2_WARP": "0"}, {"WDNO_LINATTN_TC": "0", "WDNO_LA1_NPH": "2"}, {"WDNO_CLUSTER": "1"}]


@pytest.mark.parametrize("knob", KNOBS, ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()) or "defaults")
def test_kernel_switching_knobs_keep_parity(knob):
    env = dict(os.environ, **knob)
    r = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "knob_probe.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("KNOB_JSON ")][-1][len("KNOB_JSON "):])
    assert d["unet3d"] < 4e-3 and d["unet2d"] < 5e-3, (knob, d)
    assert d["wavedec3"] < 1e-5 and d["waverec3"] < 1e-5, (knob, d)
