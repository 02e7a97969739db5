"""The reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores) runs without a GPU and prints
ONE JSON line with the keys the driver reads; under torchrun only rank 0 works.  The engine arm needs a B200 and is run by
the driver (its keys are checked here statically against the same list)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "cpu_baseline")


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--cpu-budget", "4"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and all(k in d for k in KEYS)
    assert d["unit"] == "steps/s" and d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    # the real reference modules where /root/reference is mounted (this container), the oracle port on the GPU box
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["steps"] == 1 and d["warmup"] == 1   # --steps / --warmup honoured
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"metric": cfg["metric"]') == 2     # both arms report the configuration's one metric string
    for k in KEYS + ("clocks", "gpu_launches", "roofline"):
        assert f'"{k}"' in src, k


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
