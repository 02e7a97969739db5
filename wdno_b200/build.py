"""Build libwdno_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libwdno_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


if os.environ.get("WDNO_PROF"):  # debug: per-role stall counters in tapgemm (tools/prof_roles.py)
    NVCC_FLAGS.append("-DWDNO_PROF")


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found; libwdno_b200.so cannot be built")
    return nv


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu and link the shared library. Returns the library path."""
    nv = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hdrs.append(os.path.join(HERE, "..", "include", "wdno_b200.h"))
    jobs = []
    objs = []
    # a change of compile flags (e.g. WDNO_PROF) invalidates every object: the flag set is part of the staleness check
    stamp = os.path.join(OBJ, "flags.txt")
    flags_now = " ".join(NVCC_FLAGS)
    if not os.path.exists(stamp) or open(stamp).read() != flags_now:
        force = True
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nv] + NVCC_FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([nv, "-shared", "-o", LIB] + objs)
    with open(stamp, "w") as fh:
        fh.write(flags_now)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
