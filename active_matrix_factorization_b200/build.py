"""In-tree build of libamf_b200.so (sm_100a only).  `python -m active_matrix_factorization_b200.build`."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libamf_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
# AMF_B200_DEBUG=1: device-side bounds checks on every computed write address (common.cuh);
# build with --force so that every object is recompiled with the flag
if os.environ.get("AMF_B200_DEBUG") == "1":
    FLAGS = FLAGS + ["-DAMF_BOUNDS_CHECK"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(f) > t for f in deps)


def _compile(src):
    obj = src[:-3] + ".o"
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    if os.path.exists(obj) and all(os.path.getmtime(obj) > os.path.getmtime(f) for f in [src] + hdrs):
        return obj
    subprocess.check_call([NVCC] + FLAGS + ["-c", src, "-o", obj])
    return obj


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s; cannot build libamf_b200.so" % NVCC)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(sources())))) as ex:
        objs = list(ex.map(_compile, sources()))
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", LIB] + objs + ["-cudart", "static"])
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
