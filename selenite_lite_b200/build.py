"""Builds the CUDA shared library in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libselenite_b200.so")
# the same sources with -DSL_TC_STRESS: every hand-over point of the tensor-core kernels sleeps a pseudo-random time first
# (csrc/sl_tc_common.cuh). Test infrastructure (tests/test_gpu_stress_build.py), never loaded by default.
STRESS_LIB_PATH = os.path.join(LIB_DIR, "libselenite_b200_stress.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, stress=False):
    """Compile every CUDA source into lib/libselenite_b200.so (and, with stress=True, the stress build beside it, in parallel).
    nvcc cross-compiles without a GPU."""
    if not force and not is_stale() and (not stress or os.path.exists(STRESS_LIB_PATH)):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmds = [[nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()]
    if stress:
        cmds.append([nvcc] + NVCC_FLAGS + ["-DSL_TC_STRESS", "-o", STRESS_LIB_PATH] + sources())
    procs = [subprocess.Popen(c) for c in cmds]
    rcs = [p.wait() for p in procs]
    for c, rc in zip(cmds, rcs):
        if rc != 0:
            raise subprocess.CalledProcessError(rc, c)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
