"""Builds libcmdlmc_b200.so in-tree with nvcc for sm_100a (no other architecture, no JIT)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcmdlmc_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
              "--fmad=true", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + sources() + ["-o", LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    # --check: a second library with index checks inside the kernels (-DCMD_BOUNDS_CHECK, csrc/common.cuh),
    # for `CMDLMC_B200_LIB=cmdlmc_b200/libcmdlmc_b200_check.so python -m pytest tests -m gpu`
    if "--check" in sys.argv:
        out = os.path.join(HERE, "libcmdlmc_b200_check.so")
        subprocess.run([os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + ["-DCMD_BOUNDS_CHECK"] + sources() +
                       ["-o", out], check=True)
        print(out)
        sys.exit(0)
    build(force="--force" in sys.argv, verbose=True,
          extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
    print(LIB)
