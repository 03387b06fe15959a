#!/usr/bin/env python
"""Build the reference's own Cython geometry modules into oracle/_ref/  (TEST INFRASTRUCTURE).

This compiles the three real Cython sources of the reference *where they lie* under
/root/reference (nothing is copied into the repository):

    mdlmc/cython_exts/helper/math_helper.pyx
    mdlmc/cython_exts/atoms/numpyatom.pyx
    mdlmc/cython_exts/LMC/PBCHelper.pyx

with the reference's own flags (`-O3 -ffast-math`, C++, setup.py:52-63) minus GSL: the only
GSL symbol used is `round` (numpyatom.pyx:8), which we bind to C99 `round()` through a two
line `cython_gsl` stub that is generated into oracle/_ref/stubs/.

Outputs (all git-ignored, but they travel to the GPU box with the snapshot):

    oracle/_ref/mdlmc/...                      empty package skeleton + the three .so modules
    oracle/_ref/build/*.cpp                    generated C++

The reference's *Python* layers (topology.py, MDMC.py, ...) are never copied; in this container
they are imported straight from /root/reference by oracle/ref_import.py to pin the restatement
in oracle/ and to generate tests/golden/*.npz.  On the GPU box only the compiled AtomBox classes
are available (used as the "reference" CPU baseline of bench.py).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CMDLMC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

MODULES = [
    ("mdlmc/cython_exts/helper", "math_helper"),
    ("mdlmc/cython_exts/atoms", "numpyatom"),
    ("mdlmc/cython_exts/LMC", "PBCHelper"),
]


def have_reference():
    return os.path.isfile(os.path.join(REF, "mdlmc/cython_exts/LMC/PBCHelper.pyx"))


def is_built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.isfile(os.path.join(OUT, d, m + suffix)) for d, m in MODULES)


def build(force=False, verbose=True):
    if is_built() and not force:
        return True
    if not have_reference():
        if verbose:
            print("oracle/_ref: reference sources not present, nothing to build", file=sys.stderr)
        return False
    import numpy

    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    py_inc = sysconfig.get_paths()["include"]
    np_inc = numpy.get_include()
    stubs = os.path.join(OUT, "stubs")
    bdir = os.path.join(OUT, "build")
    os.makedirs(os.path.join(stubs, "cython_gsl"), exist_ok=True)
    os.makedirs(bdir, exist_ok=True)
    # the one GSL symbol the reference uses (numpyatom.pyx:8) == C99 round()
    with open(os.path.join(stubs, "cython_gsl", "__init__.pxd"), "w") as f:
        f.write('cdef extern from "math.h":\n    double round(double x) nogil\n')
    open(os.path.join(stubs, "cython_gsl", "__init__.py"), "w").close()
    # empty package skeleton so the compiled modules import as mdlmc.cython_exts.*
    for d in ("mdlmc", "mdlmc/cython_exts", "mdlmc/cython_exts/helper",
              "mdlmc/cython_exts/atoms", "mdlmc/cython_exts/LMC"):
        os.makedirs(os.path.join(OUT, d), exist_ok=True)
        open(os.path.join(OUT, d, "__init__.py"), "w").close()

    for d, m in MODULES:
        pyx = os.path.join(REF, d, m + ".pyx")
        cpp = os.path.join(bdir, m + ".cpp")
        so = os.path.join(OUT, d, m + suffix)
        cmd = [sys.executable, "-m", "cython", "--cplus", "-3", "-I", REF, "-I", stubs,
               "-o", cpp, pyx]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, cwd=REF, stdout=subprocess.DEVNULL if not verbose else None)
        cmd = ["g++", "-O3", "-ffast-math", "-w", "-shared", "-fPIC", "-std=c++17",
               "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
               "-I", py_inc, "-I", np_inc, "-I", REF, cpp, "-o", so, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "oracle/_ref NOT built")
