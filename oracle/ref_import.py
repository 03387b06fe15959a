"""Import helpers for the REAL reference (TEST INFRASTRUCTURE, never on the product path).

Two levels:

* `import_ref_atombox()`   -- the reference's compiled Cython AtomBox classes from oracle/_ref
                              (built by oracle/build_ref.py; travels to the GPU box).
* `import_ref_python()`    -- additionally the reference's pure-Python layers (topology.py,
                              MDMC.py, output.py, jumprate_generators.py ...) imported in place
                              from /root/reference.  Only possible in the build container; used
                              by oracle/make_golden.py and by the CPU tests that pin the oracle.

Compatibility shims: the reference targets Python 3.6 / NumPy 1.14.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CMDLMC_REFERENCE", "/root/reference")
REF_BUILD = os.path.join(HERE, "_ref")


def _numpy_shims():
    import warnings
    import numpy as np
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)
    for name, typ in (("int", int), ("float", float), ("bool", bool)):
        if name not in np.__dict__:
            setattr(np, name, typ)
    if "warnings" not in np.__dict__:
        np.warnings = warnings


def _stub_modules():
    # unconditional `import h5py` in mdlmc/IO/trajectory_parser.py:17; daiquiri only in tests
    if "h5py" not in sys.modules:
        try:
            importlib.import_module("h5py")
        except ImportError:
            sys.modules["h5py"] = types.ModuleType("h5py")
    if "tables" not in sys.modules:
        try:
            importlib.import_module("tables")
        except ImportError:
            sys.modules["tables"] = types.ModuleType("tables")


def ref_atombox_available():
    from . import build_ref
    return build_ref.is_built()


def ref_python_available():
    return ref_atombox_available() and os.path.isfile(os.path.join(REF, "mdlmc/LMC/MDMC.py"))


def import_ref_atombox():
    """Returns the module mdlmc.cython_exts.LMC.PBCHelper of the reference build."""
    if not ref_atombox_available():
        raise ImportError("oracle/_ref is not built (run python oracle/build_ref.py)")
    _numpy_shims()
    if REF_BUILD not in sys.path:
        sys.path.insert(0, REF_BUILD)
    mod = importlib.import_module("mdlmc.cython_exts.LMC.PBCHelper")
    assert os.path.abspath(mod.__file__).startswith(REF_BUILD), mod.__file__
    return mod


def import_ref_python():
    """Returns the reference's `mdlmc` package with compiled parts from oracle/_ref and Python
    parts imported in place from /root/reference (never copied)."""
    import_ref_atombox()
    if not ref_python_available():
        raise ImportError("reference python sources not present at %s" % REF)
    _stub_modules()
    import mdlmc
    ref_pkg = os.path.join(REF, "mdlmc")
    if ref_pkg not in list(mdlmc.__path__):
        mdlmc.__path__.append(ref_pkg)
    importlib.import_module("mdlmc.topo.topology")
    importlib.import_module("mdlmc.LMC.MDMC")
    importlib.import_module("mdlmc.LMC.output")
    importlib.import_module("mdlmc.LMC.jumprate_generators")
    importlib.import_module("mdlmc.IO.trajectory_parser")
    return mdlmc
