"""Drop-in registration under the reference's module paths.

`install()` puts thin module objects into sys.modules so that code written against the reference
-- `from mdlmc.cython_exts.LMC.PBCHelper import AtomBoxCubic`, `from mdlmc.topo.topology import
NeighborTopology`, `from mdlmc.LMC.MDMC import KMCLattice`, ... -- resolves to the GPU-backed
classes of this package.  Nothing of the reference tree is modified or required; modules of the
reference that are already imported are left alone unless `force=True`."""
import sys
import types

MODULES = {
    "mdlmc.cython_exts.LMC.PBCHelper": ("atombox", ["AtomBox", "AtomBoxCubic", "AtomBoxMonoclinic",
                                                    "AtomBoxWater", "AtomBoxWaterLinearConversion",
                                                    "AtomBoxWaterRampConversion"]),
    "mdlmc.topo.topology": ("topology", ["NeighborTopology", "AngleTopology", "HydroniumTopology",
                                         "DistanceTransformation", "ReLUTransformation",
                                         "InterpolatedTransformation", "DistanceInterpolator"]),
    "mdlmc.LMC.jumprate_generators": ("jumprate", ["JumpRate", "Fermi", "FermiAngle"]),
    "mdlmc.LMC.MDMC": ("kmc", ["KMCLattice", "Output", "XYZOutput", "ObservablesOutput"]),
    "mdlmc.LMC.output": ("output", ["CovalentAutocorrelation", "MeanSquareDisplacement"]),
    "mdlmc.IO.trajectory_parser": ("trajectory", ["Frame", "Trajectory", "XYZTrajectory",
                                                  "HDF5Trajectory"]),
    "mdlmc.main": ("main", ["main"]),
}


def install(force=False):
    """Registers the shims; returns the list of module names that now point at this package."""
    import importlib
    done = []
    for pkg in ("mdlmc", "mdlmc.cython_exts", "mdlmc.cython_exts.LMC", "mdlmc.topo", "mdlmc.LMC",
                "mdlmc.IO"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    for name, (src, symbols) in MODULES.items():
        if name in sys.modules and not force:
            continue
        source = importlib.import_module("cmdlmc_b200." + src)
        mod = types.ModuleType(name)
        mod.__doc__ = "cmdlmc_b200 shim for %s" % name
        for sym in symbols:
            setattr(mod, sym, getattr(source, sym))
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
        done.append(name)
    return done
