"""`mdmc` driver on top of the GPU pipeline -- same INI sections and keys as the reference's
mdlmc/main.py:56-158 ([Trajectory] [AtomBox] [NeighborTopology] [JumpRate] [KMCLattice] [Output],
`type=` selects the class, every other key is a constructor argument converted by the parameter's
annotation), Python-3.12-safe, and with the sub-commands README.md:54-55 documents:

    python -m cmdlmc_b200.main config_load <file.ini>     (or just <file.ini>)
    python -m cmdlmc_b200.main config_help

Each output item is printed like upstream (`for x in output: print(x)`, main.py:157-158)."""
import argparse
import configparser
import inspect
import logging
import sys
import typing

import numpy as np

from . import atombox, jumprate, kmc, topology, trajectory

SECTIONS = {
    # section -> (registry of `type` values, positional objects it is built from)
    "Trajectory": {"XYZTrajectory": trajectory.XYZTrajectory, "NpzTrajectory": trajectory.NpzTrajectory,
                   "HDF5Trajectory": trajectory.HDF5Trajectory},
    "AtomBox": {"AtomBoxCubic": atombox.AtomBoxCubic, "AtomBoxMonoclinic": atombox.AtomBoxMonoclinic},
    "NeighborTopology": {"NeighborTopology": topology.NeighborTopology,
                         "AngleTopology": topology.AngleTopology,
                         "HydroniumTopology": topology.HydroniumTopology},
    "DistanceTransformation": {"ReLUTransformation": topology.ReLUTransformation,
                               "InterpolatedTransformation":
                                   topology.InterpolatedTransformation.from_file},
    "DistanceInterpolator": {"DistanceInterpolator": topology.DistanceInterpolator},
    "JumpRate": {"Fermi": jumprate.Fermi, "FermiAngle": jumprate.FermiAngle,
                 "ActivationEnergy": jumprate.ActivationEnergy, "Exponential": jumprate.Exponential},
    "KMCLattice": {"KMCLattice": kmc.KMCLattice},
    "Output": {"XYZOutput": kmc.XYZOutput, "ObservablesOutput": kmc.ObservablesOutput},
}


def _convert(value, annotation):
    """One INI string -> the annotated type (unions: first member that accepts it)."""
    if value == "EMPTY":
        raise ValueError("a key is EMPTY: please specify a value in the config file")
    if value == "None":
        return None
    if annotation is inspect.Parameter.empty or isinstance(annotation, str):
        return value
    members = typing.get_args(annotation) if typing.get_origin(annotation) is typing.Union else ()
    for typ in members or (annotation,):
        if typ is bool:
            return value.strip().lower() in ("1", "true", "yes", "on")
        if typ in (int, float, str):
            try:
                return typ(value)
            except (TypeError, ValueError):
                continue
    return value


def build_section(cp, section, *args, **fixed):
    opts = dict(cp[section])
    registry = SECTIONS[section]
    cls = registry[opts.pop("type")] if "type" in opts else next(iter(registry.values()))
    params = inspect.signature(cls).parameters
    kw = {}
    for key, val in opts.items():
        if key not in params:
            raise KeyError("[%s] %s is not a parameter of %s" % (section, key, cls.__name__))
        kw[key] = _convert(val, params[key].annotation)
    kw.update(fixed)
    return cls(*args, **kw)


def config_help(out=sys.stdout):
    """INI template of every class the driver can build (the job of mdlmc_config upstream)."""
    for section, registry in SECTIONS.items():
        for name, cls in registry.items():
            out.write("[%s]\ntype = %s\n" % (section, name))
            skip = set(getattr(cls, "__no_config_parameter__", [])) | {"self", "args", "kwargs"}
            for pname, p in inspect.signature(cls).parameters.items():
                if pname in skip or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
                    continue
                default = "EMPTY" if p.default is p.empty else p.default
                ann = getattr(p.annotation, "__name__", str(p.annotation))
                out.write("%s = %s    # %s\n" % (pname, default, ann))
            out.write("\n")


def run(configfile, out=sys.stdout):
    cp = configparser.ConfigParser(inline_comment_prefixes=("#",))
    with open(configfile, "r") as f:
        cp.read_file(f)
    if "Logging" in cp:
        logging.basicConfig(level=cp["Logging"]["level"])
    traj = build_section(cp, "Trajectory")
    box_opts = dict(cp["AtomBox"])
    pbc = np.array([float(x) for x in box_opts["periodic_boundaries"].strip("[]()").split(",")])
    box = SECTIONS["AtomBox"][box_opts["type"]](pbc)
    extra = {}
    if cp["NeighborTopology"].get("type") == "HydroniumTopology":   # main.py:89-127 upstream
        if "DistanceTransformation" not in cp:
            raise NameError("Distance Transformation needs to be specified!")
        extra["distance_transformation_function"] = build_section(cp, "DistanceTransformation")
        extra["distance_interpolator"] = (build_section(cp, "DistanceInterpolator")
                                          if "DistanceInterpolator" in cp else None)
    topo = build_section(cp, "NeighborTopology", traj, box, **extra)
    rate = build_section(cp, "JumpRate")
    lattice = build_section(cp, "KMCLattice", topo, jumprate_function=rate, atom_box=box)
    output = build_section(cp, "Output", lattice)
    for x in output:
        print(x, file=out)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="mdmc", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("command", help="config_load <file>, config_help, or directly a config file")
    ap.add_argument("configfile", nargs="?", help="INI file holding the cMD/LMC configuration")
    args = ap.parse_args(argv)
    if args.command == "config_help":
        config_help()
    elif args.command == "config_load":
        if not args.configfile:
            ap.error("config_load needs a config file")
        run(args.configfile)
    else:
        run(args.command)


if __name__ == "__main__":
    main()
