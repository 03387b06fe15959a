"""`jumpstat` -- jump probability as a function of the O-O distance (README.md:57-58 of the
reference describes the tool; its code is not in the tree).

Per distance bin:  p(d) = (#proton jumps over pairs at distance d) / (#pair-frames at distance d
that COULD have carried a jump) -- here normalised by all listed directed pair-frames, the
quantity the two device histograms deliver (k_jump_hist / k_pair_hist).
Replica ensembles make the numerator as smooth as wanted; both histograms are all-reduced over the
GPUs of the box."""
import argparse
import sys

import numpy as np

from .ensemble import run_kmc_ensemble


def jump_statistics(atom_box, frames_source, n_frames, *, n_sites, n_protons, cutoff, buffer, jumprate,
                    time_step, n_replicas=64, seed=0, lo=0.0, hi=None, nbins=100, chunk=1024):
    """Returns dict(edges, centers, pair_frames, jumps, probability).  `pair_frames` counts every
    listed directed pair of every frame once per replica (each replica sees every frame)."""
    hi = float(cutoff + buffer) if hi is None else float(hi)
    res = run_kmc_ensemble(atom_box, frames_source, n_frames, n_sites=n_sites, n_protons=n_protons,
                           cutoff=cutoff, buffer=buffer, jumprate=jumprate, time_step=time_step,
                           n_replicas=n_replicas, seed=seed, chunk=chunk, histogram=(lo, hi, nbins))
    edges = np.linspace(lo, hi, nbins + 1)
    pair_frames = res["pair_hist"].astype(np.float64) * res["n_replicas"]
    jumps = res["jump_hist"].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        prob = np.where(pair_frames > 0, jumps / pair_frames, 0.0)
    return dict(edges=edges, centers=0.5 * (edges[1:] + edges[:-1]), pair_frames=pair_frames,
                jumps=jumps, probability=prob, events=res["events"], n_replicas=res["n_replicas"])


def main(argv=None):
    """jumpstat <run.ini> [--replicas R] [--bins N]: same INI sections as `mdmc` (Trajectory,
    AtomBox, NeighborTopology, JumpRate, KMCLattice); prints `distance pair_frames jumps p`."""
    import configparser
    from . import main as driver
    ap = argparse.ArgumentParser(prog="jumpstat")
    ap.add_argument("configfile")
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--bins", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    cp = configparser.ConfigParser(inline_comment_prefixes=("#",))
    with open(args.configfile) as f:
        cp.read_file(f)
    traj = driver.build_section(cp, "Trajectory")
    box_opts = dict(cp["AtomBox"])
    pbc = np.array([float(x) for x in box_opts["periodic_boundaries"].strip("[]()").split(",")])
    box = driver.SECTIONS["AtomBox"][box_opts["type"]](pbc)
    rate = driver.build_section(cp, "JumpRate")
    topo_opts, kmc_opts = cp["NeighborTopology"], cp["KMCLattice"]
    donor = topo_opts["donor_atoms"]
    stat = jump_statistics(box, lambda a, b: traj.block(donor, a, b), len(traj),
                           n_sites=int(kmc_opts["lattice_size"]), n_protons=int(kmc_opts["proton_number"]),
                           cutoff=float(topo_opts.get("cutoff", 3.0)), buffer=float(topo_opts.get("buffer", 2.0)),
                           jumprate=rate, time_step=float(kmc_opts["time_step"]),
                           n_replicas=args.replicas, seed=args.seed, nbins=args.bins)
    print("# distance pair_frames jumps probability   (%d replicas, %d jumps)" % (stat["n_replicas"],
                                                                                 stat["events"]))
    for c, n, j, p in zip(stat["centers"], stat["pair_frames"], stat["jumps"], stat["probability"]):
        print("%.5f %d %d %.6e" % (c, n, j, p))


if __name__ == "__main__":
    main(sys.argv[1:])
