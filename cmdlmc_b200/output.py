"""Observables -- host mirror of mdlmc/LMC/output.py:6-49.

The device pipeline (csrc/kmc.cu kmc_observe) computes the same quantities per replica inside
the KMC kernel; these classes keep the reference's small API for callers that update the
observables themselves.  Distances go through AtomBox.distance, i.e. the CUDA kernels."""
import numpy as np


class CovalentAutocorrelation:
    """Number of sites that still hold the proton label they held at the last reset."""

    def __init__(self, lattice):
        self.reset(lattice)

    def reset(self, lattice):
        self.lattice = np.array(lattice, copy=True)

    def calculate(self, lattice):
        lattice = np.asarray(lattice)
        return np.sum((lattice == self.lattice) & (lattice != 0))


class MeanSquareDisplacement:
    """Per-axis MSD of the protons with periodic unwrapping: every update adds the minimum-image
    vector between a proton's previous and current site position."""

    def __init__(self, atom_positions, lattice, atombox):
        lattice = np.asarray(lattice)
        n_protons = int(np.sum(lattice > 0))
        self.snapshot = np.zeros((n_protons, 3))
        self.displacement = np.zeros((n_protons, 3))
        self.snapshot = self.determine_proton_positions(atom_positions, lattice)
        self.atombox = atombox

    def determine_proton_positions(self, atom_positions, lattice):
        """Row (label - 1) holds the position of the site carrying proton `label`."""
        lattice = np.asarray(lattice)
        sites = np.flatnonzero(lattice)
        out = np.zeros_like(self.snapshot)
        out[lattice[sites] - 1] = np.asarray(atom_positions)[sites]
        return out

    def update_proton_positions(self, atom_positions, lattice):
        self.snapshot[:] = self.determine_proton_positions(atom_positions, lattice)

    def update_displacement(self, new_positions, lattice):
        new = self.determine_proton_positions(new_positions, lattice)
        step = self.atombox.distance(self.snapshot, new)
        self.displacement += np.reshape(step, self.displacement.shape)
        self.snapshot = new

    def reset_displacement(self):
        self.displacement[:] = 0

    def msd(self):
        return np.sum(self.displacement ** 2, axis=0) / self.displacement.shape[0]


def diffusion_coefficient(rows, reset_frequency, fit_start=0):
    """Interval averaging of printed observables, the job of mdlmc/LMC/average_MC_out.py
    (`avg` :115-125, `get_slope` :149-205, its default branch): the rows (frame, time, msd_x,
    msd_y, msd_z, autocorr) are cut into the intervals between MSD resets, a line m t + y0 is
    fitted to the summed MSD of every interval from `fit_start` on, and D = mean(m) / 6 with the
    standard deviation of the slopes as its error.  Host-side post-processing of a few rows."""
    rows = np.asarray(rows, dtype=float)
    if rows.ndim != 2 or rows.shape[1] < 5:
        raise ValueError("rows must be [n, >= 5]: frame, time, msd_x, msd_y, msd_z, ...")
    interval = (rows[:, 0] // reset_frequency).astype(np.int64)
    slopes, offsets = [], []
    for k in np.unique(interval):
        part = rows[interval == k][fit_start:]
        if len(part) < 2:
            continue
        t = part[:, 1] - part[0, 1]
        y = part[:, 2:5].sum(axis=1)
        m, y0 = np.polyfit(t, y, 1)
        slopes.append(m)
        offsets.append(y0)
    if not slopes:
        raise ValueError("no interval holds two or more rows")
    slopes = np.asarray(slopes)
    return dict(slope=float(slopes.mean()), slope_err=float(slopes.std()),
                diffusion_coefficient=float(slopes.mean() / 6), error=float(slopes.std() / 6),
                intervals=len(slopes), offset=float(np.mean(offsets)))
