"""Row F4: interval averaging / diffusion-coefficient fit of printed observables against the
reference's analysis script, restated with the same SciPy call.

mdlmc/LMC/average_MC_out.py cannot be imported here (it needs numba, pint and matplotlib), so its
default branch is restated literally below: `load_intervals_intelligently` reshapes the rows into
[interval_number, interval_length, 7] (:103-111), `get_slope` (:149-186) takes the time column of
the FIRST interval, fits m x + y with scipy.optimize.curve_fit to the summed MSD of every interval
from `msd_fitstart` on, and reports mean and standard deviation of the slopes; D = m / 6 (:190)."""
import numpy as np
import pytest

from cmdlmc_b200.output import diffusion_coefficient


def reference_get_slope(rows, interval_length, fit_start):
    from scipy.optimize import curve_fit
    interval_number = rows.shape[0] // interval_length
    data = rows[:interval_number * interval_length].reshape(interval_number, interval_length, rows.shape[1])
    time = data[0, :, 1]
    ms, y0s = [], []
    for interval in data:
        y = interval[:, 2:5].sum(axis=-1)
        (m, y0), _ = curve_fit(lambda x, m, y: m * x + y, time[fit_start:], y[fit_start:])
        ms.append(m)
        y0s.append(y0)
    ms = np.asarray(ms, dtype=float)
    return ms.mean(), np.std(ms)


@pytest.mark.parametrize("fit_start", [0, 3])
def test_diffusion_coefficient_matches_the_reference_fit(fit_start):
    rng = np.random.RandomState(5)
    reset_frequency, print_frequency, sweeps, dt = 2000, 100, 20000, 0.5
    rows = []
    for frame in range(0, sweeps, print_frequency):
        t_in = (frame % reset_frequency) * dt            # the MSD restarts at every reset
        slope = 3e-4 * (1 + 0.2 * rng.normal())
        msd = np.abs(slope * t_in * np.array([0.5, 0.3, 0.2]) + 1e-3 * rng.normal(size=3)) * (t_in > 0)
        rows.append([frame, frame * dt, *msd, 0.5 + 0.01 * rng.normal(), 96])
    rows = np.array(rows)
    want_m, want_err = reference_get_slope(rows, reset_frequency // print_frequency, fit_start)
    got = diffusion_coefficient(rows, reset_frequency, fit_start=fit_start)
    assert got["intervals"] == sweeps // reset_frequency
    assert got["slope"] == pytest.approx(want_m, rel=1e-6)
    assert got["slope_err"] == pytest.approx(want_err, rel=1e-5)
    assert got["diffusion_coefficient"] == pytest.approx(want_m / 6, rel=1e-6)
    assert got["error"] == pytest.approx(want_err / 6, rel=1e-5)


def test_diffusion_coefficient_rejects_rows_without_msd_columns():
    with pytest.raises(ValueError):
        diffusion_coefficient(np.zeros((4, 3)), 10)


class _FakeH5File(dict):
    """Just enough of h5py.File for HDF5Trajectory: datasets that slice like arrays."""

    def __init__(self, filename, mode="r"):
        super().__init__(np.load(filename))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def test_hdf5_trajectory_reads_the_reference_layout_chunk_by_chunk(tmp_path, monkeypatch):
    """Row F3: HDF5Trajectory (trajectory_parser.py:290-337) on the reference's file layout
    (IO/converters.py:38-43: `trajectory` float32 [frames, atoms, 3], `atom_names`).  h5py is not
    in this image; a stand-in with the same slicing interface serves the arrays."""
    import sys
    import types
    from cmdlmc_b200.trajectory import HDF5Trajectory
    rng = np.random.RandomState(3)
    names = np.array(["O", "H", "O", "P", "O", "H"])
    traj = rng.uniform(0, 10, size=(25, names.size, 3)).astype(np.float32)
    fn = tmp_path / "t.npz"
    np.savez(fn, trajectory=traj, atom_names=names.astype("S2"))
    fake = types.ModuleType("h5py")
    fake.File = _FakeH5File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    t = HDF5Trajectory(str(fn), time_step=0.5, chunk_size=10)
    assert len(t) == 25 and list(t.atom_names) == list(names)
    frames = list(t)
    assert len(frames) == 25 and frames[7].time == pytest.approx(3.5)
    assert frames[7].atom_positions.dtype == np.float64
    np.testing.assert_array_equal(frames[7]["O"].atom_positions, traj[7][names == "O"].astype(float))
    blk = t.block("O", 5, 17)                      # what the GPU pipeline uploads: stored precision
    assert blk.dtype == np.float32 and blk.shape == (12, 3, 3)
    np.testing.assert_array_equal(blk, traj[5:17][:, names == "O"])
    with pytest.warns(UserWarning):
        HDF5Trajectory(str(fn), time_step=0.5, selection="O")
