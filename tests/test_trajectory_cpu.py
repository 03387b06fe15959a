"""Row F3, host side: the reference's own parser tests (tests/IO/test_parsers.py:50-103) run
against this repo's Frame / XYZTrajectory, and its chunking helper test
(tests/misc/test_tools.py:29-44) against the block interface the GPU pipeline uploads from."""
from io import StringIO

import numpy as np
import pytest

from cmdlmc_b200.trajectory import ArrayTrajectory, Frame, XYZTrajectory

dtype_xyz = np.dtype([("name", np.str_, 2), ("pos", np.float64, (3,))])   # atoms/numpy_atom.py

MOCK_XYZ = """
3
comment
O 0 0 0
H 0 1 0
H 1 0 0
3
comment
O 0 0 0
H 0 1 0
H 1.2 0 0
3
comment
O 0 0 0
H 0 1 0
H 1.4 0 0
""".strip()


@pytest.fixture
def xyz_array():
    return np.array([("O", [0, 0, 0]), ("H", [0, 1, 0]), ("H", [1, 0, 0])], dtype=dtype_xyz)


def test_frame(xyz_array):                                   # test_parsers.py:50-59
    frame = Frame.from_recarray(xyz_array, time=0.5)
    np.testing.assert_equal(frame["H"].atom_positions, xyz_array["pos"][xyz_array["name"] == "H"])
    np.testing.assert_equal(frame[[0, -1]].atom_positions, xyz_array["pos"][[0, -1]])
    assert frame.atom_number == 3
    assert frame.time == 0.5
    with pytest.raises(ValueError):
        frame[3.5]


def test_frame_append(xyz_array):                            # test_parsers.py:62-69
    f1 = Frame.from_recarray(xyz_array)
    f2 = Frame.from_recarray(xyz_array)
    result = f1.append(f2)
    assert result.atom_number == 6
    np.testing.assert_array_equal(result.atom_names, ["O", "H", "H", "O", "H", "H"])


def test_xyz_trajectory():                                   # test_parsers.py:78-86
    parser = XYZTrajectory(StringIO(MOCK_XYZ), number_of_atoms=3, time_step=0.5)
    frames = list(parser)
    for frame in frames:
        assert frame.atom_names.shape == (3,)
    assert len(frames) == 3
    assert [f.time for f in frames] == [0.0, 0.5, 1.0]
    np.testing.assert_array_equal(frames[2].atom_positions[2], [1.4, 0, 0])


@pytest.mark.parametrize("selection, expected_shape", [((0, 2), (2,)), (("O", "H"), (3,))])
def test_xyz_selection(selection, expected_shape):           # test_parsers.py:89-103
    parser = XYZTrajectory(StringIO(MOCK_XYZ), number_of_atoms=3, selection=selection, time_step=0.5)
    frames = list(parser)
    assert frames[0].atom_names.shape == expected_shape


def test_blocks_tile_the_trajectory_like_chunk_trajectory():  # test_tools.py:29-44
    """chunk_trajectory yields (start, stop, frames[start:stop]); the block interface serves the
    same slices (donor rows only, stored precision kept)."""
    rng = np.random.RandomState(0)
    names = np.array(["O", "H", "O", "H", "P"])
    pos = rng.normal(size=(23, 5, 3)).astype(np.float32)
    traj = ArrayTrajectory(pos, names, time_step=0.4)
    assert len(traj) == 23
    got = []
    for start in range(0, 23, 10):
        stop = min(start + 10, 23)
        blk = traj.block("O", start, stop)
        assert blk.dtype == np.float32 and blk.shape == (stop - start, 2, 3)
        got.append(np.array(blk))
    np.testing.assert_array_equal(np.concatenate(got), pos[:, names == "O"])
    frames = list(traj)
    assert len(frames) == 23 and frames[5].time == pytest.approx(2.0)
    assert traj.current_frame_number == 22
    with pytest.raises(ValueError):
        ArrayTrajectory(pos[:, :, :2], names, time_step=0.4)
