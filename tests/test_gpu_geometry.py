"""GPU parity, rows A1-A6 + A9 of SURVEY.md section 8: the CUDA AtomBox / jump-rate kernels,
called through the C ABI, against (1) the CPU oracle on seeded inputs -- BIT-EXACT, because the
kernels mirror the reference arithmetic operation by operation without FMA contraction -- and
(2) the golden vectors produced by the real reference (1e-12 relative: the reference itself is
built with -ffast-math)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CELLS = ["ortho", "cubic10", "diag9", "mono", "tri", "tri2"]
RTOL = 1e-12


def make_box(cell, **kw):
    import cmdlmc_b200 as cm
    cell = np.asarray(cell, dtype=float)
    return cm.AtomBoxCubic(cell, **kw) if cell.size == 3 else cm.AtomBoxMonoclinic(cell, **kw)


@pytest.mark.parametrize("name", CELLS)
def test_golden_reference_vectors(golden, name):
    g = golden("geometry")
    box = make_box(g[name + "_cell"])
    a, b, c = g[name + "_a"], g[name + "_b"], g[name + "_c"]
    np.testing.assert_allclose(box.length(a, b), g[name + "_length"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(box.distance(a, b), g[name + "_distance"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(box.angle(a, b, c), g[name + "_angle"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(box.length_all_to_all(a[:40], b[:50]), g[name + "_all"],
                               rtol=RTOL, atol=0)
    for i in range(20):
        idx, dist = box.next_neighbor(a[i], b[:60])
        assert idx == g[name + "_nn_idx"][i]
        assert dist == pytest.approx(g[name + "_nn_dist"][i], rel=RTOL)


@pytest.mark.parametrize("name", CELLS)
@pytest.mark.parametrize("n", [1, 31, 1000, 70001])
def test_bit_exact_vs_oracle(golden, orc, name, n):
    g = golden("geometry")
    cell = g[name + "_cell"]
    box, obox = make_box(cell), orc.OracleBox(cell)
    rng = np.random.RandomState(n)
    scale = 4.0 * np.abs(cell).max()
    a = rng.uniform(-scale, scale, size=(n, 3))
    b = rng.uniform(-scale, scale, size=(n, 3))
    c = rng.uniform(-scale, scale, size=(n, 3))
    np.testing.assert_array_equal(box.length(a, b), obox.length(a, b))
    np.testing.assert_array_equal(box.distance(a, b), obox.distance(a, b))
    got, want = box.angle(a, b, c), obox.angle(a, b, c)
    np.testing.assert_allclose(got, np.atleast_1d(want), rtol=0, atol=4e-15)  # acos: <= 2 ulp
    m = min(n, 300)
    np.testing.assert_array_equal(box.length_all_to_all(a[:m], b[:257]),
                                  obox.length_all_to_all(a[:m], b[:257]))
    for i in range(min(n, 5)):
        assert box.next_neighbor(a[i], b[:5000]) == obox.next_neighbor(a[i], b[:5000])


def test_reference_known_answers():
    """tests/cython_exts/LMC/test_AtomBox.py:19-75,143-174 against the CUDA classes."""
    import cmdlmc_b200 as cm
    box = cm.AtomBoxCubic([10.0, 10, 10])
    mono = cm.AtomBoxMonoclinic(np.array([10.0, 0, 0, 0, 10, 0, 0, 0, 10]))
    a1, a2 = np.zeros(3), np.array([6.0, 6, 6])
    for i in range(-5, 5):
        assert box.length(a1, a2 + i * 10) == pytest.approx(np.sqrt(48))
    atoms_2 = np.arange(-10, 10)[:, None] * np.array([10.0, 10, 10]) + 3
    assert np.isclose(box.length(np.zeros((20, 3)), atoms_2), np.sqrt(27)).all()
    assert np.allclose(box.distance(a1, a2), [-4, -4, -4])
    t = np.array([[1.0, 1, 1], [2, 2, 2], [3, 3, 3]])
    assert np.allclose(box.distance(np.zeros((3, 3)), t), t)
    assert box.angle(np.zeros(3), np.array([3.0, 0, 0]), np.array([3.0, 34, 0])) == \
        pytest.approx(np.pi / 2)
    rng = np.random.RandomState(0)
    big = cm.AtomBoxCubic([100.0, 100, 100])
    atoms = rng.uniform(0.3, 50, size=(20, 3))
    for _ in range(10):
        atom = rng.uniform(0, 50, size=3)
        index, _ = big.next_neighbor(atom, atoms)
        assert index == np.argmin(np.sqrt(((atom - atoms) ** 2).sum(axis=-1)))
    p1, p2, p3 = (rng.uniform(-10, 10, size=(10, 3)) for _ in range(3))
    for i in range(10):
        assert np.isclose(box.distance(p1[i], p2[i]), mono.distance(p1[i], p2[i])).all()
        assert np.allclose(box.length(p1[i], p2[i]), mono.length(p1[i], p2[i]))
        assert box.angle(p1[i], p2[i], p3[i]) == pytest.approx(mono.angle(p1[i], p2[i], p3[i]))
    atoms = np.array([[0.0, 0, 0], [1, 1, 1], [5, 5, 5], [10, 10, 10]])
    s3 = np.sqrt(3)
    want = np.array([[0, s3, 5 * s3, 0], [s3, 0, 4 * s3, s3], [5 * s3, 4 * s3, 0, 5 * s3],
                     [0, s3, 5 * s3, 0]])
    np.testing.assert_allclose(box.length_all_to_all(atoms, atoms), want)


def test_ties_and_attributes(golden):
    import cmdlmc_b200 as cm
    g = golden("geometry")
    z = np.zeros((2, 3))
    t = np.array([[5.0, 5.0, 5.0], [-5.0, -5.0, -5.0]])
    np.testing.assert_array_equal(cm.AtomBoxCubic(g["cubic10_cell"]).distance(z, t), g["tie_ortho"])
    np.testing.assert_array_equal(cm.AtomBoxMonoclinic(g["diag9_cell"]).distance(z, t),
                                  g["tie_general"])
    tri = cm.AtomBoxMonoclinic(g["tri_cell"])
    np.testing.assert_array_equal(tri.pbc_matrix, g["tri_cell"].reshape(3, 3))
    np.testing.assert_array_equal(tri.h, g["tri_cell"].reshape(3, 3).T)
    np.testing.assert_array_equal(tri.h_inv, np.linalg.inv(tri.h))


def test_extended_box(golden, orc):
    import cmdlmc_b200 as cm
    g = golden("geometry")
    box = cm.AtomBoxCubic(np.array([10.0, 10, 10]), box_multiplier=(2, 3, 4))
    np.testing.assert_array_equal(box.periodic_boundaries_extended, g["ext_pbc"])
    pos = np.array([box.position_extended_box(i, g["ext_frame"]) for i in range(5 * 24)])
    np.testing.assert_allclose(pos, g["ext_pos"], rtol=0, atol=1e-13)
    # tests/cython_exts/LMC/test_AtomBox.py:77-120
    atom1 = np.zeros((1, 3))
    b5 = cm.AtomBoxCubic([10.0, 10, 10], box_multiplier=(5, 5, 5))
    index = 0
    for i in range(5):
        for j in range(5):
            for k in range(5):
                np.testing.assert_allclose(b5.position_extended_box(index, atom1),
                                           [10.0 * i, 10.0 * j, 10.0 * k])
                index += 1
    # next_neighbor_extended_box: (0,0,0) image vs atoms at z = 9 in a 1x1x5 box
    b = cm.AtomBoxCubic([10.0, 10, 10], box_multiplier=(1, 1, 5))
    idx, dist = b.next_neighbor_extended_box(0, np.zeros((1, 3)), np.array([[0.0, 0, 9]]))
    assert (idx, dist) == (4, 1.0)   # image k=4 sits at z=49 == -1 in the 50 A box


def test_water_conversions():
    import cmdlmc_b200 as cm
    a, b, d0, lb, rb = 0.5, 2.3, 2.45, 2.3, 3.33
    par = dict(a=a, b=b, d0=d0, left_bound=lb, right_bound=rb)
    box = cm.AtomBoxCubic([10.0, 10, 10])
    ramp = cm.AtomBoxWaterRampConversion([10.0, 10, 10], par)
    z = np.zeros((1, 3))
    len1 = float(box.length(z, np.array([2.7, 0, 0]))[0])
    assert a * (len1 - d0) + b == float(ramp.length(z, np.array([2.7, 0, 0]))[0])
    assert b == float(ramp.length(z, np.array([2.4, 0, 0]))[0])
    lin = cm.AtomBoxWaterLinearConversion([10.0, 10, 10], dict(a=0.5, b=1.1, left_bound=2.2,
                                                                right_bound=3.3))
    assert float(lin.length(z, np.array([2.5, 0, 0]))[0]) == pytest.approx(0.5 * 2.5 + 1.1)
    rng = np.random.RandomState(0)
    atoms2 = np.zeros((100, 3))
    atoms2[:, 2] = rng.uniform(2.343, 2.9, size=100)
    par2 = dict(a=0.97672, b=2.342541, d0=2.578514, left_bound=2.34, right_bound=3.058)
    diffs = cm.AtomBoxWaterRampConversion([100.0, 100, 100], par2).length(np.zeros((100, 3)), atoms2)
    assert (diffs <= atoms2[:, 2]).all()


def test_rates_vs_oracle(orc):
    import cmdlmc_b200 as cm
    x = np.random.RandomState(0).uniform(1.5, 6.0, size=20000)
    th = np.random.RandomState(1).uniform(0, np.pi, size=20000)
    np.testing.assert_allclose(cm.Fermi(0.06, 2.3, 0.1)(x), orc.rates("Fermi", (0.06, 2.3, 0.1), x),
                               rtol=1e-10, atol=0)
    got = cm.FermiAngle(0.06, 2.3, 0.1, np.pi / 2)(x, th)
    want = orc.rates("FermiAngle", (0.06, 2.3, 0.1, np.pi / 2), x, th)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=0)
    assert ((got == 0) == (th < np.pi / 2)).all()
    par = (0.06, 1.2, 30.0, 2.2, 510.0)
    np.testing.assert_allclose(cm.ActivationEnergy(*par)(x), orc.rates("ActivationEnergy", par, x),
                               rtol=1e-10, atol=0)
    np.testing.assert_allclose(cm.Exponential(2.0, -1.5)(x), orc.rates("Exponential", (2.0, -1.5), x),
                               rtol=1e-10, atol=0)
    # reference formula straight from jumprate_generators.py:33-34
    np.testing.assert_allclose(cm.Fermi(0.06, 2.3, 0.1)(x), 0.06 / (1 + np.exp((x - 2.3) / 0.1)),
                               rtol=1e-10)


def test_empty_and_errors():
    import cmdlmc_b200 as cm
    box = cm.AtomBoxCubic([10.0, 10, 10])
    assert box.length(np.zeros((0, 3)), np.zeros((0, 3))).shape == (0,)
    assert box.length_all_to_all(np.zeros((0, 3)), np.zeros((4, 3))).shape == (0, 4)
    with pytest.raises(ValueError):
        cm.AtomBoxCubic([10.0, 10, 10, 5])
    with pytest.raises(cm._abi.CmdError):
        cm.AtomBoxCubic([10.0, -1, 10])
    with pytest.raises(ValueError):
        box.length(np.zeros((2, 3)), np.zeros((3, 3)))


def test_mean_square_displacement_host_mirror_like_the_reference_tests():
    """tests/LMC/test_output.py:21-47 on the host mirror of mdlmc/LMC/output.py (its distances go
    through AtomBox.distance, i.e. the CUDA kernels), plus the covalent autocorrelation (:6-14)."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.output import CovalentAutocorrelation, MeanSquareDisplacement
    atom_positions = np.arange(1, 19).reshape(6, 3)
    lattice = np.array([0, 3, 0, 0, 1, 2])
    atombox = cm.AtomBoxCubic([10, 10, 10])
    msd = MeanSquareDisplacement(atom_positions, lattice, atombox=atombox)
    np.testing.assert_equal(msd.snapshot, np.array([[13, 14, 15], [16, 17, 18], [4, 5, 6]]))
    auto = CovalentAutocorrelation(lattice)
    assert auto.calculate(lattice) == 3
    # protons 1 and 2 swap positions
    lattice[-2], lattice[-1] = lattice[-1], lattice[-2]
    msd.update_displacement(atom_positions, lattice)
    displacement = np.zeros((3, 3), int)
    displacement[0] = [3, 3, 3]
    displacement[1] = [-3, -3, -3]
    np.testing.assert_equal(msd.displacement, displacement)
    assert auto.calculate(lattice) == 1
    # proton 2 jumps to an empty site
    lattice[-2], lattice[-3] = lattice[-3], lattice[-2]
    msd.update_displacement(atom_positions, lattice)
    displacement[1] += np.array([-3, -3, -3])
    np.testing.assert_equal(msd.displacement, displacement)
    np.testing.assert_allclose(msd.msd(), (displacement ** 2).sum(axis=0) / 3)
    msd.reset_displacement()
    assert (msd.displacement == 0).all()
