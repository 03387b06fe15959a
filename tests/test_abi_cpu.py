"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cmdlmc_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cmdlmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmd_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from cmdlmc_b200 import build, _abi
    build.build()
    return _abi.lib()


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) > 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header(lib):
    from cmdlmc_b200 import _abi
    assert set(declared_symbols()) <= set(_abi.SIGNATURES), \
        sorted(set(declared_symbols()) - set(_abi.SIGNATURES))
    assert lib.cmd_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import cmdlmc_b200 as cm
    from cmdlmc_b200._abi import CmdError
    with pytest.raises(CmdError) as e:
        cm.AtomBoxCubic([10.0, 10, 10])
    assert e.value.code == -2
    with pytest.raises(CmdError):
        cm.Fermi(0.06, 2.3, 0.1)(np.array([2.5]))


def test_product_never_imports_oracle():
    """The product package must never import, load or link the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "cmdlmc_b200")
    bad = re.compile(r"(^\s*(from|import)\s+(\.*)oracle)|libcmdlmc_oracle|#include\s*[\"<].*oracle|"
                     r"oracle\.(oracle|ref_import)|orc_[a-z_]+\s*\(", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)


def test_compat_shims_and_config_help(tmp_path):
    """Host logic only (no compute call): the reference's import paths resolve to this package's
    classes with the reference's constructor signatures, and the driver's INI template lists
    every class main.py:73-155 can build."""
    import inspect
    import io
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from cmdlmc_b200 import compat\n"
        "done = compat.install()\n"
        "from mdlmc.cython_exts.LMC.PBCHelper import AtomBoxCubic, AtomBoxMonoclinic\n"
        "from mdlmc.topo.topology import NeighborTopology, AngleTopology\n"
        "from mdlmc.LMC.jumprate_generators import Fermi, FermiAngle\n"
        "from mdlmc.LMC.MDMC import KMCLattice, XYZOutput, ObservablesOutput\n"
        "from mdlmc.IO.trajectory_parser import Frame, XYZTrajectory\n"
        "import cmdlmc_b200\n"
        "assert AtomBoxCubic is cmdlmc_b200.AtomBoxCubic and len(done) == 7\n"
        "print('ok')\n" % str(__import__('os').path.dirname(__import__('os').path.dirname(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
    from cmdlmc_b200 import kmc, main, topology
    sig = inspect.signature(kmc.KMCLattice).parameters
    for key in ("lattice_size", "proton_number", "donor_atoms", "time_step", "extra_atoms"):
        assert key in sig                       # MDMC.py:34-41
    sig = inspect.signature(topology.AngleTopology).parameters
    assert [k for k in sig][:2] == ["trajectory", "atom_box"] and "group_size" in sig
    buf = io.StringIO()
    main.config_help(buf)
    text = buf.getvalue()
    for word in ("[Trajectory]", "XYZTrajectory", "AtomBoxMonoclinic", "AngleTopology", "FermiAngle",
                 "KMCLattice", "ObservablesOutput", "reset_frequency"):
        assert word in text
    assert main._convert("3", int) == 3 and main._convert("None", float) is None
    import typing
    assert main._convert("0.5", typing.Union[int, float]) == 0.5
    # xyz reader: two frames, selection by name
    from cmdlmc_b200.trajectory import XYZTrajectory
    p = tmp_path / "t.xyz"
    p.write_text("3\n\nO 0 0 0\nH 1 0 0\nO 2 0 0\n3\n\nO 0 1 0\nH 1 1 0\nO 2 1 0\n")
    t = XYZTrajectory(str(p), time_step=0.5, selection="O")
    frames = list(t)
    assert len(frames) == 2 and frames[1].atom_positions.tolist() == [[0, 1, 0], [2, 1, 0]]
    assert frames[1].time == 0.5 and list(frames[0].atom_names) == ["O", "O"]


def test_reference_arm_prints_one_json_line():
    """bench.py --impl reference: the CPU arm (the reference's compiled AtomBox from oracle/_ref
    when present, else the oracle's C port) prints one JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--cpu-seconds", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames*O-pairs/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_distance_transformations_host_mirrors():
    """ReLUTransformation / InterpolatedTransformation / DistanceInterpolator host classes
    (topology.py:273-353): known answers and scipy's interp1d, which the reference uses."""
    import numpy as np
    from scipy.interpolate import interp1d
    from cmdlmc_b200.topology import (DistanceInterpolator, InterpolatedTransformation,
                                      ReLUTransformation)
    relu = ReLUTransformation(a=0.9, b=2.35, d0=2.5, left_bound=2.2, right_bound=3.2)
    d = np.array([2.0, 2.2, 2.3, 2.5, 2.8, 3.2, 3.5])
    np.testing.assert_array_equal(relu(d), [2.0, 2.2, 2.35, 0.9 * 0.0 + 2.35, 0.9 * (2.8 - 2.5) + 2.35, 3.2, 3.5])
    xs = np.linspace(2.0, 3.4, 15)
    ys = 2.3 + 0.8 * (xs - 2.0) ** 1.5
    it = InterpolatedTransformation(xs, ys)
    q = np.array([1.5, 2.0, 2.05, 2.77, 3.4, 3.9])
    ref = interp1d(xs, ys, kind="linear")
    want = q.copy()
    inside = (xs[0] <= q) & (q <= xs[-1])
    want[inside] = ref(q[inside])
    want[want < xs[0]] = ys[0]
    np.testing.assert_array_equal(it(q), want)
    ip = DistanceInterpolator(4.0)
    out = ip(np.array([0.0, 2.0, np.inf]), np.full((3, 2), 3.0), np.full((3, 2), 2.0))
    np.testing.assert_array_equal(out, [[3.0, 3.0], [2.5, 2.5], [2.0, 2.0]])


def test_diffusion_coefficient_from_rows():
    """average_MC_out.get_slope equivalent: D = slope / 6 of the summed MSD per reset interval."""
    import numpy as np
    from cmdlmc_b200.output import diffusion_coefficient
    rng = np.random.RandomState(1)
    rows = []
    for frame in range(10, 3001, 10):
        t = frame * 0.5
        t_in = (frame % 1000) * 0.5
        msd = np.array([0.2, 0.3, 0.1]) * t_in + rng.normal(scale=1e-3, size=3)
        rows.append([frame, t, *msd, 40])
    r = diffusion_coefficient(rows, reset_frequency=1000)
    assert r["intervals"] == 4 or r["intervals"] == 3
    assert abs(r["slope"] - 0.6) < 5e-3 and abs(r["diffusion_coefficient"] - 0.1) < 1e-3


def test_unknown_jump_rate_is_rejected_with_a_clear_error():
    """The reference takes any callable as JumpRate; the device evaluates the rate inside the list
    kernels, so anything it does not know is refused before a handle is created."""
    from cmdlmc_b200.topology import DeviceTopology
    with pytest.raises(TypeError, match="cannot be evaluated on the device"):
        DeviceTopology(None, 4, 3.0, 2.0, 0, jumprate=lambda d: d)
