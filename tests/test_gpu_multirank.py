"""Real multi-GPU run (NCCL, one process per GPU under torchrun) of the two sharding rows of
SURVEY.md 8(e); skipped on boxes with one GPU.  The reduced statistics must equal a one-process run."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from cmdlmc_b200 import synth

pytestmark = pytest.mark.gpu


def test_two_ranks_nccl_equal_one_process(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _two_ranks_equal_one_process(tmp_path, "nccl", 29533)


def test_two_ranks_on_one_gpu_equal_one_process(tmp_path):
    """The same comparison on a one-GPU box: two processes share cuda:0, torch.distributed on
    gloo (NCCL refuses two ranks on one device), the statistics reduce through
    parallel.allreduce_sum's torch path and the frame blocks through the coordinate skip pass."""
    _two_ranks_equal_one_process(tmp_path, "gloo", 29537)


def _two_ranks_equal_one_process(tmp_path, backend, port):
    import cmdlmc_b200 as cm
    from cmdlmc_b200.ensemble import run_kmc_ensemble
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET, build_with_retry
    here = os.path.dirname(os.path.abspath(__file__))
    out = tmp_path / "multi.json"
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        str(port), os.path.join(here, "multirank_worker.py"), str(out), backend],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    got = json.load(open(out))
    assert got["world"] == 2
    if backend == "nccl":      # the statistics travelled through the library's own communicator
        assert got["comm_world"] == 2 and got["nccl_version"] > 20000
    w = synth.workload("C4")
    nfr, R = 160, 12
    frames = synth.trajectory(w, nfr)
    box = cm.AtomBoxMonoclinic(w.cell)
    rate = cm.Fermi(*w.rate_params)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                       MODE_VERLET, rate, cap), frames)
    np.testing.assert_array_equal(got["hist"], topo.distance_histogram(0.0, 5.0, 50))
    cnt, reb, rs = topo.frame_info()
    assert got["pairs"] == cnt.sum() and got["rebuilds"] == reb.sum()
    np.testing.assert_allclose(got["rate_sum"], rs.sum(), rtol=1e-12)
    one = run_kmc_ensemble(box, lambda a, b: frames[a:b], nfr, n_sites=w.n_oxygen,
                           n_protons=w.n_protons, cutoff=w.cutoff, buffer=w.buffer, jumprate=rate,
                           time_step=w.time_step, n_replicas=R, seed=21, reset_frequency=80,
                           print_frequency=20, chunk=64, histogram=(0.0, 5.0, 50), rank=0, world=1)
    assert got["events"] == one["events"] and got["n_replicas"] == R
    np.testing.assert_array_equal(got["jump_hist"], one["jump_hist"])
    np.testing.assert_array_equal(got["occupancy"], one["occupancy_counts"])
    np.testing.assert_allclose(got["msd_mean"], one["observables"]["mean"], rtol=1e-12)
    np.testing.assert_allclose(got["msd_sem"], one["observables"]["sem"], rtol=1e-9)
