#!/usr/bin/env python
"""bench.py -- throughput of the cMD/LMC per-frame hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the CPU path (reference arm)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W      # one rank per GPU

Workload (config.workload): BASELINE.json configs[1] -- "C2": CsH2PO4-like monoclinic cell, 400 O,
Fermi jump rate, synthetic trajectory (cmdlmc_b200/synth.py, seed 1).  One STEP = one pass of the
geometry -> neighbour list -> jump rate stages (SURVEY.md 8(d) metric M1) over one block of
`frames_per_step` frames per GPU, every unordered O-O pair of every frame evaluated
(NeighborTopology.topology_bruteforce_generator semantics, topology.py:55-78), the per-frame
lists (start, dest, dist, omega) materialised in HBM in the reference's order, followed -- when
more than one GPU runs -- by one NCCL all-reduce of the block statistics.  Frame blocks are
independent, so GPUs take disjoint blocks (weak scaling, no data-path collective).

    value    frames x O-pairs / s, whole job, inputs resident in HBM
    e2e      the same through the host-pointer C-ABI calls (cmd_topo_build + cmd_topo_frame_info):
             pinned-host -> device copy of the block and device -> host read of the per-frame
             results inside the timed region
    roofline the dense pair kernel (k_pairs_dense) against the measured FP64 pipe peak
    m2       KMC site-updates / s (SURVEY.md 8(d) metric M2): Verlet-mode topology of a sub-block
             + Philox-mode KMC of `--replicas` replicas, timed on its own
    cpu_baseline  the oracle's C port of the same step (OpenMP, all host cores) on a bounded
             sample, plus the reference's own Cython AtomBox on a few frames when oracle/_ref
             is present
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "frames*O-pairs/s"
UNIT = "pairs/s"
# algorithmic FP64 cost per unordered O-O pair
FLOP_ORTHO = 20
FLOP_GENERAL_REFERENCE = 279      # the reference's 27-image algorithm
# issue slots of the dense kernel's filter per unordered pair: 3 IADD + 3 I2FP +
# 3 (ortho) / 6 (triangular general cell) multiply-adds + 3 norm + 1 compare + 1 mask
FILTER_SLOTS_ORTHO = 14
FILTER_SLOTS_GENERAL = 17


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--frames-per-step", type=int, default=16384)
    ap.add_argument("--replicas", type=int, default=1024)
    ap.add_argument("--kmc-frames", type=int, default=2048)
    ap.add_argument("--cpu-seconds", type=float, default=12.0,
                    help="target CPU time of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-m2", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--e2e-frames", type=int, default=4096,
                    help="frames of the M2 end-to-end run through the Python API")
    return ap.parse_args()


def pairs_per_frame(n):
    return n * (n - 1) // 2


# ------------------------------------------------------------------------------ clocks --------
class ClockSampler:
    """nvidia-smi sampled every 50 ms while the timed regions run (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU legs ------
def cpu_port_rate(w, seconds, threads=None):
    """The oracle's C port of the step (all-pairs topology + rates per frame, OpenMP over frames)
    on a bounded sample of the same workload.  Returns (pairs/s, cores, sample text)."""
    from oracle import oracle as orc
    from cmdlmc_b200 import synth
    cores = threads or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    orc.build()
    box = orc.OracleBox(w.cell)
    rc = w.cutoff + w.buffer
    ppf = pairs_per_frame(w.n_oxygen)
    probe = max(cores, 8)
    fr = synth.trajectory(w, probe)
    orc.bench_frames(box, fr, rc, w.rate_kind, w.rate_params)   # warm the thread pool
    t = time.perf_counter()
    orc.bench_frames(box, fr, rc, w.rate_kind, w.rate_params)
    dt = time.perf_counter() - t
    nfr = int(max(probe, min(16384, seconds / max(dt / probe, 1e-9))))
    nfr = (nfr + cores - 1) // cores * cores
    fr = synth.trajectory(w, nfr)
    best = None
    for _ in range(2):
        t = time.perf_counter()
        orc.bench_frames(box, fr, rc, w.rate_kind, w.rate_params)
        dt = time.perf_counter() - t
        best = dt if best is None or dt < best else best
    sample = "%d frames of %s (%d O, %d pairs/frame), best of 2, gcc -O2 -fopenmp port" % (
        nfr, w.name, w.n_oxygen, ppf)
    return nfr * ppf / best, cores, sample, nfr, best


def reference_compiled_available():
    try:
        from oracle import build_ref
        return build_ref.is_built()
    except Exception:
        return False


def reference_compiled_run(w, frames, procs, repeats=1):
    """The REFERENCE's own compiled AtomBox (oracle/_ref, built from the reference's Cython
    sources) on `frames` frames of the workload, `procs` worker processes: oracle/ref_bench.py in
    a child process (it forks workers; this process may hold a CUDA context).  Returns its dict."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), "--workload", w.name,
           "--frames", str(int(frames)), "--procs", str(int(procs)), "--repeats", str(int(repeats))]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1800)
    if out.returncode != 0:
        raise RuntimeError("oracle/ref_bench.py failed: " + out.stderr[-500:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def reference_compiled_rate(w, seconds):
    """Bounded sample sized from a one-frame-per-process probe.  (pairs/s, cores, sample text)."""
    procs = os.cpu_count() or 1
    probe = reference_compiled_run(w, procs, procs)
    per_round = max(probe["seconds"], 1e-3)                 # one frame per process
    k = int(max(1, min(64, seconds / per_round)))
    r = reference_compiled_run(w, procs * k, procs)
    sample = ("%d frames of %s (%d O) through the reference's compiled AtomBox.length_all_to_all "
              "(oracle/_ref, Cython -O3 -ffast-math) + NumPy cutoff and Fermi, %d processes"
              % (r["frames"], w.name, w.n_oxygen, r["procs"]))
    return r["pairs_per_s"], r["procs"], sample, r


def cpu_baseline_block(w, seconds):
    """cpu_baseline object: the reference's compiled code when oracle/_ref travelled with the
    snapshot (kind "reference"), else the oracle's C port (kind "port"); the other one rides
    along as extra keys."""
    port_v, port_cores, port_sample, _, _ = cpu_port_rate(w, min(seconds, 8.0))
    one_core = {}
    try:   # the reference is single-threaded: its one-core figures beside the all-cores ones
        # (a fresh process: libgomp reads OMP_NUM_THREADS once)
        code = ("import sys, json; sys.path.insert(0, %r); import bench; from cmdlmc_b200 import synth; "
                "print(json.dumps(bench.cpu_port_rate(synth.workload(%r), 2.0, threads=1)[0]))" % (ROOT, w.name))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, OMP_NUM_THREADS="1"))
        one_core["port_pairs_per_s"] = float(out.stdout.strip().splitlines()[-1])
    except Exception as e:
        one_core["port_error"] = str(e)
    if reference_compiled_available():
        try:
            v, cores, sample, _ = reference_compiled_rate(w, seconds)
            try:
                one_core["reference_pairs_per_s"] = reference_compiled_run(w, 4, 1)["pairs_per_s"]
            except Exception as e:
                one_core["reference_error"] = str(e)
            return {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                    "one_core": one_core,
                    "port_pairs_per_s": port_v, "port_cores": port_cores, "port_sample": port_sample}
        except Exception as e:   # the checker must never take the bench down
            note = "oracle/_ref failed: %s" % e
    else:
        note = "oracle/_ref not present"
    return {"value": port_v, "unit": UNIT, "cores": port_cores, "kind": "port",
            "sample": port_sample, "note": note, "one_core": one_core}


def run_reference(args):
    """Reference arm: the CPU implementation of the same step on the host cores -- the reference's
    own compiled AtomBox (oracle/_ref) when it is present, else the oracle's C port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cmdlmc_b200 import synth
    w = synth.workload(args.workload)
    ppf = pairs_per_frame(w.n_oxygen)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # bounded sample per step so that the whole run ends within a few minutes
    budget = min(args.cpu_seconds, 120.0 / (steps + warm))
    kind = "port"
    if reference_compiled_available():
        try:
            procs = os.cpu_count() or 1
            probe = reference_compiled_run(w, procs, procs)
            k = int(max(1, min(64, budget / max(probe["seconds"], 1e-3))))
            r = reference_compiled_run(w, procs * k, procs, repeats=steps + warm)
            times = r["seconds_all"][warm:]
            nfr, cores, dt = r["frames"], r["procs"], float(sum(times))
            sample = ("%d frames of %s per step through the reference's compiled "
                      "AtomBox.length_all_to_all (oracle/_ref) + NumPy cutoff and Fermi, "
                      "%d processes" % (nfr, w.name, cores))
            kind = "reference"
        except Exception as e:
            print("bench.py: oracle/_ref arm failed (%s); using the C port" % e, file=sys.stderr)
    if kind == "port":
        _, cores, sample, nfr, _ = cpu_port_rate(w, budget)
        from oracle import oracle as orc
        box = orc.OracleBox(w.cell)
        fr = synth.trajectory(w, nfr)
        rc = w.cutoff + w.buffer
        for _ in range(warm):
            orc.bench_frames(box, fr, rc, w.rate_kind, w.rate_params)
        t = time.perf_counter()
        for _ in range(steps):
            orc.bench_frames(box, fr, rc, w.rate_kind, w.rate_params)
        dt = time.perf_counter() - t
    value = steps * nfr * ppf / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": 0,
        "config": workload_config(w, nfr, "bruteforce", extra={"l2": "n/a (CPU)"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(w, frames_per_step, mode, extra=None):
    cfg = {"workload": "%s: %s cell, %d O, %s jump rate, synthetic trajectory (seed %d)" % (
        w.name, "orthorhombic" if w.is_ortho else "monoclinic/triclinic", w.n_oxygen,
        w.rate_kind, w.seed),
        "frames_per_step_per_gpu": frames_per_step,
        "o_pairs_per_frame": pairs_per_frame(w.n_oxygen),
        "cutoff_plus_buffer": w.cutoff + w.buffer, "topology_mode": mode,
        "parallelism": "frame-block per GPU"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------ GPU arm -------
def _workload_objects(w):
    from cmdlmc_b200 import AtomBoxCubic, AtomBoxMonoclinic, Fermi, ActivationEnergy
    cell = np.asarray(w.cell, dtype=float)
    box = AtomBoxCubic(cell) if cell.size == 3 else AtomBoxMonoclinic(cell)
    rate = Fermi(*w.rate_params) if w.rate_kind == "Fermi" else ActivationEnergy(*w.rate_params)
    return box, rate


class Ranks:
    """The few collectives the bench itself needs (max of times, sum of work counts over ranks) and
    the library communicator the statistics travel over (cmd_comm_init, NCCL bound by the library)."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    def timed(self, fn, reps=1):
        """Best of `reps`: CUDA events on the launching stream, barrier + synchronize on both
        sides, max over ranks."""
        torch = self.torch
        best = None
        for _ in range(reps):
            self.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            self.barrier()
            ms = a.elapsed_time(b)
            best = ms if best is None or ms < best else best
        return self.max([best])[0]


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from cmdlmc_b200 import runtime, synth, parallel
    from cmdlmc_b200 import _abi
    from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's log (NCCL_DEBUG=INFO / VERSION, when the caller asks for it) goes to stderr so that
        # stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    runtime.init(local)
    runtime.use_torch_stream()
    dev = torch.device("cuda", local)
    numa = runtime.numa_topology(local) if hasattr(runtime, "numa_topology") else {}
    numa_node = runtime.bind_to_gpu_numa_node(local) if world > 1 else None
    comm = parallel.comm_init()          # the library's own NCCL communicator (statistics)
    R = Ranks(torch, dist, world, dev)
    lib = _abi.lib()

    w = synth.workload(args.workload)
    n, B = w.n_oxygen, args.frames_per_step
    ppf = pairs_per_frame(n)
    box, rate = _workload_objects(w)
    kind = 0 if w.is_ortho else 1

    # this rank's frame block of the synthetic trajectory, in pinned host memory and in HBM
    host = torch.empty((B, n, 3), dtype=torch.float64, pin_memory=True)
    host.numpy()[...] = synth.trajectory(w, B, start=rank * B)
    d_frames = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    in_bytes = B * n * 24

    topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, 0)
    # block statistics (listed directed pairs, sum of the listed rates): accumulated on the device
    # by the library after every step, summed over the ranks ONCE per timed region
    stats = torch.zeros(2, dtype=torch.float64, device=dev)

    def step_resident():
        topo.build_dev(d_frames.data_ptr(), B)
        _abi.check(lib.cmd_topo_block_stats_dev(topo.handle, C.c_void_p(stats.data_ptr())))

    def reduce_stats():
        _abi.check(lib.cmd_stats_allreduce_dev(C.c_void_p(stats.data_ptr()), 2, None, 0))

    W = max(3, args.warmup)
    for _ in range(W):
        step_resident()
    reduce_stats()
    R.barrier()
    counts, _, rate_sum = topo.frame_info()
    assert (counts >= 0).all()

    sampler = ClockSampler(local)
    sampler.start()
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(K)]
    stats.zero_()
    l0 = runtime.launch_count()
    R.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        ev[k][0].record()
        topo.build_dev(d_frames.data_ptr(), B)
        ev[k][1].record()
        _abi.check(lib.cmd_topo_block_stats_dev(topo.handle, C.c_void_p(stats.data_ptr())))
    reduce_stats()                       # the path's one collective: a 16-byte sum over the ranks
    e1.record()
    R.barrier()
    launches = runtime.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    got = stats.tolist()
    want_pairs = R.sum([float(counts.sum()) * K])[0]
    assert abs(got[0] - want_pairs) < 0.5, "reduced pair count %r != %r" % (got[0], want_pairs)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region -------------
    # Headline: float32 host frames -- the layout the reference holds trajectories in
    # (IO/trajectory_parser.py:324, HDF5 storage; up-cast to f64 before any arithmetic, here on the
    # device).  The same call on float64 host frames rides along.
    h_counts = np.zeros(B, np.int64)
    h_rebuilt = np.zeros(B, np.uint8)
    h_rsum = np.zeros(B)
    host32 = torch.empty(host.shape, dtype=torch.float32, pin_memory=True)
    host32.copy_(host)

    def e2e_leg(ptr, itemsize):
        def step():
            _abi.check(lib.cmd_topo_build(topo.handle, C.c_void_p(ptr), itemsize, B))
            _abi.check(lib.cmd_topo_frame_info(topo.handle, _abi.ptr(h_counts, C.c_int64),
                                               _abi.ptr(h_rebuilt, C.c_uint8), _abi.ptr(h_rsum)))
            _abi.check(lib.cmd_topo_block_stats_dev(topo.handle, C.c_void_p(stats.data_ptr())))
        for _ in range(2):
            step()
        R.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        reduce_stats()
        f1.record()
        R.barrier()
        return max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)

    e2e32_ms = e2e_leg(host32.data_ptr(), 4)
    e2e64_ms = e2e_leg(host.data_ptr(), 8)

    # The streaming form of the same calls (cmd_topo_build_async on two topologies used
    # alternately, two page-locked host blocks): the upload of step k+1 overlaps the kernels of
    # step k, the result of a step (cmd_topo_frame_info) is read while the next one is in flight.
    # Every step still copies its own input block to the GPU and reads its own result back.
    topo_b = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, topo.stride)
    host32_b = torch.empty(host.shape, dtype=torch.float32, pin_memory=True)
    host32_b.copy_(host32)
    pair = ((topo, host32), (topo_b, host32_b))

    def stream_leg(steps):
        def finish(t):
            _abi.check(lib.cmd_topo_wait(t.handle))
            _abi.check(lib.cmd_topo_frame_info(t.handle, _abi.ptr(h_counts, C.c_int64),
                                               _abi.ptr(h_rebuilt, C.c_uint8), _abi.ptr(h_rsum)))
        for k in range(steps):
            t, h = pair[k % 2]
            _abi.check(lib.cmd_topo_build_async(t.handle, C.c_void_p(h.data_ptr()), 4, B))
            _abi.check(lib.cmd_topo_block_stats_dev(t.handle, C.c_void_p(stats.data_ptr())))
            if k >= 1:
                finish(pair[(k - 1) % 2][0])
        finish(pair[(steps - 1) % 2][0])

    _abi.check(lib.cmd_topo_build(topo_b.handle, C.c_void_p(host32_b.data_ptr()), 4, B))
    stream_leg(2)
    R.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    t0 = time.perf_counter()
    stream_leg(K)
    reduce_stats()
    f1.record()
    R.barrier()
    e2e_stream_ms = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
    del topo_b, host32_b, pair
    # a caller that hands over plain (pageable) NumPy memory: the library's page-locked staging ring
    pageable32 = host32.numpy().copy()
    st0 = runtime.staging_stats()
    e2e_page_ms = e2e_leg(pageable32.ctypes.data, 4)
    st1 = runtime.staging_stats()
    del pageable32

    # the hardware ceiling of that step: the bare pinned host -> device copies, all ranks at once
    def h2d_ceiling(src):
        dst = torch.empty_like(src, device=dev)
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        R.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            dst.copy_(src, non_blocking=True)
        b.record()
        R.barrier()
        return a.elapsed_time(b)

    copy32_ms = h2d_ceiling(host32)
    copy64_ms = h2d_ceiling(host)
    clocks = sampler.stop()
    del host32

    # max over ranks
    ms_total, e2e32_ms, e2e64_ms, kernel_ms, copy32_ms, copy64_ms, e2e_page_ms, e2e_stream_ms = R.max(
        [ms_total, e2e32_ms, e2e64_ms, kernel_ms, copy32_ms, copy64_ms, e2e_page_ms, e2e_stream_ms])
    value = world * K * B * ppf / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (k_pairs_dense) --------------------------------------
    # HBM: algorithmic bytes per launch = frames read once (24 B per atom-frame) + every listed
    # directed pair written once (start i32, dest i32, dist f64, omega f64 = 24 B), DESIGN.md 3.
    peak_tf = runtime.fp64_peak_tflops(40000)
    out_bytes = float(counts.sum()) * 24.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_gbs = (in_bytes + out_bytes) / (kernel_ms * 1e-3) / 1e9
    traffic = None
    traffic_src = None
    try:   # dram bytes per launch of the same kernel from the committed ncu --set full capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "r2_dense_traffic.json")))
        traffic = prof["dram_bytes_per_frame"] * B
        traffic_src = prof.get("source")
    except Exception:
        pass
    roofline = {
        "kernel": "k_pairs_dense", "bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak,
        "unit": "GB/s", "frac": hbm_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": in_bytes + out_bytes,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else
                       "fallback of B200_PROFILING.md",
        "kernel_ms_per_launch": kernel_ms,
        "note": "the kernel is bound by instruction issue (ncu: issue slots 56 %, FP64 pipe 16 %, "
                "DRAM 21 % of peak; profiles/r2p_ncu_dense_summary.txt), not by HBM",
        "fp64_peak_tflops_measured": peak_tf,
        "reference_equivalent_tflops": B * ppf * (FLOP_ORTHO if kind == 0 else
                                                  FLOP_GENERAL_REFERENCE) / (kernel_ms * 1e-3) / 1e12,
    }

    def e2e_block(ms, h2d, copy_ms, api_note):
        return {"value": world * K * B * ppf / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": B * 13 + 4, "ms_per_step": ms / K,
                "h2d_only_ms_per_step": copy_ms / K,
                "h2d_gbs_per_rank": h2d / (copy_ms / K * 1e-3) / 1e9, "api": api_note}

    e2e = e2e_block(e2e_stream_ms, in_bytes // 2, copy32_ms,
                    "cmd_topo_build_async on two topologies used alternately (host float32 frames: the "
                    "reference's trajectory storage, trajectory_parser.py:324) + cmd_topo_wait + "
                    "cmd_topo_frame_info: the upload of step k+1 overlaps the kernels of step k")
    e2e["one_block_at_a_time"] = e2e_block(e2e32_ms, in_bytes // 2, copy32_ms,
                                           "cmd_topo_build (blocking) + cmd_topo_frame_info on the same "
                                           "float32 host frames")
    e2e["f64_frames"] = e2e_block(e2e64_ms, in_bytes, copy64_ms,
                                  "the same calls on float64 host frames")
    e2e["pageable_f32_frames"] = {
        "value": world * K * B * ppf / (e2e_page_ms * 1e-3), "unit": UNIT,
        "ms_per_step": e2e_page_ms / K, "h2d_bytes_per_step": in_bytes // 2,
        "host_gbs_per_rank": in_bytes / 2 / (e2e_page_ms / K * 1e-3) / 1e9,
        "ring_bytes_staged": int(st1[0] - st0[0]), "ring_host_threads": int(st1[2]),
        "api": "the same calls on a plain NumPy array: cmd_topo_build stages it through the library's "
               "page-locked ring (csrc/staging.cu)"}
    e2e["numa"] = dict(numa, bound_node=numa_node)
    e2e["note"] = ("h2d_only_ms_per_step = the bare pinned host->device copies of the same bytes on all "
                   "ranks at once: the hardware ceiling of the step on this host")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, B, "bruteforce", extra={
            "l2": "inputs (%.0f MB) and outputs (%.0f MB) per step exceed the 126 MB L2" % (
                in_bytes / 1e6, out_bytes / 1e6),
            "directed_pairs_per_frame_mean": float(counts.mean())}),
        "gpu_launches": int(launches),
        "e2e": e2e, "roofline": roofline, "clocks": clocks,
        "collective": {"what": "sum over ranks of (listed pairs, listed rate sum) accumulated on the "
                               "device over the timed region: one 16-byte all-reduce per region",
                       "api": "cmd_stats_allreduce_dev", "nccl_version": comm.get("nccl_version"),
                       "comm_world": comm.get("world"), "listed_pairs_all_ranks": got[0],
                       "rate_sum_all_ranks": got[1]},
    }

    # ---- M2: KMC / LMC site-updates per second (own timed regions) ----------------------------
    m2 = None
    if not args.no_m2:
        m2 = run_m2(args, w, box, rate, d_frames, R, rank)
        line["m2"] = m2
    # ---- the other BASELINE.json configs, short legs ------------------------------------------
    cfgs = None
    if not args.no_configs:
        cfgs = run_config_legs(args, R, rank, world)
        line["configs"] = cfgs

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block(w, args.cpu_seconds)
        if m2 is not None:
            m2["cpu_baseline"] = m2_cpu_baseline(w, m2)
    # the numbers the metric names, last on the line
    line["headline"] = headline(line, m2, cfgs)
    if rank == 0:
        print(json.dumps(line))
    parallel.comm_destroy()
    if world > 1:
        dist.destroy_process_group()


def headline(line, m2, cfgs):
    h = {"m1_frames_x_o_pairs_per_s": line["value"], "m1_e2e_pairs_per_s": line["e2e"]["value"],
         "m1_roofline_frac": line["roofline"]["frac"]}
    if m2:
        h.update(m2_kmc_site_updates_per_s=m2["value"],
                 m2_e2e_site_updates_per_s=m2.get("e2e", {}).get("value"),
                 m2_lmc_attempts_per_s=m2["lmc_sweep"]["value"],
                 verlet_pair_frames_per_s=m2["verlet_pipeline"]["listed_pair_frames_per_s"],
                 single_replica_frames_per_s=m2["single_replica_replay"]["frames_per_s"])
        if "cpu_baseline" in m2:
            h["m2_cpu_site_updates_per_s"] = m2["cpu_baseline"]["value"]
    if "cpu_baseline" in line:
        h["m1_cpu_pairs_per_s"] = line["cpu_baseline"]["value"]
    if cfgs:
        for k, v in cfgs.items():
            if isinstance(v, dict) and "headline" in v:
                h[k] = v["headline"]
    return h


def run_m2(args, w, box, rate, d_frames, R, rank):
    """Verlet-mode topology of a sub-block + Philox KMC of R replicas per GPU (replica-sharded:
    every GPU walks its own replicas over its own frames; statistics all-reduced)."""
    import torch
    from cmdlmc_b200 import synth
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX, RNG_REPLAY
    from cmdlmc_b200.lmc import DeviceLMC
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET
    world = R.world
    F = min(args.kmc_frames, d_frames.shape[0])
    NR = args.replicas
    n = w.n_oxygen
    topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, 0)
    topo.build_dev(d_frames.data_ptr(), F)
    counts, rebuilt, _ = topo.frame_info()
    # the Verlet pipeline itself (k_dr, rebuild schedule, rebuilds, k_refresh) on the whole block the
    # M1 step used: a fresh object with a known capacity so that every run starts from frame 0
    Fv = d_frames.shape[0]
    tv = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, topo.stride)
    tv.build_dev(d_frames.data_ptr(), Fv)
    vms = R.timed(lambda: tv.build_dev(d_frames.data_ptr(), Fv), reps=3)
    vcounts, vreb, _ = tv.frame_info()
    vsum = R.sum([float(vcounts.sum())])[0]
    verlet = {"frames_per_gpu": int(Fv), "ms": vms, "rebuilds": int(vreb.sum()),
              "listed_pair_frames_per_s": vsum / vms * 1e3,
              "frames_x_o_pairs_equiv_per_s": world * Fv * pairs_per_frame(n) / vms * 1e3,
              "list_traffic_gbs_per_gpu": float(vcounts.sum()) * 32.0 / vms / 1e6,
              "note": "continuation blocks of one trajectory (state carried); 8 B of indices read + "
                      "24 B written per listed pair-frame; k_refresh is FP64-bound"}
    del tv
    lattices = np.stack([synth.initial_lattice(n, w.n_protons, 4000 + r)[0] for r in range(NR)])
    times, updates, events = [], 0, 0
    for it in range(4):
        kmc = DeviceKMC(box, lattices, w.time_step, RNG_PHILOX, seed=11 + it)
        ms = R.timed(lambda: kmc.advance(topo))
        st = kmc.state()
        if it > 0:
            times.append(ms)
            updates, events = int(st["site_updates"].sum()), int(st["n_events"].sum())
    ms = float(np.mean(times))
    updates, events = R.sum([updates, events])
    rate_su = updates / (ms * 1e-3)
    # legacy LMC sweep (row A14): one sweep = P attempts per frame and replica, Philox
    lmc_times, lmc_attempts, lmc_jumps = [], 0, 0
    for it in range(3):
        lmc = DeviceLMC(lattices, RNG_PHILOX, seed=21 + it)
        t = R.timed(lambda: lmc.advance(topo, w.time_step, 1))
        if it > 0:
            lmc_times.append(t)
            ls = lmc.state()
            lmc_attempts, lmc_jumps = int(ls["attempts"].sum()), int(ls["jumps"].sum())
    lmc_ms = float(np.mean(lmc_times))
    lmc_attempts, lmc_jumps = R.sum([lmc_attempts, lmc_jumps])
    lmc_block = {"metric": "LMC site-updates/s (jump attempts)", "kernel": "k_lmc_sweep",
                 "value": lmc_attempts / (lmc_ms * 1e-3), "unit": "attempts/s", "ms": lmc_ms,
                 "jumps": lmc_jumps, "sweeps_per_frame": 1, "rng": "philox4x32-10",
                 "parity": "unpinned upstream (engine not in the reference tree)"}
    # one replica in exact-replay mode: what a reference `mdmc` run is (k_kmc_solo, one CTA)
    Fs = min(F, 2048)
    u = np.random.RandomState(5).random_sample((1, 32 * Fs + 1000))
    ts = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, topo.stride)
    ts.build_dev(d_frames.data_ptr(), Fs)
    solo_ms = []
    solo_events = 0
    for it in range(2):
        one = DeviceKMC(box, lattices[:1], w.time_step, RNG_REPLAY)
        one.set_replay_stream(u)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one.advance(ts)
        b.record()
        torch.cuda.synchronize()
        solo_ms.append(a.elapsed_time(b))
        solo_events = int(one.state()["n_events"].sum())
    del ts
    single = {"kernel": "k_kmc_solo", "rng": "replay (reference np.random protocol, bit-exact mode)",
              "frames": int(Fs), "ms": float(min(solo_ms)), "events": solo_events,
              "frames_per_s": Fs / (min(solo_ms) * 1e-3),
              "site_updates_per_s": float(counts[:Fs].sum()) / (min(solo_ms) * 1e-3)}
    from cmdlmc_b200 import runtime as _rt
    smem_peak = _rt.smem_peak_gbs()
    out = {"metric": "KMC site-updates/s", "value": rate_su, "unit": "site-updates/s",
           "replicas_per_gpu": NR, "frames": F, "ms": ms, "events": events,
           "rng": "philox4x32-10", "directed_pairs_per_frame_mean": float(counts.mean()),
           "verlet_rebuilds": int(rebuilt.sum()), "kernel": "k_kmc_stream",
           "roofline": {"bound": "smem", "unit": "GB/s", "achieved": rate_su / world * 16 / 1e9,
                        "peak": smem_peak, "frac": rate_su / world * 16 / 1e9 / smem_peak,
                        "peak_source": "measured: cmd_smem_peak (conflict-free 16-byte loads on every SM)",
                        "nominal_peak": 148 * 128 * 1.965,
                        "note": "per GPU; 16 B of (start, dest, omega) read from the shared-memory ring "
                                "per site-update; the kernel is bound by per-warp latency and event "
                                "handling, not by this rate (DESIGN.md 6.1)"},
           "verlet_pipeline": verlet, "lmc_sweep": lmc_block, "single_replica_replay": single}
    out["e2e"] = m2_e2e(args, w, box, rate, float(counts.mean()), R)
    return out


def m2_e2e(args, w, box, rate, pairs_per_frame_mean, R):
    """M2 end to end through the reference-facing Python API, the way `mdmc` drives it
    (mdlmc/main.py:73-158): host frames -> ArrayTrajectory -> NeighborTopology (Verlet) -> Fermi ->
    KMCLattice -> ObservablesOutput rows back on the host.  One replica per GPU (a KMC replica is
    sequential in time), exact-replay mode (the reference's np.random protocol) and Philox mode."""
    import gc
    import cmdlmc_b200 as cm
    from cmdlmc_b200 import synth
    from cmdlmc_b200.kmc import KMCLattice, ObservablesOutput
    from cmdlmc_b200.topology import NeighborTopology
    from cmdlmc_b200.trajectory import ArrayTrajectory
    nfr = int(args.e2e_frames)
    frames = synth.trajectory(w, nfr).astype(np.float32)       # the reference's storage dtype
    names = np.array(["O"] * w.n_oxygen)
    res = {}
    for mode in ("replay", "philox"):
        best, rows_n, ev_n = None, 0, 0
        for it in range(2):
            np.random.seed(3)
            R.barrier()
            t0 = time.perf_counter()
            top = NeighborTopology(ArrayTrajectory(frames, names, time_step=w.time_step), box,
                                   donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer)
            kmc = KMCLattice(top, atom_box=box, jumprate_function=rate, lattice_size=w.n_oxygen,
                             proton_number=w.n_protons, donor_atoms="O", time_step=w.time_step,
                             rng=mode, chunk_size=4096)
            rows = list(ObservablesOutput(kmc, 1000, 100))
            dt = time.perf_counter() - t0
            rows_n, ev_n = len(rows), len(kmc.event_log["frame"])
            del rows, kmc, top
            gc.collect()
            best = dt if best is None or dt < best else best
        best = R.max([best])[0]
        res[mode] = {"seconds": best, "frames_per_s": R.world * nfr / best, "rows": rows_n, "events": ev_n,
                     "site_updates_per_s": R.world * nfr * pairs_per_frame_mean / best}
    # the reference's own generator protocol (topology.py:80-114): one (start, destination, distance,
    # frame) tuple of HOST arrays per frame -- every list leaves the GPU again, the slowest way to
    # use the library and the one an unmodified reference driver takes
    gen_frames = min(nfr, 2048)
    R.barrier()
    t0 = time.perf_counter()
    top = NeighborTopology(ArrayTrajectory(frames[:gen_frames], names, time_step=w.time_step), box,
                           donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer)
    listed = 0
    for start, dest, dist, frame in top.topology_verlet_list_generator():
        listed += len(start)
    gen_s = R.max([time.perf_counter() - t0])[0]
    del top
    gc.collect()
    res["generator"] = {"frames": gen_frames, "seconds": gen_s, "frames_per_s": R.world * gen_frames / gen_s,
                        "listed_pair_frames_per_s": R.world * listed / gen_s,
                        "d2h_bytes": int(listed) * 16,
                        "api": "NeighborTopology.topology_verlet_list_generator(): host arrays per frame"}
    return {"value": res["philox"]["site_updates_per_s"], "unit": "site-updates/s",
            "frames": nfr, "replicas_per_gpu": 1, "reference_protocol_generator": res["generator"],
            "h2d_bytes_per_run": int(frames.nbytes), "d2h": "observable rows + event log",
            "api": "ArrayTrajectory -> NeighborTopology -> KMCLattice -> ObservablesOutput "
                   "(float32 host frames, wall clock incl. the Python layer, best of 2)",
            "philox": res["philox"], "replay_bit_exact_mode": res["replay"]}


def m2_cpu_baseline(w, m2):
    """The oracle's C port of one KMCLattice iteration per frame (Verlet refresh + rates + exact
    replay KMC) on the host cores: independent replicas side by side (oracle/kmc_bench.py)."""
    procs = os.cpu_count() or 1
    try:
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "kmc_bench.py"), "--workload", w.name,
               "--frames", "256", "--procs", str(procs)]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                             env=dict(os.environ, OMP_NUM_THREADS="1"))
        if out.returncode != 0:
            raise RuntimeError(out.stderr[-300:])
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"value": None, "unit": "site-updates/s", "kind": "port", "note": "failed: %s" % e}
    return {"value": r["site_updates_per_s"], "unit": "site-updates/s", "cores": r["procs"], "kind": "port",
            "sample": "%d independent replicas x %d frames of %s: Verlet refresh + rates + exact-replay "
                      "KMC per frame (MDMC.py:77-171), gcc -O2 port, one process per core"
                      % (r["procs"], r["frames"], w.name),
            "one_core_site_updates_per_s": r["one_core_site_updates_per_s"],
            "one_core_frames_per_s": r["frames_per_s_per_core"],
            "one_core_kmc_only_site_updates_per_s": r["one_core_kmc_only_site_updates_per_s"],
            "topology_share_of_cpu_time": r["topology_share"],
            "reference_python_frames_per_s": "675 (N=144) / 215 (N=400), SURVEY.md section 6 "
                                             "(needs /root/reference: not on the GPU box)"}


def run_config_legs(args, R, rank, world):
    """Short legs on the other BASELINE.json configs (C2 is the line itself): C1 small ortho box,
    C3 triclinic + activation-energy rate with the trajectory frame-block sharded over the ranks
    (Verlet mode, schedule from all-gathered step lengths), C4 1024 replicas replica-sharded incl.
    an exact-replay subset, C5 32k-O box through the cell list + the pair-distance histogram."""
    import ctypes as C
    import torch
    from cmdlmc_b200 import synth, parallel
    from cmdlmc_b200 import _abi
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX, RNG_REPLAY
    from cmdlmc_b200.lmc import DeviceLMC
    from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE, MODE_VERLET
    out = {}
    dev = R.dev

    def frames_dev(w, nfr, start):
        return torch.from_numpy(synth.trajectory(w, nfr, start=start)).to(dev)

    # ---- C1: reference integration size, 144 O, ortho ----------------------------------------
    w = synth.workload("C1")
    box, rate = _workload_objects(w)
    n, B = w.n_oxygen, 8192
    d = frames_dev(w, B, rank * B)
    t = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, 0)
    t.build_dev(d.data_ptr(), B)
    ms = R.timed(lambda: t.build_dev(d.data_ptr(), B), reps=3)
    tv = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, t.stride)
    tv.build_dev(d.data_ptr(), B)
    vms = R.timed(lambda: tv.build_dev(d.data_ptr(), B), reps=2)
    vc = tv.frame_info()[0]
    lat = np.stack([synth.initial_lattice(n, w.n_protons, 100 + r)[0] for r in range(256)])
    tk = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, t.stride)
    tk.build_dev(d.data_ptr(), 2048)
    kmc = DeviceKMC(box, lat, w.time_step, RNG_PHILOX, seed=5)
    kms = R.timed(lambda: kmc.advance(tk))
    su = R.sum([float(kmc.state()["site_updates"].sum())])[0]
    out["C1"] = {"n_oxygen": n, "frames_per_gpu": B, "bruteforce_ms": ms,
                 "frames_x_o_pairs_per_s": world * B * pairs_per_frame(n) / ms * 1e3,
                 "verlet_pair_frames_per_s": R.sum([float(vc.sum())])[0] / vms * 1e3,
                 "kmc_site_updates_per_s": su / kms * 1e3, "kmc_replicas_per_gpu": 256,
                 "headline": world * B * pairs_per_frame(n) / ms * 1e3}
    del t, tv, tk, kmc, d

    # ---- C3: triclinic, 2048 O, activation-energy rate, frame-block sharded Verlet run ---------
    w = synth.workload("C3")
    box, rate = _workload_objects(w)
    n = w.n_oxygen
    F3 = 512                                   # frames per GPU; the trajectory has world * F3 frames
    total = world * F3
    a3, b3 = parallel.frame_block(total, rank, world)
    lo3 = max(a3 - 1, 0)
    own3 = synth.trajectory(w, b3 - lo3, start=lo3).astype(np.float32)   # host frames of this rank (+ halo)

    def src(a, b):      # only the last rebuild frame before the block lies outside `own3`
        if a >= lo3 and b <= b3:
            return own3[a - lo3:b - lo3]
        return synth.trajectory(w, b - a, start=a)
    sh = parallel.ShardedTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, src, total,
                                  rank=rank, world=world, chunk=F3)
    R.barrier()
    t0 = time.perf_counter()
    pf = 0.0
    for first, tp in sh.blocks():
        pf += float(tp.frame_info()[0].sum())
    torch.cuda.synchronize()
    e2e_s = R.max([time.perf_counter() - t0])[0]
    # resident: this rank's block again, continuing blocks of one trajectory
    d = frames_dev(w, F3, rank * F3)
    vms = R.timed(lambda: sh.topo.build_dev(d.data_ptr(), F3), reps=2)
    vc = sh.topo.frame_info()[0]
    tb = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, 0)
    tb.build_dev(d.data_ptr(), 256)
    bms = R.timed(lambda: tb.build_dev(d.data_ptr(), 256), reps=2)
    pf_all = R.sum([pf])[0]
    out["C3"] = {"n_oxygen": n, "rate": w.rate_kind, "frames_total": total, "frames_per_gpu": F3,
                 "sharding": "frame block per GPU; rebuild schedule from the all-gathered step lengths "
                             "(cmd_allgather_dev), each rank uploads only its own block (float32 host "
                             "frames, wall clock incl. capacity probe and allocations)",
                 "verlet_e2e_seconds": e2e_s, "verlet_e2e_pair_frames_per_s": pf_all / e2e_s,
                 "verlet_resident_ms": vms,
                 "verlet_pair_frames_per_s": R.sum([float(vc.sum())])[0] / vms * 1e3,
                 "cell_list_us_per_frame": bms * 1e3 / 256,
                 "frames_x_o_pairs_equiv_per_s": world * 256 * pairs_per_frame(n) / bms * 1e3,
                 "headline": R.sum([float(vc.sum())])[0] / vms * 1e3}
    del sh, tb, d

    # ---- C4: 1024 replicas on a 384-O lattice, replica-sharded over the ranks ------------------
    w = synth.workload("C4")
    box, rate = _workload_objects(w)
    n = w.n_oxygen
    NR, F4 = 1024, 1024
    ids = parallel.replica_ids(NR, rank, world)
    d = frames_dev(w, F4, 0)                    # every rank walks the same trajectory
    tv = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, 0)
    tv.build_dev(d.data_ptr(), F4)
    cnt = tv.frame_info()[0]
    lat = np.stack([synth.initial_lattice(n, w.n_protons, 7000 + int(r))[0] for r in ids])
    best = None
    for it in range(2):
        kmc = DeviceKMC(box, lat, w.time_step, RNG_PHILOX, seed=31)
        kmc.set_replica_ids(rank, world)
        ms = R.timed(lambda: kmc.advance(tv))
        best = ms if best is None or ms < best else best
    su, evn = R.sum([float(kmc.state()["site_updates"].sum()), float(kmc.state()["n_events"].sum())])
    lmc = DeviceLMC(lat, RNG_PHILOX, seed=41)
    lms = R.timed(lambda: lmc.advance(tv, w.time_step, 1))
    att = R.sum([float(lmc.state()["attempts"].sum())])[0]
    # exact-replay subset: 64 replicas in total, the one-CTA-per-replica kernel against the
    # warp-per-replica kernel on the same uniform streams -- final lattices must be identical
    vids = parallel.replica_ids(64, rank, world)
    Fr = 256
    same = 1.0
    rms = 0.0
    if len(vids):
        tr = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, tv.stride)
        tr.build_dev(d.data_ptr(), Fr)
        vlat = np.stack([synth.initial_lattice(n, w.n_protons, 7000 + int(r))[0] for r in vids])
        u = np.stack([np.random.RandomState(1000 + int(r)).random_sample(64 * Fr + 1000) for r in vids])
        finals = []
        for solo in ("1", "0"):
            os.environ["CMDLMC_B200_KMC_SOLO"] = solo
            k = DeviceKMC(box, vlat, w.time_step, RNG_REPLAY)
            k.set_replay_stream(u)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); k.advance(tr); b.record()
            torch.cuda.synchronize()
            if solo == "1":
                rms = a.elapsed_time(b)
            finals.append(k.state()["lattices"].copy())
        os.environ.pop("CMDLMC_B200_KMC_SOLO", None)
        same = 1.0 if np.array_equal(finals[0], finals[1]) else 0.0
        del tr
    same_all = R.sum([same])[0]
    out["C4"] = {"n_oxygen": n, "replicas_total": NR, "frames": F4, "sharding": "replica r on rank r mod G",
                 "kmc_philox_ms": best, "kmc_site_updates_per_s": su / best * 1e3, "events": evn,
                 "lmc_attempts_per_s": att / lms * 1e3,
                 "replay_subset": {"replicas_total": 64, "frames": Fr, "ms": R.max([rms])[0],
                                   "two_kernels_bit_identical": bool(same_all == world),
                                   "note": "k_kmc_solo vs k_kmc_advance on the reference's RandomState "
                                           "streams; oracle parity of the same mode: tests/test_gpu_kmc.py"},
                 "headline": su / best * 1e3, "scaling": "strong (1024 replicas in total)"}
    del tv, kmc, lmc, d

    # ---- C5: 32768-O water-like box, cell list + pair-distance histogram, frames sharded -------
    w = synth.workload("C5")
    box, rate = _workload_objects(w)
    n, F5 = w.n_oxygen, 32    # BASELINE C5: 200 frames over the GPUs of a box (25 each at N = 8)
    d = frames_dev(w, F5, rank * F5)
    t = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, 0)
    t.build_dev(d.data_ptr(), F5)
    ms = R.timed(lambda: t.build_dev(d.data_ptr(), F5), reps=3)
    cnt = t.frame_info()[0]
    hist = np.zeros(500, np.int64)
    hms = R.timed(lambda: t.distance_histogram(0.0, 5.0, 500, out=hist))
    hist[:] = 0
    t.distance_histogram(0.0, 5.0, 500, out=hist)
    tot = parallel.allreduce_sum({"hist": hist})["hist"]
    pairs_all = R.sum([float(cnt.sum())])[0]
    out["C5"] = {"n_oxygen": n, "frames_per_gpu": F5, "path": "cell list",
                 "us_per_frame": ms * 1e3 / F5, "listed_pairs_per_frame": float(cnt.mean()),
                 "list_bytes_gbs_per_gpu": float(cnt.sum()) * 24 / ms / 1e6,
                 "frames_x_o_pairs_equiv_per_s": world * F5 * pairs_per_frame(n) / ms * 1e3,
                 "pair_histogram_ms": hms, "pair_histogram_total_all_ranks": int(tot.sum()),
                 "pair_histogram_consistent": bool(int(tot.sum()) == int(pairs_all)),
                 "headline": world * F5 * pairs_per_frame(n) / ms * 1e3}
    del t, d
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: whatever libraries print on file descriptor 1 while the
    # bench runs (NCCL's version / INFO lines when the caller sets NCCL_DEBUG) is sent to stderr,
    # where it stays readable, and the line itself goes to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
