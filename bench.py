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
    if reference_compiled_available():
        try:
            v, cores, sample, _ = reference_compiled_rate(w, seconds)
            return {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                    "port_pairs_per_s": port_v, "port_cores": port_cores, "port_sample": port_sample}
        except Exception as e:   # the checker must never take the bench down
            note = "oracle/_ref failed: %s" % e
    else:
        note = "oracle/_ref not present"
    return {"value": port_v, "unit": UNIT, "cores": port_cores, "kind": "port",
            "sample": port_sample, "note": note}


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
def run_b200(args):
    import torch
    import torch.distributed as dist
    from cmdlmc_b200 import AtomBoxCubic, AtomBoxMonoclinic, Fermi, ActivationEnergy, runtime, synth
    from cmdlmc_b200 import _abi
    from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE, MODE_VERLET

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
            os.environ["NCCL_DEBUG"] = "WARN"     # NCCL logs to stdout: keep it to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    runtime.init(local)
    runtime.use_torch_stream()
    dev = torch.device("cuda", local)
    numa_node = runtime.bind_to_gpu_numa_node(local) if world > 1 else None

    w = synth.workload(args.workload)
    n, B = w.n_oxygen, args.frames_per_step
    ppf = pairs_per_frame(n)
    cell = np.asarray(w.cell, dtype=float)
    box = AtomBoxCubic(cell) if cell.size == 3 else AtomBoxMonoclinic(cell)
    rate = Fermi(*w.rate_params) if w.rate_kind == "Fermi" else ActivationEnergy(*w.rate_params)

    # this rank's frame block of the synthetic trajectory, in pinned host memory and in HBM
    host = torch.empty((B, n, 3), dtype=torch.float64, pin_memory=True)
    host.numpy()[...] = synth.trajectory(w, B, start=rank * B)
    d_frames = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    in_bytes = B * n * 24

    topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, 0)
    stats = torch.zeros(2, dtype=torch.float64, device=dev)

    def step_resident():
        topo.build_dev(d_frames.data_ptr(), B)
        if world > 1:
            dist.all_reduce(stats)       # block statistics (pair count, rate sum): 16 bytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    counts, _, rate_sum = topo.frame_info()
    assert (counts >= 0).all()

    sampler = ClockSampler(local)
    sampler.start()
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(K)]
    l0 = runtime.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        ev[k][0].record()
        topo.build_dev(d_frames.data_ptr(), B)
        ev[k][1].record()
        if world > 1:
            dist.all_reduce(stats)
    e1.record()
    barrier()
    launches = runtime.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region -------------
    hptr = host.data_ptr()
    import ctypes as C
    lib = _abi.lib()
    h_counts = np.zeros(B, np.int64)
    h_rebuilt = np.zeros(B, np.uint8)
    h_rsum = np.zeros(B)

    def step_e2e():
        _abi.check(lib.cmd_topo_build(topo.handle, C.c_void_p(hptr), 8, B))
        _abi.check(lib.cmd_topo_frame_info(topo.handle, _abi.ptr(h_counts, C.c_int64),
                                           _abi.ptr(h_rebuilt, C.c_uint8), _abi.ptr(h_rsum)))
        if world > 1:
            dist.all_reduce(stats)

    for _ in range(2):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    f1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(f0.elapsed_time(f1), e2e_wall * 1e3)
    clocks = sampler.stop()

    # the same call on float32 host frames (the reference's HDF5 storage: f32 on disk, up-cast to
    # f64 before any arithmetic, trajectory_parser.py:324): half the bytes over PCIe.  Reported
    # beside the headline, which stays on f64 host frames.
    host32 = torch.empty(host.shape, dtype=torch.float32, pin_memory=True)
    host32.copy_(host)
    h32 = host32.data_ptr()

    def step_e2e32():
        _abi.check(lib.cmd_topo_build(topo.handle, C.c_void_p(h32), 4, B))
        _abi.check(lib.cmd_topo_frame_info(topo.handle, _abi.ptr(h_counts, C.c_int64),
                                           _abi.ptr(h_rebuilt, C.c_uint8), _abi.ptr(h_rsum)))

    for _ in range(2):
        step_e2e32()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e32()
    barrier()
    e2e32_ms = (time.perf_counter() - t0) * 1e3
    del host32

    # max over ranks
    tm = torch.tensor([ms_total, e2e_ms, kernel_ms, e2e32_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms, e2e32_ms = [float(x) for x in tm.tolist()]

    value = world * K * B * ppf / (ms_total * 1e-3)
    e2e_value = world * K * B * ppf / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (k_pairs_dense) --------------------------------------
    # HBM: algorithmic bytes per launch = frames read once (24 B per atom-frame) + every listed
    # directed pair written once (start i32, dest i32, dist f64, omega f64 = 24 B), DESIGN.md 3.
    # The kernel is far from that bound: its time goes into instruction issue (17 issue slots per
    # unordered pair in the FP32/INT filter + the FP64 exact stage of the ~4 % survivors), so the
    # issue-slot utilisation and the FP64 peak are reported beside it.
    peak_tf = runtime.fp64_peak_tflops(40000)
    n_img = topo.n_images
    kind = 0 if cell.size == 3 else 1
    out_bytes = float(counts.sum()) * 24.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_gbs = (in_bytes + out_bytes) / (kernel_ms * 1e-3) / 1e9
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    issue_peak = 148 * 4 * sm_mhz * 1e6                     # warp instructions / s
    slots = FILTER_SLOTS_ORTHO if kind == 0 else FILTER_SLOTS_GENERAL + 7 * n_img
    filter_issue = B * ppf * slots / 32.0 / (kernel_ms * 1e-3)
    traffic = None
    try:   # dram bytes per launch of the same kernel from the committed ncu --set full capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_dense_traffic.json")))
        traffic = prof["dram_bytes_per_frame"] * B
    except Exception:
        pass
    roofline = {
        "kernel": "k_pairs_dense", "bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak,
        "unit": "GB/s", "frac": hbm_gbs / hbm_peak, "traffic": traffic,
        "algorithmic_bytes_per_launch": in_bytes + out_bytes,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else
                       "fallback of B200_PROFILING.md",
        "kernel_ms_per_launch": kernel_ms,
        "issue": {"filter_slots_per_pair": slots, "images_kept": n_img,
                  "filter_warp_instr_per_s": filter_issue, "issue_peak_warp_instr_per_s": issue_peak,
                  "filter_share_of_issue_peak": filter_issue / issue_peak},
        "fp64_peak_tflops_measured": peak_tf,
        "reference_equivalent_tflops": B * ppf * (FLOP_ORTHO if kind == 0 else
                                                  FLOP_GENERAL_REFERENCE) / (kernel_ms * 1e-3) / 1e12,
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(3, args.warmup), "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, B, "bruteforce", extra={
            "l2": "inputs (%.0f MB) and outputs (%.0f MB) per step exceed the 126 MB L2" % (
                in_bytes / 1e6, out_bytes / 1e6),
            "directed_pairs_per_frame_mean": float(counts.mean())}),
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": B * 13 + 4, "ms_per_step": e2e_ms / K,
                "api": "cmd_topo_build(host f64 frames) + cmd_topo_frame_info",
                "numa_node_bound": numa_node,
                "f32_storage": {"value": world * K * B * ppf / (e2e32_ms * 1e-3), "unit": UNIT,
                                "h2d_bytes_per_step": in_bytes // 2, "ms_per_step": e2e32_ms / K,
                                "note": "same call on float32 host frames (HDF5 storage layout), "
                                        "up-cast on the device; not the headline"}},
        "roofline": roofline, "clocks": clocks,
    }

    # ---- M2: KMC site-updates/s (own timed region) --------------------------------------------
    if not args.no_m2:
        line["m2"] = run_m2(args, w, box, rate, d_frames, world, dev, dist, barrier)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block(w, args.cpu_seconds)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_m2(args, w, box, rate, d_frames, world, dev, dist, barrier):
    """Verlet-mode topology of a sub-block + Philox KMC of R replicas per GPU (replica-sharded:
    every GPU walks its own replicas over its own frames; statistics all-reduced)."""
    import torch
    from cmdlmc_b200 import runtime, synth
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET
    F = min(args.kmc_frames, d_frames.shape[0])
    R = args.replicas
    n = w.n_oxygen
    topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, 0)
    topo.build_dev(d_frames.data_ptr(), F)
    counts, rebuilt, _ = topo.frame_info()
    # the Verlet pipeline itself (k_dr, rebuild schedule, rebuilds, k_refresh) on the whole block the
    # M1 step used: fresh objects with a known capacity so that every run starts from frame 0
    Fv = d_frames.shape[0]
    vt = []
    for it in range(3):
        tv = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, topo.stride)
        tv.build_dev(d_frames.data_ptr(), min(Fv, 64))        # allocates the block arrays
        tv = None
    tv = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, topo.stride)
    tv.build_dev(d_frames.data_ptr(), Fv)
    for it in range(3):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tv.build_dev(d_frames.data_ptr(), Fv)
        b.record()
        barrier()
        vt.append(a.elapsed_time(b))
    vcounts, vreb, _ = tv.frame_info()
    vms = float(min(vt))
    verlet = {"frames": int(Fv), "ms": vms, "rebuilds": int(vreb.sum()),
              "listed_pair_frames_per_s": float(vcounts.sum()) / vms * 1e3,
              "frames_x_o_pairs_equiv_per_s": Fv * pairs_per_frame(n) / vms * 1e3,
              "list_traffic_gbs": float(vcounts.sum()) * 32.0 / vms / 1e6,
              "note": "continuation blocks of one trajectory (state carried); list traffic = 8 B of "
                      "indices read + 24 B written per listed pair-frame; k_refresh is bound by the "
                      "FP64 pipe (~190 FP64 instructions per pair-frame: reference-order distance "
                      "+ Fermi rate), not by HBM"}
    del tv
    lattices = np.stack([synth.initial_lattice(n, w.n_protons, 4000 + r)[0] for r in range(R)])
    times = []
    updates = 0
    events = 0
    reps = 3
    for it in range(reps + 1):
        kmc = DeviceKMC(box, lattices, w.time_step, RNG_PHILOX, seed=11 + it)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        kmc.advance(topo)
        b.record()
        barrier()
        st = kmc.state()
        if it > 0:
            times.append(a.elapsed_time(b))
            updates = int(st["site_updates"].sum())
            events = int(st["n_events"].sum())
    ms = float(np.mean(times))
    # legacy LMC sweep (row A14): one sweep = P attempts per frame and replica, Philox
    from cmdlmc_b200.lmc import DeviceLMC
    lmc_times, lmc_attempts, lmc_jumps = [], 0, 0
    for it in range(3):
        lmc = DeviceLMC(lattices, RNG_PHILOX, seed=21 + it)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lmc.advance(topo, w.time_step, 1)
        b.record()
        barrier()
        if it > 0:
            lmc_times.append(a.elapsed_time(b))
            ls = lmc.state()
            lmc_attempts, lmc_jumps = int(ls["attempts"].sum()), int(ls["jumps"].sum())
    lmc_ms = float(np.mean(lmc_times))
    tl = torch.tensor([lmc_ms], dtype=torch.float64, device=dev)
    tot_l = torch.tensor([float(lmc_attempts), float(lmc_jumps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_l)
    lmc_block = {"metric": "LMC site-updates/s (jump attempts)", "kernel": "k_lmc_sweep",
                 "value": float(tot_l[0].item()) / (float(tl.item()) * 1e-3), "unit": "attempts/s",
                 "ms": float(tl.item()), "jumps": float(tot_l[1].item()), "sweeps_per_frame": 1,
                 "rng": "philox4x32-10", "parity": "unpinned upstream (engine not in the reference tree)"}
    # one replica in exact-replay mode: what a reference `mdmc` run is (k_kmc_solo, one CTA)
    from cmdlmc_b200.kmc import RNG_REPLAY
    Fs = min(F, 2048)
    u = np.random.RandomState(5).random_sample((1, 32 * Fs + 1000))
    solo_ms = []
    for it in range(2):
        one = DeviceKMC(box, lattices[:1], w.time_step, RNG_REPLAY)
        one.set_replay_stream(u)
        # a topology view of the first Fs frames is not needed: the kernel walks the block it is given,
        # so time a block of its own
        if it == 0:
            ts = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, topo.stride)
            ts.build_dev(d_frames.data_ptr(), Fs)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one.advance(ts)
        b.record()
        torch.cuda.synchronize()
        solo_ms.append(a.elapsed_time(b))
        solo_events = int(one.state()["n_events"].sum())
    single = {"kernel": "k_kmc_solo", "rng": "replay (reference np.random protocol, bit-exact mode)",
              "frames": int(Fs), "ms": float(min(solo_ms)), "events": solo_events,
              "frames_per_s": Fs / (min(solo_ms) * 1e-3),
              "site_updates_per_s": float(counts[:Fs].sum()) / (min(solo_ms) * 1e-3)}
    tm = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(updates), float(events)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms = float(tm.item())
    updates, events = [float(x) for x in tot.tolist()]
    rate_su = updates / (ms * 1e-3)
    return {"metric": "KMC site-updates/s", "value": rate_su, "unit": "site-updates/s",
            "replicas_per_gpu": R, "frames": F, "ms": ms, "events": events,
            "rng": "philox4x32-10", "directed_pairs_per_frame_mean": float(counts.mean()),
            "verlet_rebuilds": int(rebuilt.sum()), "verlet_pipeline": verlet, "lmc_sweep": lmc_block,
            "single_replica_replay": single, "kernel": "k_kmc_stream",
            "roofline": {"bound": "smem", "unit": "GB/s", "achieved": rate_su * 16 / 1e9,
                         "peak": 148 * 128 * 1.965, "frac": rate_su * 16 / 1e9 / (148 * 128 * 1.965),
                         "note": "16 B of (start, dest, omega) read from the shared-memory ring per "
                                 "site-update; peak = 148 SMs x 128 B/clk x 1.965 GHz"},
            "hbm_algorithmic_gbs": float(counts.sum()) * 16.0 * world / (ms * 1e-3) / 1e9}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
