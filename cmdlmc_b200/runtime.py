"""Process-level runtime: one process drives one GPU (torch.distributed launches one rank per
GPU; LOCAL_RANK picks the device).  PyTorch is used only as plumbing -- device memory, streams,
torch.distributed -- never for the arithmetic of the hot path."""
import os

from . import _abi

_state = {"device": None}


def device_count():
    import ctypes as C
    n = C.c_int(0)
    _abi.lib().cmd_device_count(C.byref(n))
    return n.value


def init(device=None):
    """Binds this process to a CUDA device.  Raises if there is none (no CPU fallback)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _state["device"] == device:
        return device
    _abi.check(_abi.lib().cmd_init(int(device)))
    _state["device"] = device
    return device


def ensure_init():
    if _state["device"] is None:
        init()
    return _state["device"]


def use_torch_stream():
    """Launch on PyTorch's current stream so torch.cuda.Event timing brackets our kernels."""
    import torch
    ensure_init()
    torch.cuda.set_device(_state["device"])
    s = torch.cuda.current_stream().cuda_stream
    _abi.check(_abi.lib().cmd_set_stream(s))
    return s


def sync():
    _abi.check(_abi.lib().cmd_sync())


def launch_count():
    return int(_abi.lib().cmd_launch_count())


def _host_free(address):
    import ctypes as C
    try:
        _abi.lib().cmd_host_free(C.c_void_p(address))
    except Exception:
        pass


def pinned_empty(shape, dtype):
    """NumPy array in page-locked host memory (cmd_host_alloc): what a trajectory reader should fill
    so that cmd_topo_build copies from it directly and the copy overlaps the kernels
    (trajectory_parser.py:296,322 reads its 1000-frame chunks into plain arrays).  The memory is
    freed when the last array (or view) referring to it goes away."""
    import ctypes as C
    import weakref
    import numpy as np
    ensure_init()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dtype.itemsize, 1)
    p = C.c_void_p()
    _abi.check(_abi.lib().cmd_host_alloc(nbytes, C.byref(p)))
    buf = (C.c_ubyte * nbytes).from_address(p.value)
    # the arrays keep `buf` alive through the buffer protocol; the allocation follows its lifetime
    weakref.finalize(buf, _host_free, p.value)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def staging_stats():
    """(bytes uploaded through the library's staging ring, bytes copied straight from page-locked
    memory, host copy threads of the ring)."""
    import ctypes as C
    a, b, t = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
    _abi.check(_abi.lib().cmd_staging_stats(C.byref(a), C.byref(b), C.byref(t)))
    return a.value, b.value, t.value


def fp64_peak_tflops(iters=20000):
    import ctypes as C
    ensure_init()
    v = C.c_double(0)
    _abi.check(_abi.lib().cmd_fp64_peak(int(iters), C.byref(v)))
    return v.value


def smem_peak_gbs(iters=4000):
    """Measured shared-memory read rate of the whole GPU in GB/s (cmd_smem_peak)."""
    import ctypes as C
    ensure_init()
    v = C.c_double(0)
    _abi.check(_abi.lib().cmd_smem_peak(int(iters), C.byref(v)))
    return v.value


def numa_topology(device=None):
    """What the host says about NUMA placement of this GPU: node count of the machine, the node of
    the GPU's PCIe function (-1: the platform reports none, as on single-node or virtualised
    hosts) and the CPUs this process may run on.  Facts for the bench line, nothing is changed."""
    info = {"nodes": None, "gpu_node": None, "cpus_allowed": None}
    try:
        info["cpus_allowed"] = len(os.sched_getaffinity(0))
        info["nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except OSError:
        pass
    try:
        import torch
        dev = _state["device"] if device is None else device
        p = torch.cuda.get_device_properties(0 if dev is None else dev)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            info["gpu_node"] = int(f.read().strip())
    except Exception:
        pass
    return info


def bind_to_gpu_numa_node(device=None):
    """Pins this process (one process per GPU) to the CPUs of the NUMA node its GPU hangs off, so
    that pinned staging buffers allocated afterwards are local to the GPU's PCIe root: with eight
    ranks streaming frames concurrently, cross-socket copies otherwise halve the host->device
    rate.  Best effort: returns the node, or None when the topology cannot be read."""
    try:
        import torch
        dev = _state["device"] if device is None else device
        if dev is None:
            dev = int(os.environ.get("LOCAL_RANK", "0"))
        p = torch.cuda.get_device_properties(dev)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
