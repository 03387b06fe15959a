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


def fp64_peak_tflops(iters=20000):
    import ctypes as C
    ensure_init()
    v = C.c_double(0)
    _abi.check(_abi.lib().cmd_fp64_peak(int(iters), C.byref(v)))
    return v.value
