"""Runtime = loaded C-ABI library + the torch device that owns the buffers.

The default runtime binds libttipm_b200.so to the current CUDA device and refuses to exist
without one.  tests/ may install another runtime (the -DTTIPM_EMU build of the same kernel
sources on torch CPU tensors) with use_runtime(); product code never does.
"""
import contextlib

import numpy as np
import torch

from . import _cabi

_default = None
_override = None


class TTIPMError(RuntimeError):
    pass


class Runtime:
    def __init__(self, lib_path=None, device=None):
        if device is None:
            if not torch.cuda.is_available():
                raise TTIPMError("ttipm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.lib = _cabi.load(lib_path)
        self.is_cuda = self.device.type == "cuda"
        self.launches = 0
        sm = _cabi.C.c_int(0)
        sh = _cabi.C.c_int(0)
        if self.is_cuda:
            torch.cuda.init()
            with torch.cuda.device(self.device):
                torch.zeros(1, device=self.device)
                self.lib.ttipm_device_info(sm, sh)
        else:
            self.lib.ttipm_device_info(sm, sh)
        self.sm_count, self.smem_optin = sm.value, sh.value

    # ---- memory ------------------------------------------------------------------------
    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64, device=self.device)

    def to_device(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        t = torch.from_numpy(a)
        return t.to(self.device, non_blocking=False) if self.is_cuda else t.clone()

    def to_host(self, t):
        return t.detach().cpu().numpy() if self.is_cuda else t.detach().numpy().copy()

    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.is_cuda else 0

    def sync(self):
        if self.is_cuda:
            torch.cuda.synchronize(self.device)

    def check(self, code, what):
        self.launches += 1
        if code != 0:
            raise TTIPMError(f"{what} failed ({code}): {self.lib.ttipm_last_error().decode()}")


def get_runtime():
    global _default
    if _override is not None:
        return _override
    if _default is None:
        _default = Runtime()
    return _default


@contextlib.contextmanager
def use_runtime(rt):
    """Test hook: temporarily route the host code through another Runtime."""
    global _override
    prev = _override
    _override = rt
    try:
        yield rt
    finally:
        _override = prev
