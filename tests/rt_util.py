"""Runtimes for the two test tiers: the CUDA library on a GPU, or the -DTTIPM_EMU build of the
same kernel sources on CPU tensors (tests/emu)."""
import functools
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "tensor-train-interior-point-method_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


@functools.lru_cache(maxsize=None)
def emu_runtime():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from ttipm_b200 import Runtime
    return Runtime(lib_path=build_emu.build(), device="cpu")


@functools.lru_cache(maxsize=None)
def cuda_runtime():
    from ttipm_b200 import get_runtime
    return get_runtime()
