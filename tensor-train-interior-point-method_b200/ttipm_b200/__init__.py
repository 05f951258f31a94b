"""ttipm_b200 -- B200-native (sm_100a) TT-IPM Newton-system hot path.

Host-side Python over the C ABI of libttipm_b200.so (include/ttipm.h).  PyTorch is used for
device memory and streams only; all arithmetic runs in the hand-written CUDA kernels.  There is
no CPU fallback: creating the default runtime without a CUDA device raises.
"""
from .runtime import Runtime, get_runtime, use_runtime  # noqa: F401

__all__ = ["Runtime", "get_runtime", "use_runtime"]
