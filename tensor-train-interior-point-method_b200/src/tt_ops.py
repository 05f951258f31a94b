"""Same module name as reference src/tt_ops.py; everything comes from the B200 path."""
from ttipm_b200.tt_ops import *  # noqa: F401,F403
from ttipm_b200.tt_ops import E, cached_einsum, np, scp  # noqa: F401
