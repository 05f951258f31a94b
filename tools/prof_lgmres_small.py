"""One small-regime Krylov solve (fixture eq_small, 41 inner steps) for ncu source-level captures."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "oracle")]
import kernel_cases as KC  # noqa: E402
from ttipm_b200 import get_runtime  # noqa: E402

rt = get_runtime()
case = sys.argv[1] if len(sys.argv) > 1 else "eq_small"
c = KC.load_blp_case(case)
P1 = {k: v.copy() for k, v in c["P1"].items()}
for key in ((0, 0), (2, 1), (3, 3)):
    if key in P1:
        P1[key] = P1[key] + 8.0 * np.stack([np.eye(P1[key].shape[0])] * P1[key].shape[1], axis=1)
op, ineq = KC._reduced_op(rt, dict(c, P1=P1))
r, n, R = c["inv_I"].shape
nb = 3 if ineq else 2
b = rt.to_device(np.random.default_rng(5).standard_normal(nb * r * n * R))
restart = min(r * n * R, 100)
for _ in range(3):
    x, info = op.solve(b, restart, max(restart // 10, 3), max_it=41, rtol=1e-30, grid_hint=1)
rt.sync()
print("ok", rt.to_host(info))
