"""Small-regime Krylov kernel: per-launch and per-inner-iteration cost on the fixture local systems.
   python tools/bench_lgmres_small.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "oracle")]
import kernel_cases as KC  # noqa: E402
from ttipm_b200 import get_runtime  # noqa: E402


def ev_us(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    rt = get_runtime()
    for case in ("eq_small", "ineq_small", "eq_mid"):
        c = KC.load_blp_case(case)
        P1 = {k: v.copy() for k, v in c["P1"].items()}
        for key in ((0, 0), (2, 1), (3, 3)):
            if key in P1:
                P1[key] = P1[key] + 8.0 * np.stack([np.eye(P1[key].shape[0])] * P1[key].shape[1], axis=1)
        op, ineq = KC._reduced_op(rt, dict(c, P1=P1))
        r, n, R = c["inv_I"].shape
        nb = 3 if ineq else 2
        b = rt.to_device(np.random.default_rng(5).standard_normal(nb * r * n * R))
        x = rt.to_device(c["red_x"])
        for stage in (0,):
            t_apply = ev_us(lambda: op.matvec(x, grid_hint=1), 200)
            out = {}
            for max_it in (1, 11, 41):
                restart = min(nb * r * n * R // nb, 100)
                t = ev_us(lambda: op.solve(b, restart, max(restart // 10, 3), max_it=max_it, rtol=1e-30, grid_hint=1), 50)
                out[max_it] = t
            print(json.dumps(dict(case=case, r=r, R=R, nv=nb * r * n * R, apply_launch_us=t_apply,
                                  solve_us=out, us_per_inner_iteration=(out[41] - out[11]) / 30.0,
                                  us_per_early_iteration=(out[11] - out[1]) / 10.0)), flush=True)


if __name__ == "__main__":
    main()
