"""tt_inner_prod on device-resident trains: fused single-launch chain (csrc/ttops.cu k_tt_inner) vs what it replaced
(two GEMM launches per core), wall time per call incl. the scalar read-back.  python tools/bench_inner.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
from ttipm_b200 import get_runtime, tt as T, use_runtime  # noqa: E402


def main():
    rt = get_runtime()
    rng = np.random.default_rng(0)
    with use_runtime(rt):
        for d, rk, nn in ((5, 4, 4), (10, 4, 16), (13, 16, 4), (13, 40, 4), (13, 100, 4)):
            shapes = [(1 if k == 0 else rk, nn, 1 if k == d - 1 else rk) for k in range(d)]
            a = [rng.standard_normal(s) / np.sqrt(s[0] * s[1]) for s in shapes]
            b = [rng.standard_normal(s) / np.sqrt(s[0] * s[1]) for s in shapes]
            da, db = T.tt_add(a, [0 * c for c in a]), T.tt_add(b, [0 * c for c in b])     # lazy device trains
            v = T.tt_inner_prod(da, db)
            rt.sync()
            t0 = time.perf_counter()
            for _ in range(50):
                v = T.tt_inner_prod(da, db)
            rt.sync()
            us = (time.perf_counter() - t0) / 50 * 1e6
            print(json.dumps(dict(d=d, rank=2 * rk, mode=nn, us_per_call=us, value=v)), flush=True)


if __name__ == "__main__":
    main()
