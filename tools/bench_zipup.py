"""Zip-up product at large vector ranks (the dual update `tt_fast_matrix_vec_mul(lag_map_y, Y)` of src/tt_ipm.py:1074 with
the rank-~100 Newton direction of maxcut_13): time and result ranks per vector rank."""
import json
import sys
import time

import numpy as np

sys.path[:0] = ["tensor-train-interior-point-method_b200"]
from ttipm_b200 import get_runtime, tt as T  # noqa: E402


def rand_tt(rng, d, r, mode):
    rr = [1] + [r] * (d - 1) + [1]
    return [rng.standard_normal((a, *mode, b)) / np.sqrt(a * b) for a, b in zip(rr[:-1], rr[1:])]


def main():
    rt = get_runtime()
    rng = np.random.default_rng(0)
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    for r in (8, 16, 32, 64, 100):
        op = rand_tt(rng, d, 2, (4, 4))
        v = rand_tt(rng, d, r, (4,))
        v = T.tt_rank_reduce(v, 1e-12)
        rt.sync()
        t0 = time.perf_counter()
        out = T.tt_fast_matrix_vec_mul(op, v, 1e-10)
        ranks = T.tt_ranks(out)
        rt.sync()
        dt = time.perf_counter() - t0
        print(json.dumps(dict(d=d, vec_rank=r, seconds=dt, out_ranks=ranks)), flush=True)


if __name__ == "__main__":
    main()
