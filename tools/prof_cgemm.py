"""One grouped-GEMM local matvec (K1 large-rank path) on a scaled-grid shape, for ncu captures:
   python tools/prof_cgemm.py [r] [s] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
from ttipm_b200 import get_runtime, kernels as K  # noqa: E402


def main():
    r = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    s = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rt = get_runtime()
    rng = np.random.default_rng(0)
    ranks = {(0, 0): s, (0, 1): s, (1, 2): 1, (2, 1): s, (2, 2): s}
    dev = rt.to_device
    A = {k: dev(rng.standard_normal((q, 4, 4, q))) for k, q in ranks.items()}
    P1 = {k: dev(rng.standard_normal((r, q, r))) for k, q in ranks.items()}
    P2 = {k: dev(rng.standard_normal((r, q, r))) for k, q in ranks.items()}
    x = dev(rng.standard_normal((r, 3, 4, r)))
    tl = K.TermList()
    for (i, j) in ranks:
        tl.add(P1[i, j], A[i, j], P2[i, j], j, i)
        if (i, j) == (0, 1):
            tl.add(P1[i, j].permute(2, 1, 0), A[i, j].permute(0, 2, 1, 3), P2[i, j].permute(2, 1, 0), 0, 1)
    rt.lib.ttipm_matvec_big_min_flops(0.0)
    if "CG_CFG" in os.environ:
        rt.lib.ttipm_cgemm_force_cfg(int(os.environ["CG_CFG"]))
    for _ in range(reps):
        y = K.block_matvec(tl, x, 3, (r, r), rt=rt)
    rt.sync()
    print("ok", float(y.abs().sum()))


if __name__ == "__main__":
    main()
