"""Micro-benchmarks of the hot-path kernels with CUDA events (GPU box).  Writes JSON lines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "oracle")]
from ttipm_b200 import get_runtime, kernels as K  # noqa: E402


def term_flops(l, L, r, R, s, S, n=4):
    """2 r n R L S + 2 r L s n n S + 2 l n L r s (the reference's 3-GEMM order)."""
    return 2 * r * n * R * L * S + 2 * r * L * s * n * n * S + 2 * l * n * L * r * s


def time_cuda(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters   # us


def main():
    """K1 on traced shapes and on the scaled synthetic grid: fused one-CTA-per-slab kernel vs the grouped
    contraction-GEMM path, CUDA events over back-to-back launches, against cuBLAS DGEMM measured here."""
    rt = get_runtime()
    rng = np.random.default_rng(0)
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    us = time_cuda(lambda: a @ a, iters=5, warm=2)
    peak = 2.0 * n ** 3 / us * 1e-6          # TFLOP/s
    print(json.dumps(dict(kernel="cublas_dgemm_4096", tflops=peak)), flush=True)
    out = [dict(kernel="cublas_dgemm_4096", tflops=peak)]
    eq = lambda s: {(0, 0): s, (0, 1): s, (1, 2): 1, (2, 1): s, (2, 2): s}
    shapes = [("maxcut13", 55, 55, {(0, 0): 2, (0, 1): 1, (1, 2): 1, (2, 1): 5, (2, 2): 5}),
              ("graphm3", 29, 44, {(0, 0): 5, (0, 1): 10, (1, 2): 1, (2, 1): 1, (2, 2): 4})]
    for r in (64, 128, 256):
        for s in (8, 16, 32):
            shapes.append((f"grid_r{r}_s{s}", r, r, eq(s)))
    for name, r, R, ranks in shapes:
        dev = lambda a: rt.to_device(a)
        A = {k: dev(rng.standard_normal((s, 4, 4, s))) for k, s in ranks.items()}
        P1 = {k: dev(rng.standard_normal((r, s, r))) for k, s in ranks.items()}
        P2 = {k: dev(rng.standard_normal((R, s, R))) for k, s in ranks.items()}
        x = dev(rng.standard_normal((r, 3, 4, R)))
        tl = K.TermList()
        flops = 0
        for (i, j), s in ranks.items():
            tl.add(P1[i, j], A[i, j], P2[i, j], j, i)
            flops += term_flops(r, R, r, R, s, s)
            if (i, j) == (0, 1):
                tl.add(P1[i, j].permute(2, 1, 0), A[i, j].permute(0, 2, 1, 3), P2[i, j].permute(2, 1, 0), 0, 1)
                flops += term_flops(r, R, r, R, s, s)
        for path, thr in (("fused", 1e30), ("grouped_gemm", 0.0)):
            old = rt.lib.ttipm_matvec_big_min_flops(thr)
            try:
                us = time_cuda(lambda: K.block_matvec(tl, x, 3, (r, R), rt=rt), iters=20, warm=3)
                rec = dict(kernel="block_matvec", path=path, shape=name, r=r, R=R, us=us, tflops=flops / us * 1e-6,
                           frac_of_dgemm=flops / us * 1e-6 / peak, flops=flops)
            except Exception as e:   # report, do not hide
                rec = dict(kernel="block_matvec", path=path, shape=name, error=str(e)[:200])
            finally:
                rt.lib.ttipm_matvec_big_min_flops(old)
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_kernels.jsonl"), "w") as f:
        for rec in out:
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
