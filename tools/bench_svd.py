"""Micro-benchmark of the left-SVD / QR kernels (GPU box): single-CTA vs cooperative multi-CTA path, CUDA events,
with the cooperative kernel's own phase timers (info[1..3])."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
from ttipm_b200 import get_runtime  # noqa: E402
from ttipm_b200.kernels import _ptr  # noqa: E402


def graded(M, N, decades, rng):
    K = min(M, N)
    U, _ = np.linalg.qr(rng.standard_normal((M, K)))
    V, _ = np.linalg.qr(rng.standard_normal((N, K)))
    return (U * np.logspace(0, -decades, K)) @ V.T


def plateau(M, N, rng):
    """the spectrum traced from the maxcut_13 sweep: a third decaying over 9 decades, the rest rounding noise"""
    K = min(M, N)
    U, _ = np.linalg.qr(rng.standard_normal((M, K)))
    V, _ = np.linalg.qr(rng.standard_normal((N, K)))
    sv = np.concatenate([np.logspace(0, -9, (K + 2) // 3), 1e-14 * rng.uniform(0.1, 1.0, K - (K + 2) // 3)])
    return 12000.0 * (U * sv) @ V.T


def run(rt, a, coop, iters=5):
    M, N = a.shape
    K = min(M, N)
    old = rt.lib.ttipm_linalg_coop_min_dim(1 if coop else 1 << 30)
    try:
        A = rt.to_device(a)
        U, S, W = rt.empty(M, K), rt.empty(K), rt.empty(K, N)
        ws = rt.empty(int(rt.lib.ttipm_svd_workspace(M, N, 1)))
        info = torch.zeros(16, dtype=torch.int32, device=A.device)

        def call():
            rt.check(rt.lib.ttipm_svd_left(_ptr(A), N, 1, 0, M, N, _ptr(U), _ptr(S), _ptr(W), _ptr(ws),
                                           C.c_void_p(info.data_ptr()), 1, rt.stream()), "svd")
        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        Uh, Sh, Wh = rt.to_host(U), rt.to_host(S), rt.to_host(W)
        sref = np.linalg.svd(a, compute_uv=False)
        inf = info.cpu().numpy().tolist()
        return dict(M=M, N=N, coop=coop, ms=ms, sweeps=inf[0], qr_us=inf[1] / 1e3, q_us=inf[2] / 1e3, jac_us=inf[3] / 1e3,
                    grid=inf[4], nb=inf[5], jac_load_us=inf[6] / 1e3, jac_rot_us=inf[7] / 1e3, jac_store_us=inf[8] / 1e3,
                    jac_sync_us=inf[9] / 1e3, total_us=inf[10] / 1e3, cluster=inf[11], qr_load_us=inf[12] / 1e3,
                    qr_factor_us=inf[13] / 1e3, qr_trail_us=inf[14] / 1e3, qr_sync_us=inf[15] / 1e3, s_err=float(np.max(np.abs(Sh - sref)) / sref[0]),
                    rec=float(np.linalg.norm(Uh @ Wh - a) / np.linalg.norm(a)),
                    orth=float(np.linalg.norm(Uh.T @ Uh - np.eye(K))))
    finally:
        rt.lib.ttipm_linalg_coop_min_dim(old)


def main():
    rt = get_runtime()
    rng = np.random.default_rng(0)
    shapes = [(16, 12), (32, 24), (64, 48), (88, 66), (136, 81), (296, 102), (102, 296), (220, 165), (165, 220), (440, 330),
              (330, 440), (400, 156), (208, 96), (24, 440)]
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--single-qr" in sys.argv:
        rt.lib.ttipm_linalg_tall_triple_qr(0)       # tall matrices: one QR instead of three before the Jacobi sweeps
    if "--no-early-exit" in sys.argv:
        rt.lib.ttipm_linalg_early_exit(0)           # always run the confirming all-skip sweep (round-2 behaviour)
    for a in sys.argv[1:]:
        if a.startswith("--nb="):
            rt.lib.ttipm_linalg_block_rows(int(a[5:]))
    if args:
        shapes = [tuple(int(v) for v in s.split("x")) for s in args]
    out = []
    for (M, N) in shapes:
        for kind, a in (("graded18", graded(M, N, 18, rng)), ("plateau", plateau(M, N, rng))):
            for coop in (True, False):
                if not coop and min(M, N) > 128:
                    continue                    # one CTA per matrix is for small / batched unfoldings
                rec = run(rt, a, coop)
                rec["spectrum"] = kind
                print(json.dumps(rec), flush=True)
                out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_svd_single.jsonl" if "--single-qr" in sys.argv else "bench_svd.jsonl"), "w") as f:
        for rec in out:
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
