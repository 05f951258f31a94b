"""A few left-SVD launches of one small unfolding for ncu source captures:  python tools/prof_svd_small.py 216 6"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tools")]
from ttipm_b200 import get_runtime  # noqa: E402
from ttipm_b200.kernels import _ptr  # noqa: E402
from bench_svd import plateau  # noqa: E402


def main():
    M, N = int(sys.argv[1]), int(sys.argv[2])
    rt = get_runtime()
    a = plateau(M, N, np.random.default_rng(0))
    K = min(M, N)
    A = rt.to_device(a)
    U, S, W = rt.empty(M, K), rt.empty(K), rt.empty(K, N)
    ws = rt.empty(int(rt.lib.ttipm_svd_workspace(M, N, 1)))
    info = torch.zeros(16, dtype=torch.int32, device=A.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(6):
        if it == 1:
            torch.cuda.synchronize()
            e0.record()
        rt.check(rt.lib.ttipm_svd_left(_ptr(A), N, 1, 0, M, N, _ptr(U), _ptr(S), _ptr(W), _ptr(ws), C.c_void_p(info.data_ptr()),
                                       1, rt.stream()), "svd")
    e1.record()
    torch.cuda.synchronize()
    print("us per call", e0.elapsed_time(e1) * 1e3 / 5, "info", info.cpu().numpy().tolist())


if __name__ == "__main__":
    main()
