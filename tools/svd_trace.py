"""Per-call trace of the SVD kernel inside one traced block AMEn solve (Python-driven sweep, same kernels):
shape, Jacobi sweeps and the kernel's own phase timers.  python tools/svd_trace.py [workload]"""
import ctypes as C
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "oracle")]
import golden_io as G  # noqa: E402
import tt_oracle as O  # noqa: E402
from ttipm_b200 import get_runtime, kernels as K  # noqa: E402
from ttipm_b200.amen import DeviceBlockAmen  # noqa: E402

LOG = []


def traced_svd(A, rt=None):
    rt = rt or get_runtime()
    A3 = A if A.dim() == 3 else A.unsqueeze(0)
    nb, M, N = A3.shape
    Kk = min(M, N)
    U, S, W = rt.empty(nb, M, Kk), rt.empty(nb, Kk), rt.empty(nb, Kk, N)
    ws = rt.empty(int(rt.lib.ttipm_svd_workspace(M, N, nb)))
    info = torch.zeros(16 * nb, dtype=torch.int32, device=A.device)
    rt.check(rt.lib.ttipm_svd_left(K._ptr(A3), A3.stride(1), A3.stride(2), A3.stride(0) if nb > 1 else 0, M, N, K._ptr(U),
                                   K._ptr(S), K._ptr(W), K._ptr(ws), C.c_void_p(info.data_ptr()), nb, rt.stream()), "svd")
    inf = info.cpu().numpy().tolist()
    s = S[0].cpu().numpy()
    LOG.append(dict(M=M, N=N, sweeps=inf[0], qr_us=inf[1] / 1e3, q_us=inf[2] / 1e3, jac_us=inf[3] / 1e3, grid=inf[4],
                    total_us=inf[10] / 1e3, s_max=float(s[0]), s_min=float(s[-1]), s_med=float(np.median(s))))
    return (U, S, W) if A.dim() == 3 else (U[0], S[0], W[0])


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "maxcut_13"
    f = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", f"amen_{wl}_*.npz")))[-1]
    g = G.load_amen(f)
    rt = get_runtime()
    K.svd_left = traced_svd
    import ttipm_b200.amen as A
    A.K.svd_left = traced_svd
    np.random.set_state(g["rng_state"])
    x0 = [c.copy() for c in g["x0"]] if g["x0"] is not None else None
    if x0 is not None:
        x0 = O.tt_rank_retraction(x0, [len(x0)] * (len(x0) - 1))
    s = DeviceBlockAmen(g["A"], g["aliases"], g["transposes"], g["b"], g["ineq"], rt=rt)
    x, res = s.solve(g["termination_tol"], r_max=g["rank_restriction"], eps=g["eps"], nswp=g["inner_m"], x0=x0, kick_rank=2,
                     amen=True)
    print("res", res, "svd calls", len(LOG), "total ms", sum(r["total_us"] for r in LOG) / 1e3)
    for r in sorted(LOG, key=lambda r: -r["total_us"])[:25]:
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) and v > 1 else v) for k, v in r.items()}))


if __name__ == "__main__":
    main()
