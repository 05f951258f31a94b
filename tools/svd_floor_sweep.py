"""Effect of the Jacobi noise floor (ttipm_linalg_noise_floor) on a traced solve: ranks, sweeps, residual and time
against the committed oracle record.  python tools/svd_floor_sweep.py maxcut_13 0 1e-4 1e-2 1"""
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "oracle")]
import golden_io as G  # noqa: E402
import tt_oracle as O  # noqa: E402
from ttipm_b200 import get_runtime  # noqa: E402
from ttipm_b200.amen import NativeBlockAmen  # noqa: E402


def main():
    wl = sys.argv[1]
    f = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", f"amen_{wl}_*.npz")))[-1]
    rec = json.load(open(f.replace("amen_", "oracle_amen_").replace(".npz", ".json")))
    print("oracle", {k: rec[k] for k in ("ranks", "sweeps", "nsolves", "res", "krylov_its")})
    g = G.load_amen(f)
    rt = get_runtime()
    for fac in [float(v) for v in sys.argv[2:]]:
        if fac < 0:                         # negative: toggle the early exit of the Jacobi iteration off, floor 0
            rt.lib.ttipm_linalg_early_exit(0)
            fac = 0.0
        if fac >= 256:                      # 256 / 512: threads per CTA of the QR / SVD kernel instead of a floor factor
            rt.lib.ttipm_linalg_threads(int(fac))
            fac = 0.0
        rt.lib.ttipm_linalg_noise_floor(fac)
        best = None
        for _ in range(3):
            np.random.set_state(g["rng_state"])
            x0 = [c.copy() for c in g["x0"]] if g["x0"] is not None else None
            if x0 is not None:
                x0 = O.tt_rank_retraction(x0, [len(x0)] * (len(x0) - 1))
            s = NativeBlockAmen(g["A"], g["aliases"], g["transposes"], g["b"], g["ineq"], rt=rt)
            rt.sync()
            t0 = time.perf_counter()
            x, res = s.solve(g["termination_tol"], r_max=g["rank_restriction"], eps=g["eps"], nswp=g["inner_m"], x0=x0,
                             kick_rank=2, amen=True)
            rt.sync()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(json.dumps(dict(floor=fac, seconds=best, ranks=s.ranks, sweeps=s.sweeps, nsolves=len(s.trace), res=res,
                              ranks_match=list(s.ranks) == list(rec["ranks"]))), flush=True)


if __name__ == "__main__":
    main()
