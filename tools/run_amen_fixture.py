"""Run ONE traced / dumped KKT system (amen_*.npz layout) through the native block AMEn driver on the GPU and print
what happened: time, sweeps, ranks, local-solve trace.  With TTIPM_AMEN_LOG=1 the driver logs every core step to
stderr, which is how a solve that does not come back is located.

  TTIPM_AMEN_LOG=1 python tools/run_amen_fixture.py gpurun_out/amen_dump_0.npz [--restarted] [--profile]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tests")]
import golden_io as G  # noqa: E402
from ttipm_b200 import get_runtime  # noqa: E402
from ttipm_b200 import tt as T  # noqa: E402
from ttipm_b200.amen import NativeBlockAmen  # noqa: E402


def main():
    path = sys.argv[1]
    g = G.load_amen(path)
    rt = get_runtime()
    for a in sys.argv[2:]:
        if a.startswith("--big-min-flops="):      # flop threshold of the grouped-GEMM form of the block matvec
            rt.lib.ttipm_matvec_big_min_flops(float(a.split("=")[1]))
        if a.startswith("--host-krylov="):       # 0 never, 1 automatic (default), 2 always
            rt.lib.ttipm_amen_host_krylov(int(a.split("=")[1]))
    print(json.dumps(dict(file=os.path.basename(path), d=g["d"], ineq=g["ineq"], rank_restriction=g["rank_restriction"],
                          termination_tol=g["termination_tol"], eps=g["eps"], inner_m=g["inner_m"],
                          op_ranks={f"{k[0]}{k[1]}": max(c.shape[0] for c in v) for k, v in g["A"].items()},
                          x0_ranks=[c.shape[-1] for c in g["x0"]] if g["x0"] is not None else None)), flush=True)
    np.random.set_state(g["rng_state"])
    x0 = [np.asarray(c).copy() for c in g["x0"]] if g["x0"] is not None else None
    if x0 is not None:
        x0 = [np.asarray(c) for c in T.tt_rank_retraction(x0, [len(x0)] * (len(x0) - 1))]
    profile = "--profile" in sys.argv
    for rep in range(2 if profile else 1):         # --profile: one warm-up solve, then the instrumented one
        np.random.set_state(g["rng_state"])
        s = NativeBlockAmen(g["A"], g["aliases"], g["transposes"], g["b"], g["ineq"], rt=rt,
                            stats={"profile": {}} if (profile and rep == 1) else None)
        if rep == 0 and profile:
            s.solve(g["termination_tol"], r_max=g["rank_restriction"], eps=g["eps"], nswp=g["inner_m"],
                    x0=[c.copy() for c in x0] if x0 is not None else None, kick_rank=2, amen=True)
    rt.sync()
    t0 = time.perf_counter()
    x, res = s.solve(g["termination_tol"], r_max=g["rank_restriction"], eps=g["eps"], nswp=g["inner_m"], x0=x0, kick_rank=2,
                     amen=True)
    rt.sync()
    print(json.dumps(dict(seconds=time.perf_counter() - t0, res=res, sweeps=s.sweeps, ranks=list(s.ranks),
                          local_solves=len(s.trace), stats={k: v for k, v in s.stats.items() if not isinstance(v, list)})),
          flush=True)
    if profile:
        # CUDA events around every launch of the native driver: seconds, algorithmic flops (bytes for the memory-bound
        # helpers), launches per kernel category
        for name, (sec, work, nl) in s.native_profile.items():
            if nl:
                print(json.dumps(dict(category=name, ms=sec * 1e3, launches=int(nl), work=work,
                                      rate_T_per_s=work / sec / 1e12 if sec > 0 else None)), flush=True)


if __name__ == "__main__":
    main()
