"""AMEn-level parity: the device sweep vs the oracle on KKT systems traced from the reference."""
import glob
import os
import time

import numpy as np

import golden_io as G
import tt_oracle as O


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def dense_block(tt, i):
    t = O.tt_get_block(i, tt)
    d = t[0]
    for c in t[1:]:
        d = np.tensordot(d, c, axes=(-1, 0))
    return d


def block_inner(bm, x, y=None):
    """<x, y> summed over blocks, in TT form (size independent)."""
    y = x if y is None else y
    nb = O.block_size_of(bm)
    return sum(O.tt_inner_prod(O.tt_get_block(i, x), O.tt_get_block(i, y)) for i in range(nb))


def run_block_amen(rt, path, use_oracle=True, native=True):
    """tt_block_amen on one traced system with the device sweep (and optionally the oracle) from the
    same warm start and RNG state; returns a dict of comparisons."""
    from ttipm_b200.amen import DeviceBlockAmen, NativeBlockAmen
    g = G.load_amen(path)
    bm = O.BlockMatrix(g["A"], g["aliases"], g["transposes"])
    bv = O.BlockVector(g["b"])
    ls = O.local_solver_ineq if g["ineq"] else O.local_solver_eq

    def prep():
        np.random.set_state(g["rng_state"])
        x0 = [c.copy() for c in g["x0"]] if g["x0"] is not None else None
        if x0 is not None:
            x0 = O.tt_rank_retraction(x0, [len(x0)] * (len(x0) - 1))
        return x0

    kw = dict(r_max=g["rank_restriction"], eps=g["eps"], nswp=g["inner_m"], kick_rank=2, amen=True)
    out = {"file": os.path.basename(path), "d": g["d"], "ineq": g["ineq"], "native": native}
    dev = (NativeBlockAmen if native else DeviceBlockAmen)(g["A"], g["aliases"], g["transposes"], g["b"], g["ineq"], rt=rt)
    x0 = prep()
    rt.sync()
    t0 = time.perf_counter()
    xd, resd = dev.solve(g["termination_tol"], x0=x0, **kw)
    rt.sync()
    out.update(t_dev=time.perf_counter() - t0, res_dev=resd, sweeps_dev=dev.sweeps, ranks_dev=dev.ranks,
               solves_dev=len(dev.trace))
    if use_oracle:
        trace = []
        x0 = prep()
        t0 = time.perf_counter()
        xo, reso, sto = O.tt_block_amen(bm, bv, g["termination_tol"], x0=x0, local_solver=ls, trace=trace, **kw)
        out.update(t_oracle=time.perf_counter() - t0, res_oracle=reso, sweeps_oracle=sto.sweeps,
                   ranks_oracle=O.tt_ranks(xo), solves_oracle=len(trace))
        nn = block_inner(bm, xo)
        diff2 = nn - 2 * block_inner(bm, xd, xo) + block_inner(bm, xd)
        out["sol_rel_diff"] = float(np.sqrt(max(diff2, 0.0) / nn))
        if len(trace) == len(dev.trace):
            tro = np.array([[t["res_old"], t["res_new"]] for t in trace])
            trd = np.array([[t[2], t[3]] for t in dev.trace])
            out["trace_absdiff"] = float(np.max(np.abs(tro - trd)))
    return out
