"""Shared runner for the step-size eigen sweep fixtures (tests/golden/eigen_*.npz, traced from the reference IPM by
oracle/ref_harness/make_golden.py eigen): inputs, NumPy RNG state at entry, the reference's returned step size /
eigenvector train."""
import glob
import os

import numpy as np

import golden_io as G

FILES = sorted(glob.glob(os.path.join(G.GOLD, "eigen_*.npz")))
SMALL = [f for f in FILES if "maxcut_5" in f or f.endswith("s208_2.npz") or f.endswith("s208_11.npz")]


def load(path):
    z = np.load(path)
    g = dict(kind=int(z["kind"]), tol=float(z["tol"]), A=G.get_tt(z, "A"), out_x=G.get_tt(z, "out/x"),
             scalar=float(z["out_scalar"]), rng_state=("MT19937", z["rng_keys"], int(z["rng_pos"]), 0, 0.0))
    g["Delta"] = G.get_tt(z, "Delta") if "Delta/n" in z.files else None
    g["x0"] = G.get_tt(z, "x0") if "x0/n" in z.files else None
    return g


def rayleigh(inner, matvec, A, x):
    """x^T A x / x^T x with the given TT inner product / mat-vec (gauge-free comparison of eigenvector trains)."""
    return inner(x, matvec(A, [c.copy() for c in x], 1e-12)) / inner(x, x)


def run(g, gen_fn, min_fn):
    """Run one fixture with an implementation's (tt_max_generalised_eigen, tt_min_eig); returns a result dict."""
    np.random.set_state(g["rng_state"])
    x0 = [c.copy() for c in g["x0"]] if g["x0"] is not None else None
    A = [c.copy() for c in g["A"]]
    if g["kind"] == 0:
        D = [c.copy() for c in g["Delta"]]
        step, x = gen_fn(A, D, x0=x0, tol=g["tol"])
        return dict(step=float(step), x=x)
    x, _ = min_fn(A, x0=x0, tol=g["tol"])
    return dict(step=None, x=x)
