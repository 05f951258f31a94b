"""Readers for the fixtures under tests/golden/ (formats: oracle/ref_harness/make_golden.py)."""
import glob
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def get_tt(z, prefix):
    n = int(z[prefix + "/n"])
    return [z[f"{prefix}/{k}"] for k in range(n)]


def sub(z, prefix):
    """All arrays under 'prefix/' keyed by the remainder of the name."""
    pl = len(prefix) + 1
    return {k[pl:]: z[k] for k in z.files if k.startswith(prefix + "/")}


def keyed(z, prefix):
    """{(i,j): array} for names 'prefix/ij' ; {i: array} for names 'prefix/i'."""
    out = {}
    for k, v in sub(z, prefix).items():
        if "/" in k:
            continue
        out[(int(k[0]), int(k[1])) if len(k) == 2 else int(k)] = v
    return out


def amen_files(pattern="amen_*.npz"):
    return sorted(glob.glob(os.path.join(GOLD, pattern)))


def load_amen(path):
    """-> dict(A={(i,j):[cores]}, aliases, transposes, b={i:[cores]}, x0, rng_state, args..., out_x, out_res, trace)"""
    z = np.load(path)
    d = int(z["d"])
    A, b = {}, {}
    for name in z.files:
        parts = name.split("/")
        if parts[0] == "A":
            A.setdefault((int(parts[1][0]), int(parts[1][1])), [None] * d)[int(parts[2])] = z[name]
        elif parts[0] == "b":
            b.setdefault(int(parts[1]), [None] * d)[int(parts[2])] = z[name]
    x0 = [z[f"x0/{k}"] for k in range(d)] if "x0/0" in z.files else None
    out_x = [z[f"out/x/{k}"] for k in range(d)] if "out/x/0" in z.files else None
    al = {(int(r[0]), int(r[1])): (int(r[2]), int(r[3])) for r in z["aliases"]}
    tr = {(int(r[0]), int(r[1])): (int(r[2]), int(r[3])) for r in z["transposes"]}
    a = z["args"]
    rng_state = ("MT19937", z["rng_keys"], int(z["rng_pos"][0]), int(z["rng_pos"][1]), float(z["rng_gauss"]))
    return dict(A=A, aliases=al, transposes=tr, b=b, x0=x0, rng_state=rng_state, d=d,
                rank_restriction=int(a[0]), op_tol=float(a[1]), termination_tol=float(a[2]), eps=float(a[3]),
                num_restarts=int(a[4]), inner_m=int(a[5]), ineq=bool(a[6]), out_x=out_x,
                out_res=float(z["out_res"]) if "out_res" in z.files else None, trace=z["trace"],
                wall_s=float(z["wall_s"]), raised=bool(z["raised"]))
