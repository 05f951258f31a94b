"""Stand-in for opt_einsum 3.4.0 (absent offline); harness-only, used to run the
reference under /root/reference for golden-vector generation.  np.einsum with a
greedy path is numerically equivalent to ~1e-15 (different pairwise order)."""
import numpy as np


def contract(eq, *ops, optimize=None, **kw):
    return np.einsum(eq, *ops, optimize="greedy")


class _Expr:
    def __init__(self, eq, shapes):
        self.eq = eq
        dummies = [np.empty(s) for s in shapes]
        self.path = np.einsum_path(eq, *dummies, optimize="greedy")[0]

    def __call__(self, *ops, **kw):
        return np.einsum(self.eq, *ops, optimize=self.path)


def contract_expression(eq, *shapes, optimize="greedy", **kw):
    return _Expr(eq, shapes)
