"""TT algebra of the IPM driver on the device (placeholder header, filled in below)."""
import numpy as np


def tt_inner_prod_host(a, b):
    res = np.ones((1, 1))
    for c1, c2 in zip(a, b):
        t = np.tensordot(res, c1, axes=([0], [0]))
        ax = list(range(c1.ndim - 1))
        res = np.tensordot(t, c2, axes=(ax, ax))
    return float(res[0, 0])


def tt_scale(alpha, tt):
    idx = np.random.randint(0, len(tt))
    out = list(tt)
    out[idx] = float(np.float32(alpha)) * tt[idx]
    return out


def tt_normalise(tt, radius=1):
    return tt_scale(np.divide(int(radius), np.sqrt(tt_inner_prod_host(tt, tt))), tt)
