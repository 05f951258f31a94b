import os
import sys
import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_oracle = os.path.abspath(os.path.join(_here, "..", "..", ".."))
if _oracle not in sys.path:
    sys.path.insert(0, _oracle)
from lgmres_ref import lgmres as _lgmres  # noqa: E402

COMM_WORLD = object()
_OPTIONS = {}
STATS = {"solves": 0, "its": 0, "matvecs": 0, "max_n": 0}


class Options:
    def setValue(self, key, value):
        _OPTIONS[key] = value


class Vec:
    def __init__(self):
        self._a = None

    def createWithArray(self, arr, comm=None):
        self._a = arr
        return self

    @property
    def array_r(self):
        return self._a

    @property
    def array_w(self):
        return self._a

    def destroy(self):
        self._a = None


class Mat:
    def createPython(self, shape, comm=None):
        self.shape = shape
        return self

    def setPythonContext(self, ctx):
        self.ctx = ctx

    def setUp(self):
        pass


class KSP:
    def create(self, comm=None):
        self.opts = {}
        return self

    def setType(self, t):
        assert t == "lgmres"

    def setFromOptions(self):
        self.opts = dict(_OPTIONS)

    def setOperators(self, A):
        self.A = A

    def solve(self, b, x):
        ctx = self.A.ctx
        xin, yout = Vec(), Vec()

        def mv(v):
            xin._a = np.ascontiguousarray(v)
            yout._a = np.empty_like(xin._a)
            ctx.mult(None, xin, yout)
            return yout._a

        res = _lgmres(mv, b.array_r,
                      rtol=float(self.opts.get("-ksp_rtol", 1e-5)),
                      max_it=int(self.opts.get("-ksp_max_it", 10000)),
                      restart=int(self.opts.get("-ksp_gmres_restart", 30)),
                      augment=int(self.opts.get("-ksp_lgmres_augment", 2)))
        STATS["solves"] += 1
        STATS["its"] += res.its
        STATS["matvecs"] += res.matvecs
        STATS["max_n"] = max(STATS["max_n"], b.array_r.size)
        x.array_w[:] = res.x

    def destroy(self):
        pass
