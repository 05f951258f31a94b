"""Host mirror of the reference's cy_src/lgmres_cy.pyx: the matrix-free Schur-reduced local KKT operators.

MatVecWrapper / IneqMatVecWrapper keep the reference's constructor signatures and flat block-major vector
convention (reference cy_src/lgmres_cy.pyx:216-232, :291-331, :392-414, :490-510); operands are uploaded
once in __init__ and matvec() runs the fused device kernel.  `solve()` is the device-resident replacement
for LGMRESSolver.solve_system (reference src/tt_ipm.py:136-154)."""
import numpy as np

from . import kernels as K
from .runtime import get_runtime


class BaseMatVec:
    def matvec(self, x_core):
        raise NotImplementedError("BaseMatVec.matvec must be implemented by subclass")


class _DeviceWrapper(BaseMatVec):
    _ineq = False
    _keys = []

    def _setup(self, P1, A, P2, inv_I, r, n, R):
        self.rt = get_runtime()
        up = self.rt.to_device
        self.r, self.n, self.R = int(r), int(n), int(R)
        self.nblk = 3 if self._ineq else 2
        self.op = K.ReducedOperator({k: up(v) for k, v in P1.items()}, {k: up(v) for k, v in A.items()},
                                    {k: up(v) for k, v in P2.items()}, up(np.asarray(inv_I).reshape(r, n, R)),
                                    self._ineq, rt=self.rt)

    def matvec(self, x_core):
        v = self.rt.to_device(np.asarray(x_core, dtype=np.float64).reshape(self.nblk, self.r, self.n, self.R))
        return self.rt.to_host(self.op.matvec(v)).reshape(-1)

    def solve(self, rhs, rtol=1e-5, max_iter=300, restart=100, outer_k=10):
        """Device LGMRES on this operator; returns the solution as a flat NumPy vector."""
        b = self.rt.to_device(np.asarray(rhs, dtype=np.float64).reshape(self.nblk, self.r, self.n, self.R))
        x, info = self.op.solve(b, int(restart), int(outer_k), max_it=int(max_iter), rtol=float(rtol))
        self.last_info = self.rt.to_host(info)
        return self.rt.to_host(x).reshape(-1)


class MatVecWrapper(_DeviceWrapper):
    def __init__(self, XAX_k_00, XAX_k_01, XAX_k_21, XAX_k_22, block_A_k_00, block_A_k_01, block_A_k_21, block_A_k_22,
                 XAX_kp1_00, XAX_kp1_01, XAX_kp1_21, XAX_kp1_22, inv_I, r, n, R):
        keys = [(0, 0), (0, 1), (2, 1), (2, 2)]
        self._setup(dict(zip(keys, (XAX_k_00, XAX_k_01, XAX_k_21, XAX_k_22))),
                    dict(zip(keys, (block_A_k_00, block_A_k_01, block_A_k_21, block_A_k_22))),
                    dict(zip(keys, (XAX_kp1_00, XAX_kp1_01, XAX_kp1_21, XAX_kp1_22))), inv_I, r, n, R)


class IneqMatVecWrapper(_DeviceWrapper):
    _ineq = True

    def __init__(self, XAX_k_00, XAX_k_01, XAX_k_21, XAX_k_22, XAX_k_31, XAX_k_33, block_A_k_00, block_A_k_01,
                 block_A_k_21, block_A_k_22, block_A_k_31, block_A_k_33, XAX_kp1_00, XAX_kp1_01, XAX_kp1_21,
                 XAX_kp1_22, XAX_kp1_31, XAX_kp1_33, inv_I, r, n, R):
        keys = [(0, 0), (0, 1), (2, 1), (2, 2), (3, 1), (3, 3)]
        self._setup(dict(zip(keys, (XAX_k_00, XAX_k_01, XAX_k_21, XAX_k_22, XAX_k_31, XAX_k_33))),
                    dict(zip(keys, (block_A_k_00, block_A_k_01, block_A_k_21, block_A_k_22, block_A_k_31,
                                    block_A_k_33))),
                    dict(zip(keys, (XAX_kp1_00, XAX_kp1_01, XAX_kp1_21, XAX_kp1_22, XAX_kp1_31, XAX_kp1_33))),
                    inv_I, r, n, R)
