"""Step-size eigen sweeps: the Lanczos / assembly kernels against NumPy, and the device sweeps
`ttipm_b200.eigen.tt_max_generalised_eigen` / `tt_min_eig` against the reference's traced calls
(tests/golden/eigen_*.npz: returned step size to 1e-6 relative) and the oracle restatement.
CPU tier = the same kernel sources under the emulator on the small fixtures; GPU tier = everything through the C ABI."""
import os

import numpy as np
import pytest
import scipy.linalg as sla

import eigen_cases as EC
import rt_util
import tt_oracle as O


def _kernel_checks(rt, sizes, forced):
    from ttipm_b200 import kernels as K
    rng = np.random.default_rng(3)
    try:
        for m in sizes:
            B = rng.standard_normal((m, m))
            A = B + B.T
            D = rng.standard_normal((m, m))
            D = D + D.T
            M = 0.7 * A + D
            w, v = np.linalg.eigh(M)
            v0 = rng.standard_normal(m)
            for G in forced:
                rt.lib.ttipm_eig_force_cluster(G if m >= 8 else 0)
                x, out = K.eig_lanczos(rt.to_device(A), 0.7, rt.to_device(D), 1.0, v0=rt.to_device(v0), K=24, tol=1e-10, rt=rt)
                o, xx = rt.to_host(out), rt.to_host(x)
                assert o[3] == 1.0 and abs(o[0] - w[0]) <= 1e-10 * max(1.0, abs(w[0])), (m, G, o, w[0])
                assert o[1] <= 1e-9 and abs(abs(xx @ v[:, 0]) - 1.0) < 1e-8
                rq = v0 @ M @ v0
                assert abs(o[4] - rq) <= 1e-11 * max(1.0, abs(rq))
                assert abs(o[5] - np.linalg.norm(M @ v0 - rq * v0)) <= 1e-10 * np.linalg.norm(M @ v0)
                assert abs(o[8] - np.linalg.norm(M @ v0 - o[0] * v0)) <= 1e-10 * np.linalg.norm(M @ v0)
            rt.lib.ttipm_eig_force_cluster(0)
            x, out = K.eig_lanczos(rt.to_device(A), 0.7, rt.to_device(D), 1.0, v0=None, largest=True, K=24, tol=1e-10, rt=rt)
            assert abs(rt.to_host(out)[0] - w[-1]) <= 1e-10 * max(1.0, abs(w[-1]))
            # a start vector that is an exact eigenvector of a LARGER eigenvalue must not be accepted
            x, out = K.eig_lanczos(rt.to_device(M), 1.0, None, 0.0, v0=rt.to_device(v[:, min(2, m - 1)].copy()), K=24,
                                   tol=1e-10, rt=rt)
            assert abs(rt.to_host(out)[0] - w[0]) <= 1e-9 * max(1.0, abs(w[0])), (m, rt.to_host(out), w[:3])
            # pencil (-D, P), P positive definite
            P = B @ B.T + m * np.eye(m)
            wg = sla.eigh(-D, P, eigvals_only=True)
            x, out = K.eig_gen_largest(rt.to_device(P), rt.to_device(D), v0=rt.to_device(v0), K=24, tol=1e-10, rt=rt)
            o, xx = rt.to_host(out), rt.to_host(x)
            assert o[3] == 1.0 and abs(o[0] - wg[-1]) <= 1e-9 * max(1.0, abs(wg[-1]))
            assert np.linalg.norm(-D @ xx - o[0] * (P @ xx)) <= 1e-7 * np.linalg.norm(P @ xx) and abs(np.linalg.norm(xx) - 1) < 1e-12
            # not positive definite -> flagged, like the exception of the reference's eigsh(-D, M=A)
            x, out = K.eig_gen_largest(rt.to_device(A - 10 * m * np.eye(m)), rt.to_device(D), K=24, rt=rt)
            assert rt.to_host(out)[3] == 0.0
    finally:
        rt.lib.ttipm_eig_force_cluster(0)


def _assemble_checks(rt):
    from ttipm_b200 import kernels as K
    rng = np.random.default_rng(4)
    for (l, s, k, S, L, n1, n2) in ((3, 2, 3, 2, 4, 2, 2), (2, 3, 2, 1, 3, 4, 4), (1, 1, 2, 2, 1, 2, 2)):
        P1 = rng.standard_normal((l, s, l))
        P2 = rng.standard_normal((L, S, L))
        A1 = rng.standard_normal((s, n1, n1, k))
        A2 = rng.standard_normal((k, n2, n2, S))
        M = np.einsum("lsr,smnk,kptS,LSR->lmpLrntR", P1, A1, A2, P2)
        m = l * n1 * n2 * L
        M = M.reshape(m, m)
        got = rt.to_host(K.eig_assemble(rt.to_device(P1), rt.to_device(A1), rt.to_device(A2), rt.to_device(P2), rt=rt))
        assert np.abs(got - 0.5 * (M + M.T)).max() <= 1e-12 * np.abs(M).max()
        got = rt.to_host(K.eig_assemble(rt.to_device(P1), rt.to_device(A1), rt.to_device(A2), rt.to_device(P2),
                                        symmetrise=False, rt=rt))
        assert np.abs(got - M).max() <= 1e-12 * np.abs(M).max()
        A1s = rng.standard_normal((s, n1, n1, S))
        M1 = np.einsum("lsr,smnS,LSR->lmLrnR", P1, A1s, P2).reshape(l * n1 * L, l * n1 * L)
        got = rt.to_host(K.eig_assemble(rt.to_device(P1), rt.to_device(A1s), None, rt.to_device(P2), rt=rt))
        assert np.abs(got - 0.5 * (M1 + M1.T)).max() <= 1e-12 * np.abs(M1).max()


def _sweep_check(rt, path, with_oracle):
    from ttipm_b200 import eigen as E, tt as T, use_runtime
    g = EC.load(path)
    with use_runtime(rt):
        stats = {}
        out = EC.run(g, lambda *a, **k: E.tt_max_generalised_eigen(*a, _stats=stats, **k),
                     lambda *a, **k: E.tt_min_eig(*a, _stats=stats, **k))
        assert abs(T.tt_norm(out["x"]) - 1.0) < 1e-6
    if g["kind"] == 0:
        assert abs(out["step"] - g["scalar"]) <= 1e-6 * abs(g["scalar"]), (out["step"], g["scalar"], stats)
        assert stats["max_res"] <= g["tol"]
    else:
        rq = EC.rayleigh(O.tt_inner_prod, O.tt_fast_matrix_vec_mul, g["A"], out["x"])
        rq_ref = EC.rayleigh(O.tt_inner_prod, O.tt_fast_matrix_vec_mul, g["A"], g["out_x"])
        assert rq <= rq_ref + 1e-8 and abs(rq - rq_ref) <= 1e-6 * max(1.0, abs(rq_ref)), (rq, rq_ref, stats)
    if with_oracle:
        import eigen_oracle as EO
        ref = EC.run(g, EO.tt_max_generalised_eigen, EO.tt_min_eig)
        if g["kind"] == 0:
            assert abs(out["step"] - ref["step"]) <= 1e-6 * abs(ref["step"])


# ---- CPU tier (emulator) ---------------------------------------------------------------------------------------
def test_eig_kernels_emu():
    rt = rt_util.emu_runtime()
    _kernel_checks(rt, (1, 2, 5, 17, 40), (0, 2))
    _assemble_checks(rt)


@pytest.mark.parametrize("path", EC.SMALL, ids=os.path.basename)
def test_eigen_sweeps_emu(path):
    _sweep_check(rt_util.emu_runtime(), path, with_oracle=False)


# ---- GPU tier ----------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_eig_kernels_gpu():
    rt = rt_util.cuda_runtime()
    _kernel_checks(rt, (1, 2, 5, 17, 40, 90, 200, 513), (0, 1, 2, 8))
    _assemble_checks(rt)


@pytest.mark.gpu
@pytest.mark.parametrize("path", EC.FILES, ids=os.path.basename)
def test_eigen_sweeps_gpu(path):
    _sweep_check(rt_util.cuda_runtime(), path, with_oracle=True)
