"""CPU tier: the REAL kernel sources, compiled with -DTTIPM_EMU (threads = OS threads), against
the oracle and the reference-generated fixtures.  Small shapes only."""
import pytest

import rt_util
import kernel_cases as KC

CASES = ["eq_small", "ineq_small"]


@pytest.fixture(scope="module")
def rt():
    return rt_util.emu_runtime()


@pytest.mark.parametrize("case", CASES)
def test_block_matvec(rt, case):
    KC.assert_small(KC.case_block_matvec(rt, case))


@pytest.mark.parametrize("case", CASES)
def test_block_matvec_grouped_gemm_path(rt, case):
    """large-rank path of K1 (csrc/cgemm.cu) forced onto the fixture shapes: compressed variants, batch, sub, norms"""
    KC.assert_small(KC.case_block_matvec_big(rt, case))


@pytest.mark.parametrize("ksplit", [0, 3])
def test_block_matvec_grouped_gemm_random(rt, ksplit):
    ranks = {(0, 0): (2, 3), (0, 1): (1, 2), (1, 2): (1, 1), (2, 1): (3, 2), (2, 2): (2, 2)}
    KC.assert_small(KC.case_block_matvec_big(rt, shape=(9, 7, 3, ranks), ksplit=ksplit))
    if ksplit:
        KC.assert_small(KC.case_block_matvec_big(rt, "eq_small", ksplit=ksplit))


@pytest.mark.parametrize("vec", [1, 0])
def test_block_matvec_grouped_gemm_vector_loads(rt, vec):
    """even ranks / aligned operands take the two-doubles-per-cp.async loaders; same numbers with them forbidden"""
    ranks = {(0, 0): (2, 2), (0, 1): (2, 4), (1, 2): (2, 2), (2, 1): (4, 2), (2, 2): (2, 2)}
    old = rt.lib.ttipm_cgemm_vector_loads(vec)
    try:
        KC.assert_small(KC.case_block_matvec_big(rt, shape=(8, 6, 3, ranks)))
        KC.assert_small(KC.case_block_matvec_big(rt, shape=(6, 10, 3, ranks), ksplit=2))
    finally:
        rt.lib.ttipm_cgemm_vector_loads(old)


@pytest.mark.parametrize("case", CASES)
def test_phi(rt, case):
    KC.assert_small(KC.case_phi(rt, case))


@pytest.mark.parametrize("case", CASES)
def test_phi_grouped_gemm_path(rt, case):
    KC.assert_small(KC.case_phi_big(rt, case))


def test_phi_grouped_gemm_random(rt):
    ranks = {(0, 0): (2, 3), (0, 1): (3, 2), (2, 2): (4, 2)}
    KC.assert_small(KC.case_phi_big(rt, shape=(5, 6, 7, 3, ranks)))
    KC.assert_small(KC.case_phi_big(rt, shape=(2, 9, 3, 8, ranks)))      # z-interface like: thin left core


@pytest.mark.parametrize("case", CASES)
def test_rhs(rt, case):
    KC.assert_small(KC.case_rhs(rt, case))


def test_diag_dense(rt):
    KC.assert_small(KC.case_diag_dense(rt, "eq_small"))


def test_gemm(rt):
    KC.assert_small(KC.case_gemm(rt))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("grid", [1, 2])
def test_reduced_matvec(rt, case, grid):
    KC.assert_small(KC.case_reduced_matvec(rt, case, grid_hint=grid))


@pytest.mark.parametrize("case,grid,restart", [("eq_small", 1, None), ("eq_small", 2, 12), ("ineq_small", 2, 16)])
def test_lgmres(rt, case, grid, restart):
    errs, meta = KC.case_lgmres(rt, case, grid_hint=grid, restart=restart)
    assert meta["its"] > 3, meta
    KC.assert_small(errs, tol=1e-8)


def test_qr_svd(rt):
    KC.assert_small(KC.case_qr_svd(rt), tol=1e-12)


def test_qr_svd_cooperative(rt):
    """the multi-CTA panel-QR / block-Jacobi kernel forced onto small shapes (2 emulated SMs)"""
    shapes = ((7, 5), (5, 7), (6, 6), (12, 3), (3, 12), (1, 4), (4, 1), (20, 12), (37, 23), (19, 40), (24, 24))
    KC.assert_small(KC.case_qr_svd(rt, shapes=shapes, coop_min_dim=1), tol=1e-12)
    KC.assert_small(KC.case_qr_svd(rt, shapes=((37, 23), (19, 40)), coop_min_dim=1, graded=True), tol=1e-12)


def test_svd_tall_single_qr_form(rt):
    """ttipm_linalg_tall_triple_qr(0): the single-QR form of the tall SVD (K x M accumulator) stays available"""
    old = rt.lib.ttipm_linalg_tall_triple_qr(0)
    try:
        KC.assert_small(KC.case_qr_svd(rt, shapes=((7, 5), (6, 6), (12, 3), (20, 12), (37, 23))), tol=1e-11)
        KC.assert_small(KC.case_qr_svd(rt, shapes=((7, 5), (6, 6), (12, 3), (20, 12), (37, 23)), coop_min_dim=1, graded=True), tol=1e-11)
    finally:
        rt.lib.ttipm_linalg_tall_triple_qr(old)


def test_svd_noise_plateau(rt):
    """strongly graded unfoldings with a rounding-noise plateau (the sweep's real spectrum): rows below eps * ||R||_F are
    left alone by the Jacobi iteration; U stays orthonormal, U W = A, significant singular values agree with LAPACK"""
    KC.assert_small(KC.case_qr_svd(rt, shapes=((37, 23), (19, 40), (30, 30)), graded="plateau"), tol=1e-11)
    KC.assert_small(KC.case_qr_svd(rt, shapes=((37, 23), (19, 40), (30, 30)), coop_min_dim=1, graded="plateau", noise_floor=1.0), tol=1e-11)


def test_qr_svd_tall_panel_in_workspace(rt):
    """unfoldings with more rows than a shared-memory reflector panel holds (> ~3500) keep the panel in the workspace"""
    KC.assert_small(KC.case_qr_svd(rt, shapes=((4100, 5), (4, 3900))), tol=1e-12)


def test_elementwise(rt):
    KC.assert_small(KC.case_elementwise(rt), tol=1e-13)


def test_tt_algebra(rt):
    KC.assert_small(KC.case_tt_algebra(rt))


def test_tt_products_above_als_threshold(rt):
    """rank products above the reference's ALS thresholds (80 mat-vec / 40 mat-mat, src/tt_als.py:1632,1766): the
    dispatchers take the ALS fit like the reference's; the dense result agrees with NumPy to the fit tolerance"""
    import numpy as np
    from ttipm_b200 import tt as T, use_runtime
    rng = np.random.default_rng(3)

    def dense(tt):
        d = tt[0]
        for c in tt[1:]:
            d = np.tensordot(d, c, axes=(-1, 0))
        return d.squeeze(0).squeeze(-1)

    rm, rv = 9, 10                                                # 90 > 80
    mat = [rng.standard_normal((1, 4, 4, rm)), rng.standard_normal((rm, 4, 4, 1))]
    vec = [rng.standard_normal((1, 4, rv)), rng.standard_normal((rv, 4, 1))]
    np.random.seed(5)
    with use_runtime(rt):
        out = T.tt_mat_vec_mul(mat, vec, 1e-10, 1e-12)
    want = np.einsum("imjn,mn->ij", dense(mat), dense(vec))      # cores (1, i, m, r), (r, j, n, 1)
    got = dense(out)
    assert np.linalg.norm(got - want) <= 1e-8 * np.linalg.norm(want)
    m2 = [rng.standard_normal((1, 4, 4, 7)), rng.standard_normal((7, 4, 4, 1))]      # 9 * 7 = 63 > 40
    with use_runtime(rt):
        out2 = T.tt_mat_mat_mul(mat, m2, 1e-10, 1e-12)
    A = dense(mat)          # (i, m, j, n): cores (1, i, m, r), (r, j, n, 1)
    B = dense(m2)
    want2 = np.einsum("imjn,mpnq->ipjq", A, B)
    assert np.linalg.norm(dense(out2) - want2) <= 1e-8 * np.linalg.norm(want2)


def test_als_products_vs_reference(rt):
    """ALS-fitted products against the reference's own outputs; the d = 6 case runs in the GPU tier"""
    KC.assert_small(KC.case_als_products(rt, KC.ALS_CASES[:3]))


def test_truncated_svd_and_kick_rank(rt):
    """reference src/tt_als.py:269-274 and cy_src/tt_ops_cy.pyx:538-579 through the device QR / SVD"""
    import numpy as np
    from ttipm_b200 import tt_als, tt_ops, use_runtime
    rng = np.random.default_rng(8)
    Mx = rng.standard_normal((12, 7))
    with use_runtime(rt):
        u, sv = tt_als.truncated_svd(Mx.copy(), 3)
        U, s, Vt = np.linalg.svd(Mx, full_matrices=False)
        assert u.shape == (12, 3) and sv.shape == (3, 7)
        assert np.linalg.norm(u @ sv - (U[:, :3] * s[:3]) @ Vt[:3]) <= 1e-12 * np.linalg.norm(Mx)
        np.random.seed(4)
        q, w, r = tt_ops.add_kick_rank(U[:, :3].copy(), (s[:3, None] * Vt[:3]).copy(), 2)
    np.random.seed(4)
    O = KC.O
    qo, wo, ro = O.add_kick_rank(U[:, :3].copy(), (s[:3, None] * Vt[:3]).copy(), 2)
    assert r == ro == 5 and np.linalg.norm(q.T @ q - np.eye(5)) <= 1e-13
    assert np.linalg.norm(q @ w - qo @ wo) <= 1e-12 * np.linalg.norm(qo @ wo)
    assert np.linalg.norm(q @ q.T - qo @ qo.T) <= 1e-12
    assert list(tt_ops.symmetric_powers_of_two(5)) == list(O.symmetric_powers_of_two(5)) == [2, 4, 8, 4, 2]
    assert list(tt_ops.symmetric_powers_of_two(4)) == [2, 4, 4, 2]


def test_large_operator_rank(rt):
    KC.assert_small(KC.case_large_operator_rank(rt, r=3, R=2, s=36, ineq=True))
    KC.assert_small(KC.case_large_operator_rank(rt, r=2, R=3, s=40, ineq=False))
