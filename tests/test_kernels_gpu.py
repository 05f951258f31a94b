"""GPU tier: product kernels through the C ABI on a B200 vs the oracle / reference fixtures."""
import numpy as np
import pytest

import rt_util
import kernel_cases as KC

pytestmark = pytest.mark.gpu
CASES = ["eq_small", "ineq_small", "eq_mid"]


@pytest.fixture(scope="module")
def rt():
    return rt_util.cuda_runtime()


@pytest.mark.parametrize("case", CASES)
def test_block_matvec(rt, case):
    KC.assert_small(KC.case_block_matvec(rt, case))


@pytest.mark.parametrize("case", CASES)
def test_phi(rt, case):
    KC.assert_small(KC.case_phi(rt, case))


@pytest.mark.parametrize("case", CASES)
def test_phi_grouped_gemm_path(rt, case):
    KC.assert_small(KC.case_phi_big(rt, case))


@pytest.mark.parametrize("l,L,r,R,s", [(55, 55, 55, 55, 5), (2, 44, 29, 44, 10), (130, 97, 130, 97, 17), (128, 128, 128, 128, 32)])
def test_phi_grouped_gemm_vs_oracle(rt, l, L, r, R, s):
    """K2 on large / ragged interfaces; (128, 128, s = 32) is a shape the fused kernel cannot hold in shared memory"""
    ranks = {(0, 0): (2, 3), (0, 1): (s, s - 1), (2, 1): (s, s), (2, 2): (s - 1, s)}
    KC.assert_small(KC.case_phi_big(rt, shape=(l, L, r, R, ranks)))


@pytest.mark.parametrize("case", CASES)
def test_rhs(rt, case):
    KC.assert_small(KC.case_rhs(rt, case))


@pytest.mark.parametrize("case", CASES)
def test_diag_dense(rt, case):
    KC.assert_small(KC.case_diag_dense(rt, case))


def test_gemm(rt):
    KC.assert_small(KC.case_gemm(rt))


@pytest.mark.parametrize("case", CASES)
def test_block_matvec_grouped_gemm_path(rt, case):
    """large-rank path of K1 (csrc/cgemm.cu) forced onto the fixture shapes: compressed variants, batch, sub, norms"""
    KC.assert_small(KC.case_block_matvec_big(rt, case))


@pytest.mark.parametrize("ksplit", [0, 5])
@pytest.mark.parametrize("r,R,nb,s", [(55, 55, 3, 5), (29, 44, 4, 10), (64, 64, 3, 8), (130, 97, 3, 17), (128, 128, 4, 16),
                                      (96, 130, 3, 9), (66, 34, 4, 6)])
def test_block_matvec_grouped_gemm_vs_einsum(rt, r, R, nb, s, ksplit):
    """128x128 / 64x64 DMMA tile kernels on ragged and on full shapes vs einsum of the reference equation; also
    checks the fused small-block kernel against it where that one still fits shared memory"""
    ranks = {(0, 0): (2, 3), (0, 1): (s, s - 1), (1, 2): (1, 1), (2, 1): (s, s), (2, 2): (s - 1, s)}
    if nb == 4:
        ranks.update({(3, 1): (1, 1), (3, 3): (2, 2)})
    KC.assert_small(KC.case_block_matvec_big(rt, shape=(r, R, nb, ranks), ksplit=ksplit))


@pytest.mark.parametrize("r,R,s", [(55, 55, 5), (29, 44, 10), (64, 64, 8)])
def test_block_matvec_large_vs_oracle(rt, r, R, s):
    """Large-regime shapes (SURVEY 8a''): maxcut_13 r2 and graphm_3 r2 local blocks."""
    import tt_oracle as O
    from ttipm_b200 import kernels as K
    rng = np.random.default_rng(r * 1000 + R)
    keys = {(0, 0): (2, 2), (0, 1): (1, 1), (1, 2): (1, 1), (2, 1): (s, s - 1), (2, 2): (s - 1, s)}
    A = {k: rng.standard_normal((a, 4, 4, b)) for k, (a, b) in keys.items()}
    P1 = {k: rng.standard_normal((r, a, r)) for k, (a, b) in keys.items()}
    P2 = {k: rng.standard_normal((R, b, R)) for k, (a, b) in keys.items()}
    x = rng.standard_normal((r, 3, 4, R))
    bm = O.BlockMatrix({k: [v] for k, v in A.items()}, transposes={(0, 1): (1, 0)})
    want = O.block_local_product(bm, 0, P1, P2, x)
    c = dict(A=A, transposes={(0, 1): (1, 0)}, aliases={})
    y = K.block_matvec(KC.full_terms(rt, c, P1, P2, False, False), rt.to_device(x), 3, (r, R), rt=rt)
    assert KC.rel(rt.to_host(y), want) < KC.TOL


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("grid", [0, 1, 7])
def test_reduced_matvec(rt, case, grid):
    KC.assert_small(KC.case_reduced_matvec(rt, case, grid_hint=grid))


@pytest.mark.parametrize("case,grid,restart,shift", [("eq_small", 1, None, 8.0), ("eq_small", 4, 12, 8.0),
                                                     ("ineq_small", 3, 16, 8.0), ("eq_mid", 0, None, 20.0),
                                                     ("eq_mid", 16, 30, 20.0), ("eq_mid", 148, None, 20.0)])
def test_lgmres(rt, case, grid, restart, shift):
    errs, meta = KC.case_lgmres(rt, case, grid_hint=grid, restart=restart, shift=shift)
    print(meta)
    KC.assert_small(errs, tol=1e-8)


def test_qr_svd(rt):
    KC.assert_small(KC.case_qr_svd(rt, shapes=((7, 5), (5, 7), (6, 6), (12, 3), (3, 12), (1, 4), (4, 1), (20, 12),
                                               (88, 66), (220, 165), (165, 220), (40, 300))), tol=1e-11)


def test_svd_tall_single_qr_form(rt):
    """ttipm_linalg_tall_triple_qr(0): the single-QR form of the tall SVD (K x M accumulator) stays available"""
    old = rt.lib.ttipm_linalg_tall_triple_qr(0)
    try:
        KC.assert_small(KC.case_qr_svd(rt, shapes=((20, 12), (88, 66), (220, 165), (440, 330))), tol=1e-11)
        KC.assert_small(KC.case_qr_svd(rt, shapes=((20, 12), (88, 66)), coop_min_dim=1, graded=True), tol=1e-11)
    finally:
        rt.lib.ttipm_linalg_tall_triple_qr(old)


def test_svd_noise_plateau(rt):
    """strongly graded unfoldings with a rounding-noise plateau (the sweep's real spectrum): rows below eps * ||R||_F are
    left alone by the Jacobi iteration; U stays orthonormal, U W = A, significant singular values agree with LAPACK"""
    KC.assert_small(KC.case_qr_svd(rt, shapes=((440, 330), (165, 220), (400, 156), (88, 81), (20, 12)), graded="plateau"), tol=1e-11)
    KC.assert_small(KC.case_qr_svd(rt, shapes=((88, 81), (20, 12)), coop_min_dim=1, graded="plateau", noise_floor=1.0), tol=1e-11)


def test_qr_svd_tall_panel_in_workspace(rt):
    """unfoldings with more rows than a shared-memory reflector panel holds keep the panel in the workspace"""
    KC.assert_small(KC.case_qr_svd(rt, shapes=((4100, 5), (4, 3900), (6000, 40))), tol=1e-11)


def test_qr_svd_cooperative(rt):
    """multi-CTA panel QR + block Jacobi: forced onto small shapes, default dispatch at the AMEn truncation sizes
    (maxcut_13 rank 2: (4 R) x (3 r) up to ~440 x 330), graded spectra, > 512 rows (memory-resident reflector path)"""
    small = ((7, 5), (5, 7), (6, 6), (12, 3), (3, 12), (1, 4), (4, 1), (20, 12), (37, 23), (19, 40), (24, 24))
    KC.assert_small(KC.case_qr_svd(rt, shapes=small, coop_min_dim=1), tol=1e-12)
    big = ((88, 66), (220, 165), (165, 220), (440, 330), (330, 440), (400, 400), (600, 200), (64, 900))
    KC.assert_small(KC.case_qr_svd(rt, shapes=big), tol=1e-11)
    KC.assert_small(KC.case_qr_svd(rt, shapes=((440, 330), (165, 220), (600, 200)), graded=True), tol=1e-11)


def test_qr_svd_beyond_one_cluster(rt):
    """unfoldings with >= 32 block pairs per Jacobi round (the rank-exploded intermediates of the zip-up products at
    maxcut_13) leave the 16-CTA cluster for a cooperative grid of up to 64 CTAs; both forms on the same matrix"""
    shapes = ((700, 640), (1200, 560))
    KC.assert_small(KC.case_qr_svd(rt, shapes=shapes), tol=1e-10)
    old = rt.lib.ttipm_linalg_use_cluster(0)
    try:
        KC.assert_small(KC.case_qr_svd(rt, shapes=((220, 165), (440, 330))), tol=1e-11)      # cooperative grid below 16 CTAs too
    finally:
        rt.lib.ttipm_linalg_use_cluster(old)


@pytest.mark.parametrize("r,R,s,ineq", [(3, 2, 36, True), (42, 21, 36, True), (16, 16, 48, False)])
def test_large_operator_rank(rt, r, R, s, ineq):
    """graphm_3 rank 2 reaches operator ranks of 36 at (r, R) = (42, 21): stage 2 without the staged operator core"""
    KC.assert_small(KC.case_large_operator_rank(rt, r=r, R=R, s=s, ineq=ineq))


@pytest.mark.parametrize("l,R,s", [(129, 2, 25), (129, 1, 25), (200, 2, 12)])
def test_block_matvec_thin_right_rank(rt, l, R, s):
    KC.assert_small(KC.case_block_matvec_thin_right_rank(rt, l=l, R=R, s=s))


def test_elementwise(rt):
    KC.assert_small(KC.case_elementwise(rt), tol=1e-13)


def test_tt_algebra(rt):
    KC.assert_small(KC.case_tt_algebra(rt))


def test_als_products_vs_reference(rt):
    """ALS-fitted TT products (reference src/tt_als.py:1502-1762) on the device against the reference's own
    outputs on the same inputs and NumPy seed: same ranks, same half sweeps as the oracle, dense product to 1e-10"""
    KC.assert_small(KC.case_als_products(rt))


def test_als_product_dispatch_large_ranks(rt):
    """size-independent property at ranks the CPU tier cannot afford: a mat-vec whose rank products (96) exceed the
    reference's threshold of 80 goes through the ALS fit (src/tt_als.py:1765-1768) and must reproduce the exact product
    (dense NumPy contraction of the same trains) to the fit tolerance, with bond ranks no larger than the exact ones"""
    import numpy as np
    from ttipm_b200 import als_product as AP, tt as T, use_runtime
    rng = np.random.default_rng(12)
    d, n = 8, 4
    ra, rd = [1, 4, 6, 6, 6, 6, 6, 4, 1], [1, 8, 16, 16, 16, 16, 16, 8, 1]
    A = [rng.standard_normal((ra[k], n, n, ra[k + 1])) / np.sqrt(ra[k] * ra[k + 1]) for k in range(d)]
    v = [rng.standard_normal((rd[k], n, rd[k + 1])) / np.sqrt(rd[k] * rd[k + 1]) for k in range(d)]
    fits = AP.STATS["fits"]
    np.random.seed(21)
    with use_runtime(rt):
        out = T.tt_mat_vec_mul([c.copy() for c in A], [c.copy() for c in v], 1e-6, 1e-12)
    assert AP.STATS["fits"] == fits + 1
    # exact product, core by core: (a b, m, A B) = sum_k A[a, m, k, A'] v[b, k, B]
    exact = [np.einsum("amkA,bkB->abmAB", a, b).reshape(a.shape[0] * b.shape[0], n, -1) for a, b in zip(A, v)]
    dense = lambda tt: KC._dense_tt(tt)
    want = dense(exact)
    got = dense(out)
    assert got.shape == want.shape
    assert np.linalg.norm(got - want) <= 1e-6 * np.linalg.norm(want)
    caps = [min(4 ** (k + 1), 4 ** (d - 1 - k), ra[k + 1] * rd[k + 1]) for k in range(d - 1)]
    assert all(c.shape[-1] <= cap for c, cap in zip(out[:-1], caps))
