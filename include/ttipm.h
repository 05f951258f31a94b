/* libttipm_b200 -- C ABI of the B200-native TT-IPM Newton-system hot path.
 *
 * The reference (FreditorK/Tensor-Train-Interior-Point-Method) has no FFI: the path sits
 * behind plain Python functions.  Each entry point below replaces the
 * reference routine cited next to it; the Python host mirror
 * (tensor-train-interior-point-method_b200/ttipm_b200, src/, cy_src/) binds them with ctypes.
 *
 * Conventions: all arrays are float64 in DEVICE memory, row-major unless strides are given;
 * strides are in elements; `stream` is a cudaStream_t (NULL = default stream); every call
 * is asynchronous on `stream`; return value 0 = ok, otherwise ttipm_last_error() describes
 * the failure.  No torch types cross this boundary.
 */
#ifndef TTIPM_H
#define TTIPM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TTIPM_MAX_TERMS 16
#define TTIPM_ABI_VERSION 1

int ttipm_abi_version(void);
const char* ttipm_last_error(void);
/* number of SMs / max opt-in shared memory per block of the current device (0 on failure) */
int ttipm_device_info(int* sm_count, int* smem_optin_bytes);
/* Programmatic dependent launch of the library's kernels (each kernel may be scheduled while its predecessor in the
 * stream still runs and waits in `griddepcontrol.wait`): 1 (default; TTIPM_PDL=0 in the environment disables) / 0.
 * Returns the previous setting; a negative argument only queries. */
int ttipm_use_pdl(int on);

/* One projected operator block  P1[l,s,r] * A[s,m,n,S] * P2[L,S,R]  of the local KKT system.
 * The strides address the LOGICAL axes, so transposed / permuted operands
 * (reference src/tt_als.py:196 'lsr,smnS,LSR,lmL->rnR', :208 'snmS', :221 'RSL', :234 'rsl')
 * are expressed by permuting strides instead of copying. */
typedef struct ttipm_term {
    const double* P1;
    const double* A;
    const double* P2;
    int64_t p1_strides[3]; /* (l, s, r) */
    int64_t a_strides[4];  /* (s, m, n, S) */
    int64_t p2_strides[3]; /* (L, S, R) */
    int32_t s, S;
    int32_t in_block, out_block;
    double alpha;
} ttipm_term;

/* K1 -- y[:, i] = y_scale * sum over terms with out_block == i of alpha * (P1 A P2) x[:, in_block]
 *                   (+ sub_scale * sub[:, i] if sub != NULL)
 * Replaces TTBlockMatrixView.block_local_product / compressed_ / lcompressed_ / rcompressed_
 * (reference src/tt_als.py:190-238).  x block j element (rho, nu, Rho) lives at
 * x[batch*x_batch_stride + j*x_block_stride + rho*x_row_stride + nu*x_mode_stride + Rho]; y likewise
 * with (l, L) -- both the (r, b, n, R) block-core layout and the (r, n, b, R) forward unfolding of the
 * sweep are expressible.  Output blocks without a term get y_scale*0 (+ sub).  sub has y's layout
 * (no batch stride): with y_scale = 1, sub_scale = -1 the call returns the local residual A x - rhs.
 * If `sumsq` != NULL it receives, per batch entry, nb_out*L partial sums of squares of the stored
 * values (sum them for the squared norm). */
int ttipm_block_matvec(const ttipm_term* terms, int nterms, int l, int L, int r, int R, int nmode, int nb_out,
                       const double* x, int64_t x_block_stride, int64_t x_row_stride, int64_t x_mode_stride,
                       int64_t x_batch_stride, double* y, int64_t y_block_stride, int64_t y_row_stride,
                       int64_t y_mode_stride, int64_t y_batch_stride, double y_scale, const double* sub,
                       double sub_scale, double* sumsq, int nbatch, void* stream);

/* Local blocks whose chain costs at least this many flops per call run as three grouped contraction-GEMM launches
 * (128x128 / 64x64 DMMA tiles over the whole machine, intermediates in L2-resident scratch from the stream-ordered
 * allocator) instead of the fused one-CTA-per-slab kernel.  Returns the previous threshold (default 2e8);
 * a negative argument only queries. */
double ttipm_matvec_big_min_flops(double min_flops);
/* Test hook of the grouped-GEMM path: force this split-K factor wherever K allows (0 = automatic: split only when the
 * output tiles alone cannot fill the machine).  Returns the previous value; a negative argument only queries. */
int ttipm_cgemm_force_ksplit(int ksplit);
/* Tuning hook: force the tile shape of the grouped-GEMM launches (0 = 128x128, 1 = 128x64, 2 = 64x64, -1 = automatic).
 * Returns the previous value; an argument below -1 only queries. */
int ttipm_cgemm_force_cfg(int cfg);
/* Tuning / test hook: allow (1, default) or forbid (0) the two-doubles-per-cp.async operand loads of the grouped-GEMM
 * path (used where the contiguous axis has unit stride and 16-byte aligned rows).  Returns the previous value; a
 * negative argument only queries. */
int ttipm_cgemm_vector_loads(int on);

/* K4 -- diag[l,m,L] = sum_s,S P1[l,s,l] A[s,m,m,S] P2[L,S,L]   (reference src/tt_ipm.py:191, :292);
 * if invert != 0 stores 1/diag (the inv_I of the Schur reduction). */
int ttipm_local_diag(const ttipm_term* term, int l, int L, int nmode, int invert, double* out, void* stream);

/* K4 -- dense local operator  out[(l,m,L),(r,n,R)] = sum_s,S P1[l,s,r] A[s,m,n,S] P2[L,S,R]
 * (reference src/tt_ipm.py:201-212, :301-315); out is (l*nmode*L) x (r*nmode*R) row-major. */
int ttipm_local_dense(const ttipm_term* term, int l, int L, int r, int R, int nmode, double* out, void* stream);

/* K2 -- interface updates for `nterms` stored blocks at once
 *   forward : out[L',S,R'] = sum Phi[l,s,r] U[l,M,L'] A[s,M,N,S] V[r,N,R']   (src/tt_als.py:256-257)
 *   backward: out[l,s,r]   = sum Phi[L,S,R] U[l,M,L]  A[s,M,N,S] V[r,N,R]    (src/tt_als.py:252-253)
 * U is (ul, nmode, uL) and V is (vr, nmode, vR), both contiguous; A is addressed through
 * a_strides (so the m<->n swapped cores of the transposed residual interfaces need no copy);
 * Phi and out are contiguous. */
typedef struct ttipm_phi_term {
    const double* Phi;
    const double* A;
    double* out;
    int64_t a_strides[4];
    int32_t s, S;
} ttipm_phi_term;
int ttipm_phi_update(const ttipm_phi_term* terms, int nterms, int forward, const double* U, int ul, int uL,
                     const double* V, int vr, int vR, int nmode, void* stream);

/* K3 -- right-hand-side contractions (reference src/tt_als.py:82, :260-265, src/tt_ipm.py:187-189)
 *   mode 0: out[r,n,R] = sum Xb1[b,r] B[b,n,B'] Xb2[B',R]      (projection; out has row stride out_row_stride)
 *   mode 1: out[B',R]  = sum Xb1[b,r] B[b,n,B'] core[r,n,R]    (forward interface)
 *   mode 2: out[b,r]   = sum Xb2[B',R] B[b,n,B'] core[r,n,R]   (backward interface)
 * In mode 0 a term with B == NULL is a "zero term": its out block (r x n x R, row stride out_row_stride) is cleared by
 * the kernel (blocks of the local right-hand side without a core; saves a memset node in the launch chain). */
typedef struct ttipm_rhs_term {
    const double* Xb1;  /* (b, r)  or NULL when unused */
    const double* B;    /* (b, n, B') */
    const double* Xb2;  /* (B', R) or NULL when unused */
    double* out;
    int32_t b, Bp;
} ttipm_rhs_term;
int ttipm_rhs_contract(const ttipm_rhs_term* terms, int nterms, int mode, const double* core, int r, int R,
                       int nmode, int64_t out_row_stride, void* stream);

/* Generic strided batched contraction  C[b](m,n) = alpha * sum_k A[b](m,k) B[b](k,n) + beta * C[b](m,n)
 * (the bond-absorption GEMMs K5 of SURVEY 8a', reference src/tt_als.py:360,369,382,465-509). */
int ttipm_gemm(int M, int N, int K, double alpha, const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs,
               const double* B, int64_t b_rs, int64_t b_cs, int64_t b_bs, double beta, double* C, int64_t c_rs,
               int64_t c_cs, int64_t c_bs, int nbatch, void* stream);

/* Planning only, no device access: grid (CTAs) and dynamic shared memory (bytes) ttipm_local_lgmres would launch with for a
 * local block (r, nmode, R) whose operator cores all have rank op_rank; 0, or the error ttipm_local_lgmres would return
 * (4: the block does not fit the kernel).  The CPU-tier tests pin the set-up's grid search with it. */
int ttipm_lgmres_plan(int ineq, int r, int R, int nmode, int op_rank, int restart, int* grid, int* smem_bytes);
/* Device-resident LGMRES(restart, augment) on the Schur-reduced local KKT operator
 *   eq   (ineq=0): [y; x]    -> [K00 y + K01 x ; K21 x - K22 (inv_I o K01^T y)]
 *   ineq (ineq=1): [y; x; t] -> [K00 y + K01 x ; K21 x - K22 (inv_I o K01^T y + t) ; K31 x + K33 t]
 * Replaces MatVecWrapper / IneqMatVecWrapper (reference cy_src/lgmres_cy.pyx:203-510) together with
 * LGMRESSolver = PETSc KSP lgmres (reference src/tt_ipm.py:101-162; options at :249-254, :362-367):
 * zero initial guess, no preconditioner, classical Gram-Schmidt, convergence ||r|| <= rtol*||rhs||,
 * at most max_it inner steps.  Vectors are block-major (nblk, r, n, R) with nblk = 2 or 3.
 * One persistent cooperative kernel runs the whole solve; `workspace` must hold at least
 * ttipm_lgmres_workspace(...) doubles.  apply_only != 0 computes x = Op(rhs) instead (the matvec of
 * the wrappers).  info (device, 6 doubles, may be NULL) = its, matvecs, reason, cycles, residual
 * estimate, grid size; reason: 1 rtol, 2 atol, 3 max_it, -1 dtol, -2 breakdown, -3 null pivot, -4 nan.
 * grid_hint = 0 lets the library choose the number of CTAs. */
int64_t ttipm_lgmres_workspace(int ineq, int r, int R, int nmode, int restart, int augment);
int ttipm_local_lgmres(int ineq, const ttipm_term* K00, const ttipm_term* K01, const ttipm_term* K21,
                       const ttipm_term* K22, const ttipm_term* K31, const ttipm_term* K33, const double* inv_I,
                       int r, int R, int nmode, const double* rhs, double* x, double* workspace, int64_t ws_doubles,
                       int restart, int augment, int max_it, double rtol, int apply_only, int grid_hint,
                       double* info, void* stream);

/* ---- dense factorisations of core unfoldings ------------------------------------------------- */
/* One kernel serves every shape (csrc/linalg.cu): a single matrix with min(M, N) >= the threshold below runs as one
 * multi-CTA launch (panel Householder QR + QR-preconditioned block one-sided Jacobi); smaller or batched
 * matrices use one CTA per matrix.  The setter returns the previous threshold (default 17); min_dim <= 0 only queries. */
int ttipm_linalg_coop_min_dim(int min_dim);
/* The multi-CTA launch is one thread-block cluster of up to 16 CTAs (hardware cluster barrier between phases): 1
 * (default) / 0 = cooperative launch with a software grid barrier.  Returns the previous setting; a negative argument
 * only queries. */
int ttipm_linalg_use_cluster(int on);

/* Threads per CTA of the QR / SVD kernel (256 or 512; tuning hook).  Returns the previous value; any other argument
 * only queries. */
int ttipm_linalg_threads(int threads);
/* Tall SVDs (M >= N) are preconditioned by three QR factorisations (A = Q1 R1, R1^T = Q2 R2, R2^T = Q3 R3; Jacobi on the
 * rows of R3 with a K x K accumulator): 1 (default) / 0 = single QR with the K x M accumulator.  Returns the previous
 * setting; a negative argument only queries. */
int ttipm_linalg_tall_triple_qr(int on);
/* tuning: smallest number of rows per Jacobi block (even; default 8); returns the previous value */
int ttipm_linalg_block_rows(int nb);
/* Rows of the Jacobi iteration whose norm is below factor * eps * ||R||_F are treated as numerically zero and left
 * alone (fewer sweeps on strongly graded unfoldings; singular values below that level then carry an absolute error of
 * that size, U stays orthonormal and U W = A).  Default 0 = off: every row pair is orthogonalised to relative accuracy,
 * which is what the rank decisions of the AMEn sweep were validated with.  Returns the previous factor; a negative
 * argument only queries. */
double ttipm_linalg_noise_floor(double factor);
/* The Jacobi iteration stops after a sweep whose rotations were all "small" (inner product below 1e-11 |a| |b| and sine
 * below 1e-6): such a sweep leaves every pair orthogonal to ~1e-15, so the all-skip sweep that would only confirm
 * convergence is not run (one sweep of ~9 on the AMEn unfoldings).  1 (default) / 0 = always run the confirming sweep.
 * Returns the previous setting; a negative argument only queries. */
int ttipm_linalg_early_exit(int on);

/* Householder QR, A (M x N, strided) = Q (M x K) R (K x N), K = min(M, N); Q, R contiguous row-major.
 * Replaces scipy.linalg.qr(mode="economic") at reference cy_src/tt_ops_cy.pyx:147, src/tt_als.py:358, :482.
 * workspace: ttipm_qr_workspace() doubles (required). */
int64_t ttipm_qr_workspace(int M, int N, int nbatch);
int ttipm_qr(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int M, int N, double* Q, double* R,
             double* workspace, int nbatch, void* stream);

/* "Left" SVD by QR-preconditioned one-sided Jacobi: A (M x N, strided) -> U (M x K), S (K, descending),
 * Wt = S * V^T (K x N).  Replaces scipy.linalg.svd at reference cy_src/tt_ops_cy.pyx:205,:291,:357,:404,:418 and
 * src/tt_als.py:270,:331,:457 -- every call site there only consumes U, s and s*Vt.
 * workspace: ttipm_svd_workspace() doubles (required).
 * info (device int32[16] per batch entry, may be NULL): [0] = Jacobi sweeps, [1..3] = ns spent in QR / Q^T set-up /
 * Jacobi, [4] = CTAs per matrix, [5] = Jacobi block rows, [6..9] = Jacobi breakdown on CTA 0 (row loads, rotations,
 * row stores, grid barriers; ns), [10] = total ns. */
int64_t ttipm_svd_workspace(int M, int N, int nbatch);
int ttipm_svd_left(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int M, int N, double* U, double* S,
                   double* Wt, double* workspace, int32_t* info, int nbatch, void* stream);

/* ---- memory-bound helpers -------------------------------------------------------------------- */
/* out = permute(in, perm) for a 4-D contiguous tensor (out axis k = in axis perm[k]); optionally every
 * element is multiplied (scale_mode 1) or divided (2) by scale[index along OUTPUT axis scale_axis]
 * (the per-block equilibration `scales` of reference src/tt_als.py:321-326, :369, :444-451, :496). */
int ttipm_permute4(const double* in, const int32_t* in_dims, const int32_t* perm, double* out, const double* scale,
                   int scale_axis, int scale_mode, void* stream);
/* out[j] = max(||x[:, j, :]||_2, floor) for x (r, b, inner) contiguous (reference src/tt_als.py:321). */
int ttipm_block_norms(const double* x, int r, int b, int inner, double floor_, double* out, void* stream);
/* out = w .* (alpha a + beta b) + gamma c on (rows x inner) panels with row strides; b, c, w, out may be
 * NULL; sumsq (256 doubles, may be NULL) receives partial sums of squares of the result. */
int ttipm_ewise(int rows, int inner, double alpha, const double* a, int64_t a_rs, double beta, const double* b,
                int64_t b_rs, double gamma, const double* c, int64_t c_rs, const double* w, int64_t w_rs, double* out,
                int64_t out_rs, double* sumsq, void* stream);
/* Residual norms of the rank-truncation loop (reference src/tt_als.py:338-345, :466-471) for all candidate
 * ranks at once: out[j*256 + part] = partial || base - sum_{i>=j} Y_i ||^2, Y is (q x len). */
int ttipm_trunc_resnorms(const double* base, const double* Y, int q, int64_t len, double* out, void* stream);

/* ---- vector algebra of the host-driven LGMRES (large local blocks, ttipm_amen_host_krylov) ------------ */
/* Number of chunks (= CTAs, <= 128) the Krylov vector kernels split a vector of nv doubles into. */
int ttipm_cgs_parts(int64_t nv);
/* One pass of classical Gram-Schmidt against the nvec (<= 112) basis vectors V_i = V + i*ldv (the Arnoldi step of
 * PETSc's LGMRES behind reference cy_src/lgmres_cy.pyx:203-510): h_i = <V_i, w> (per-chunk partial sums added in a fixed
 * order), w <- w - sum_i h_i V_i in place, h_out[0..nvec) = h, sumsq[0..parts) = partial ||w||^2 of the updated vector.
 * partials: parts x nvec doubles of scratch. */
int ttipm_cgs_project(const double* V, int64_t ldv, int nvec, double* w, int64_t nv, double* partials, double* h_out,
                      double* sumsq, void* stream);
/* out = beta * base + scale * sum_i coefs[i] * vecs[i]  (nvec <= 112 device vectors of nv doubles; vecs / coefs are HOST
 * arrays passed by value to the kernel; base may be NULL): the solution update and the error-approximation vectors of
 * the LGMRES cycle. */
int ttipm_lincomb(int nvec, const double* const* vecs, const double* coefs, double scale, const double* base, double beta,
                  double* out, int64_t nv, void* stream);

/* ---- TT primitives of the IPM driver (memory bound) ---------------------------------------------- */
/* One core of tt_add (reference cy_src/tt_ops_cy.pyx:229-258): mode 0 = first core (concatenate on the
 * last axis), 1 = middle core (block diagonal), 2 = last core (concatenate on the first axis);
 * a is (ra, n, Ra), b is (rb, n, Rb), n = product of the mode sizes. */
int ttipm_block_diag(const double* a, const double* b, double* out, int ra, int Ra, int rb, int Rb, int n, int mode,
                     void* stream);
/* Embeddings of reference src/tt_ops.py:360-375 and :312-316: mode 0 = I (x) M, 1 = M (x) I
 * (in (r,2,2,R) -> out (r,4,4,R)); mode 2 = diagonal embedding (in (r,q,R) -> out (r,q,q,R)). */
int ttipm_embed(const double* in, double* out, int r, int R, int q, int mode, void* stream);
/* out (rows x cols, contiguous) = in (strided) scaled along axis 0 (rows) or 1 (columns) by s; divide != 0
 * divides instead (entries with s == 0 are left unscaled).  Used to move singular values between the
 * two factors of the zip-up products (reference cy_src/tt_ops_cy.pyx:407-409, :421-424). */
int ttipm_scale2d(const double* in, int64_t in_rs, int64_t in_cs, int rows, int cols, const double* s, int axis,
                  int divide, double* out, void* stream);

/* ---- native sweep driver: one object = one block AMEn solve of a KKT system -------------------------
 * Replaces tt_block_amen / _bck_sweep / _fwd_sweep (reference src/tt_als.py:277-670) together with the local
 * solvers _ipm_local_solver[_ineq] (reference src/tt_ipm.py:183-401).  The host passes the operator blocks,
 * right-hand sides, the warm start and the (host-RNG generated) residual train once; run() executes every
 * sweep from C++ on `stream` and only synchronises for the scalars the reference's control flow needs.
 * All `*_host` pointers are HOST memory (NumPy buffers); cores are row-major. */
typedef struct ttipm_amen ttipm_amen;
ttipm_amen* ttipm_amen_create(int d, int block_size, int ineq, void* stream);
void ttipm_amen_destroy(ttipm_amen* h);
int ttipm_amen_set_block(ttipm_amen* h, int i, int j, int k, const double* core_host, int s, int n, int S);
int ttipm_amen_add_alias(ttipm_amen* h, int i, int j, int p, int t, int is_transpose);
int ttipm_amen_set_rhs(ttipm_amen* h, int i, int k, const double* core_host, int rb, int n, int rb2);
/* which: 0 = solution train, 1 = residual train; nb = 0 for an ordinary core (r, n, R), else the block core */
int ttipm_amen_set_core(ttipm_amen* h, int which, int k, const double* core_host, int r, int nb, int n, int R);
int ttipm_amen_run(ttipm_amen* h, double term_tol, int r_max, double eps, int nswp, int kick_rank, int use_amen,
                   int direction, double* final_res, int* sweeps);
int ttipm_amen_core_shape(ttipm_amen* h, int k, int32_t* dims /* r, nb|0, n, R */);
int ttipm_amen_get_core(ttipm_amen* h, int k, double* dst_host);
/* Which local blocks run their Krylov solve HOST-DRIVEN (matvec through ttipm_block_matvec -- grouped contraction GEMMs
 * over the whole machine -- plus ttipm_cgs_project / ttipm_lincomb, Hessenberg matrix on the host) instead of inside the
 * persistent kernel of ttipm_local_lgmres: 0 = none, 1 = automatic (default: blocks whose first matvec intermediate
 * leaves fewer than 4 output columns per shared-memory tile, i.e. left rank x operator rank large -- graphm_3 from IPM
 * iteration 3 on), 2 = all (tests).  Same algorithm, same iteration counts.  Returns the previous mode; a negative
 * argument only queries. */
int ttipm_amen_host_krylov(int mode);
/* profile != 0: time every Krylov-kernel launch of the next run() with CUDA events (read back through stats) */
int ttipm_amen_set_profile(ttipm_amen* h, int on);
/* Per-category profile of the last profiled run (call ttipm_amen_stats first): out[27], 3 per category =
 * seconds between the CUDA events bracketing the category's launches, algorithmic work (flops; bytes for the
 * memory-bound helpers), launches.  Categories: 0 block matvec (K1), 1 interface update (K2), 2 rhs contraction (K3),
 * 3 bond GEMM (K5), 4 QR, 5 SVD, 6 memory-bound helpers, 7 dense Schur fallback (cuSOLVER/cuBLAS), 8 Krylov kernel. */
int ttipm_amen_profile(ttipm_amen* h, double* out);
/* stats[12] = sweeps, local solves, dense solves, Krylov solves, Krylov inner steps, Krylov matvecs, kernel
 * launches, host syncs, peak device bytes, trace rows, seconds inside the Krylov kernel (profiling runs only), its
 * algorithmic flops; trace: 5 doubles (swp, k, res_old, res_new, r*R) per solve */
int ttipm_amen_stats(ttipm_amen* h, double* stats, double* trace, int max_trace_rows);

/* ---- device-resident tensor trains: the TT algebra of cy_src/tt_ops_cy.pyx behind handles ------------
 * A ttipm_tt holds the cores of one train in HBM; operations run their whole per-core loop natively on `stream` and
 * synchronise only for the singular values the reference's rank rule needs.  dims = 4 ints per core: r, n1, n2
 * (0 for a 3-D core (r, n1, R)), R; host buffers hold the cores concatenated, row-major. */
typedef struct ttipm_tt ttipm_tt;
ttipm_tt* ttipm_tt_create(int d, void* stream);
void ttipm_tt_destroy(ttipm_tt* h);
int ttipm_tt_length(const ttipm_tt* h);
int ttipm_tt_set_cores(ttipm_tt* h, const double* host, const int32_t* dims);   /* one H2D copy */
int ttipm_tt_shapes(const ttipm_tt* h, int32_t* dims);
int ttipm_tt_get_cores(ttipm_tt* h, double* host);                              /* one D2H copy */
ttipm_tt* ttipm_tt_clone(const ttipm_tt* src);                                   /* shares the device buffers */
/* tt_scale multiplies ONE core (cy_src/tt_ops_cy.pyx:96-114; the host draws the index and rounds alpha to float32) */
int ttipm_tt_scale_core(ttipm_tt* h, int k, double alpha);
/* tt_rl_orthogonalise (cy_src/tt_ops_cy.pyx:132-159), in place */
int ttipm_tt_rl_orthogonalise(ttipm_tt* h);
/* tt_rank_reduce (cy_src/tt_ops_cy.pyx:180-226) with the per-bond tolerance eps (the caller applies eps / sqrt(d - 1));
 * collect != 0: the tail-energy rule of tt_psd_rank_reduce / tt_mask_rank_reduce (:283-318), *dropped = discarded energy */
int ttipm_tt_round(ttipm_tt* h, double eps, int collect, double* dropped);
/* tt_add (cy_src/tt_ops_cy.pyx:244-258); NULL on failure */
ttipm_tt* ttipm_tt_add(const ttipm_tt* a, const ttipm_tt* b);
/* tt_inner_prod (cy_src/tt_ops_cy.pyx:506-520) */
int ttipm_tt_inner(ttipm_tt* a, const ttipm_tt* b, double* out);
/* Fused inner-product chain <a, b> of two trains given as device core pointers (host arrays of d pointers; core k of a
 * is (ra[k], nm[k], ra[k+1]) contiguous, of b (rb[k], nm[k], rb[k+1]); boundary ranks 1): ONE single-CTA launch that keeps
 * the running (R1 x R2) matrix and the intermediate of reference cy_src/tt_ops_cy.pyx:506-520 in shared memory, out = one
 * device double.  Returns 0, an error code (> 0), or -1 when the chain does not fit the kernel (more than 40 cores, or a
 * per-core intermediate above 2048 doubles, where two GEMM launches per core are faster: ttipm_tt_inner then runs those). */
int ttipm_tt_inner_chain(int d, const double* const* a, const double* const* b, const int32_t* ra, const int32_t* rb,
                         const int32_t* nm, double* out, void* stream);
/* zip-up products with swap_cores (cy_src/tt_ops_cy.pyx:393-502): kind 0 tt_fast_matrix_vec_mul(A, B),
 * 1 tt_fast_mat_mat_mul(A, B), 2 tt_fast_hadamard(A, B); eps as passed to the reference function; NULL on failure */
ttipm_tt* ttipm_tt_zipup(int kind, const ttipm_tt* A, const ttipm_tt* B, double eps);
/* tt_reshape without core merging (src/tt_ops.py:330-333): same buffers, modes (n1, n2) (n2 = 0: a single mode) */
ttipm_tt* ttipm_tt_reshape(const ttipm_tt* src, int n1, int n2);
/* tt_transpose of a train of 4-D cores (cy_src/tt_ops_cy.pyx:57-78), materialised */
ttipm_tt* ttipm_tt_transpose(const ttipm_tt* src);
/* per-core embeddings: kind 0 I (x) M (tt_IkronM), 1 M (x) I (tt_MkronI), 2 diagonal (tt_diag / tt_diag_op before
 * rounding) -- src/tt_ops.py:312-316, :360-375 */
ttipm_tt* ttipm_tt_embed(const ttipm_tt* src, int kind);
int ttipm_tt_counters(const ttipm_tt* h, int64_t* launches, int64_t* syncs);

/* ---- step-size eigen sweeps: local problems of tt_max_generalised_eigen / tt_min_eig ---------------
 * One- or two-site projection of a TT matrix onto the current interfaces,
 *   'lsr,smnk,kptS,LSR->lmpLrntR' (reference src/tt_als.py:952-959, :1305) or, with A2 == NULL,
 *   'lsr,smnS,LSR->lmLrnR' (:1037-1041, :1346),
 * as a dense (l n1 n2 L)^2 row-major matrix; symmetrise != 0 stores 0.5 (M + M^T) like the reference does. */
typedef struct ttipm_eig_op {
    const double* P1;      /* (l, s, l)       */
    const double* A1;      /* (s, n1, n1, k)  */
    const double* A2;      /* (k, n2, n2, S) or NULL */
    const double* P2;      /* (L, S, L)       */
    int64_t p1_strides[3], a1_strides[4], a2_strides[4], p2_strides[3];
    int32_t l, s, k, S, L, n1, n2;
} ttipm_eig_op;
int ttipm_eig_assemble(const ttipm_eig_op* op, int symmetrise, double* out, void* stream);
/* doubles of workspace for ttipm_eig_lanczos / ttipm_eig_gen_largest with a Lanczos basis of K vectors */
int64_t ttipm_eig_workspace(int m, int K);
/* Extreme eigenpair of cA * A + cD * D (dense symmetric m x m; D may be NULL): smallest algebraic (largest == 0) or
 * largest (largest != 0).  Replaces scipy.sparse.linalg.eigsh(M, k=1, which="SA", v0=...) / lobpcg of the reference
 * (src/tt_als.py:962-981, :1003-1006, :1307, :1320).  v0 (may be NULL) is the start vector; x receives the unit
 * eigenvector; out[10] = eigenvalue, residual ||M x - lambda x||, matvecs, converged flag, v0^T M v0 and
 * ||M v0 - (v0^T M v0) v0|| of the RAW start vector (the reference's `eig_val` / `old_res`, :992-993), cycles, ||v0||,
 * ||M v0 - lambda v0|| (tt_min_eig's old_res, :1315), reserved.
 * max_cycles == 0 only evaluates the start-vector quantities.  One persistent launch (a thread-block cluster). */
int ttipm_eig_lanczos(const double* A, double cA, const double* D, double cD, int m, const double* v0, int largest,
                      int K, int max_cycles, double tol, double* x, double* out, double* ws, void* stream);
/* Largest eigenpair of (-D) x = lambda A x, A positive definite (reference eigsh(-D, M=A, which="LA"),
 * src/tt_als.py:985, :1071); out[0] = lambda, out[3] = 0 if A is not positive definite (the reference's exception
 * branch).  Cholesky reduction through cuSOLVER / cuBLAS, then the kernel above. */
int ttipm_eig_gen_largest(const double* A, const double* D, int m, const double* v0, int K, int max_cycles, double tol,
                          double* x, double* out, double* ws, void* stream);
/* tuning / tests: force the number of CTAs of the Lanczos kernel's cluster (0 = automatic) */
int ttipm_eig_force_cluster(int ctas);

#ifdef __cplusplus
}
#endif
#endif
