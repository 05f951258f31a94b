// Large-rank path of K1 (local block matvec): the reference's 3-GEMM chain as three grouped contraction-GEMM
// launches (cgemm.cuh) with the two intermediates in L2-resident global scratch.
//   stage 1  T1[(rho,Lam),(nu,sig')] = sum_P      x[(rho,nu),P]          P2[Lam,sig',P]        per term
//   stage 2  T2[(sig,rho),(mu,Lam)]  = alpha sum   T1[(rho,Lam),(nu,sig')] A[sig,mu,nu,sig']     per term
//   stage 3  y[lam,(mu,Lam)]         = y_scale sum_terms sum_(sig,rho) P1[lam,sig,rho] T2[(sig,rho),(mu,Lam)] (+ sub)
// Same arithmetic and the same flop count as cy_src/lgmres_cy.pyx:146-153 / src/tt_als.py:193.
#include "cgemm.cuh"
#include "api_util.h"

namespace ttipm {

template <int BM, int BN, int WM, int WN>
static int cg_launch_cfg(CgParams& p, tt_stream_t st) {
    typedef CgCfg<BM, BN, WM, WN> Cfg;
    const int jobs = cg_plan_tiles(p, BM, BN);
    if (jobs == 0) return 0;
    return launch_kernel("k_cgemm", k_cgemm<BM, BN, WM, WN>, dim3(jobs * p.nbatch), dim3(Cfg::NT),
                         sizeof(double) * Cfg::SMEM_DOUBLES, st, false, p);
}

#ifdef TTIPM_EMU
static const int CG_TILE_M[3] = {16, 16, 16}, CG_TILE_N[3] = {16, 16, 16}, CG_PER_SM[3] = {1, 1, 1};
#else
static const int CG_TILE_M[3] = {128, 128, 64}, CG_TILE_N[3] = {128, 64, 64}, CG_PER_SM[3] = {1, 2, 3};
#endif
static int g_force_ksplit = 0;      // tests: force this split-K factor wherever K allows it
static int g_force_cfg = -1;        // tuning runs: force a tile shape (0 = 128x128, 1 = 128x64, 2 = 64x64)

static long cg_tiles(const CgParams& p, int cfg) {
    long t = 0;
    for (int q = 0; q < p.nprob; ++q)
        t += (long)((p.prob[q].M + CG_TILE_M[cfg] - 1) / CG_TILE_M[cfg]) * ((p.prob[q].N + CG_TILE_N[cfg] - 1) / CG_TILE_N[cfg]);
    return t * p.nbatch;
}

static int g_vec_enabled = 1;

long cg_plan(CgParams& p, int sms) {
    int nseg_total = 0;
    for (int q = 0; q < p.nprob; ++q) nseg_total = imax(nseg_total, p.prob[q].seg0 + p.prob[q].nseg);
    for (int e = 0; e < nseg_total; ++e) {
        if (g_vec_enabled) cg_mark_vec(p.seg[e]);
        else p.seg[e].a_vec = p.seg[e].b_vec = 0;
    }
    // largest tile shape that still gives every SM two tiles and does not waste most of a tile on a thin problem
    int min_m = 1 << 30, min_n = 1 << 30;
    for (int q = 0; q < p.nprob; ++q) {
        min_m = imin(min_m, p.prob[q].M);
        min_n = imin(min_n, p.prob[q].N);
    }
    // measured on B200 (profiles/cgemm_tile_shapes_r01.md): 128x64 tiles with two CTAs per SM beat 128x128 (one CTA
    // of 8 warps per SM leaves the DMMA pipe idle 50% of the time) and edge out 64x64 wherever both dimensions fill
    int cfg = 2;
    if (min_m >= 96 && min_n >= 48 && cg_tiles(p, 1) >= (long)sms) cfg = 1;
    if (g_force_cfg >= 0 && g_force_cfg <= 2) cfg = g_force_cfg;
    p.cfg = cfg;
    // split-K when the output tiles alone cannot fill the machine
    const long tiles = cg_tiles(p, cfg), conc = (long)sms * CG_PER_SM[cfg];
    long part = 0;
    for (int q = 0; q < p.nprob; ++q) {
        CgProb& pr = p.prob[q];
        int nsteps = 0;
        for (int e = 0; e < pr.nseg; ++e) nsteps += (p.seg[pr.seg0 + e].K + CG_BK - 1) / CG_BK;
        int ks = 1;
        if (tiles > 0 && tiles < conc + conc / 2) ks = (int)imin((int)((2 * conc + tiles - 1) / tiles), nsteps / 8);
        if (g_force_ksplit > 1) ks = imin(g_force_ksplit, nsteps);
        if (ks < 1) ks = 1;
        pr.steps_per = (nsteps + ks - 1) / ks;
        pr.ksplit = pr.steps_per > 0 ? (nsteps + pr.steps_per - 1) / pr.steps_per : 1;
        if (pr.ksplit > 1) part += (long)p.nbatch * pr.ksplit * pr.M * pr.N;
        pr.part = nullptr;
    }
    return part;
}

int cg_launch(CgParams& p, double* part, tt_stream_t st) {
    long o = 0;
    bool any_split = false;
    for (int q = 0; q < p.nprob; ++q) {
        CgProb& pr = p.prob[q];
        if (pr.ksplit > 1) {
            pr.part = part + o;
            o += (long)p.nbatch * pr.ksplit * pr.M * pr.N;
            any_split = true;
        }
    }
    int rc;
#ifdef TTIPM_EMU
    rc = cg_launch_cfg<16, 16, 1, 2>(p, st);       // tiny tiles: the emulator runs every thread as a fiber
#else
    rc = p.cfg == 0 ? cg_launch_cfg<128, 128, 4, 2>(p, st)
                    : (p.cfg == 1 ? cg_launch_cfg<128, 64, 4, 2>(p, st) : cg_launch_cfg<64, 64, 2, 2>(p, st));
#endif
    if (rc || !any_split) return rc;
    // second pass over the split problems only
    CgReduceParams rp;
    rp.nprob = 0; rp.nbatch = p.nbatch;
    int ctas = 0;
    for (int q = 0; q < p.nprob; ++q) {
        const CgProb& pr = p.prob[q];
        if (pr.ksplit <= 1) continue;
        rp.cta0[rp.nprob] = ctas;
        rp.prob[rp.nprob++] = pr;
        long want = ((long)pr.M * pr.N + 4 * block_threads() - 1) / (4 * block_threads());
        if (pr.sumsq) want = imin((int)want, (int)pr.sumsq_slots);       // one sum-of-squares slot per CTA
        ctas += (int)imax(1, (int)want);
    }
    rp.cta0[rp.nprob] = ctas;
    return launch_kernel("k_cg_reduce", k_cg_reduce, dim3(ctas * p.nbatch), dim3(block_threads()), 40 * 8, st, false, rp);
}

static double g_big_min_flops = 2.0e8;

static void* scratch_alloc(size_t bytes, tt_stream_t st) {
    void* ptr = nullptr;
#ifdef TTIPM_EMU
    (void)st;
    ptr = malloc(bytes);
#else
    pool_keep_freed_blocks();
    if (cudaMallocAsync(&ptr, bytes, st) != cudaSuccess) return nullptr;
#endif
    return ptr;
}
static void scratch_free(void* ptr, tt_stream_t st) {
#ifdef TTIPM_EMU
    (void)st;
    free(ptr);
#else
    cudaFreeAsync(ptr, st);
#endif
}

// plan, take split-K scratch from the stream-ordered allocator, launch
static int cg_run(CgParams& p, int sms, tt_stream_t st) {
    const long part = cg_plan(p, sms);
    double* buf = nullptr;
    if (part > 0) {
        buf = (double*)scratch_alloc(sizeof(double) * (size_t)part, st);
        if (!buf) return fail(5, "cgemm: cannot allocate %ld doubles of split-K scratch", part);
    }
    const int rc = cg_launch(p, buf, st);
    if (buf) scratch_free(buf, st);
    return rc;
}

bool mv_big_possible(int nterms, int l, int L, int nm, int nb_out, bool want_sumsq) {
    if (nterms < 1 || nterms > CG_MAX_PROBS || nb_out > CG_MAX_PROBS) return false;
    if (!want_sumsq) return true;        // the tile-count limit below only concerns the sum-of-squares slots of the epilogue
#ifdef TTIPM_EMU
    const int mt = 16;
#else
    const int mt = 64;
#endif
    return (long)((l + mt - 1) / mt) * ((nm * L + mt - 1) / mt) <= L;
}

bool mv_big_wanted(const MvTerm* t, int nterms, int l, int L, int r, int R, int nm, int nb_out, int nbatch) {
    if (nterms < 1 || nterms > CG_MAX_PROBS || nb_out > CG_MAX_PROBS) return false;
    double flops = 0.0;
    for (int q = 0; q < nterms; ++q)
        flops += 2.0 * r * nm * R * L * t[q].S + 2.0 * r * L * t[q].s * nm * nm * t[q].S + 2.0 * l * nm * L * r * t[q].s;
    if (flops * nbatch < g_big_min_flops) return false;
    // the per-tile sum-of-squares slots of the epilogue: at most L tiles per output block (smallest tile shape)
#ifdef TTIPM_EMU
    const int mt = 16;
#else
    const int mt = 64;
#endif
    if ((long)((l + mt - 1) / mt) * ((nm * L + mt - 1) / mt) > L) return false;
    return true;
}

int mv_big(const MvTerm* t, int nterms, int l, int L, int r, int R, int nm, int nb_out, const double* x, long x_bs,
           long x_rs, long x_ns, long x_batch, double* y, long y_bs, long y_rs, long y_ns, long y_batch, double y_scale,
           const double* sub, double sub_scale, double* sumsq, int nbatch, tt_stream_t st) {
    DevInfo di = dev_info();
    // scratch: per batch entry, per term T1 (r L nm S) and T2 (s r nm L)
    std::vector<long> o1(nterms), o2(nterms);
    long per_batch = 0;
    for (int q = 0; q < nterms; ++q) {
        o1[q] = per_batch; per_batch += (long)r * L * nm * t[q].S;
        o2[q] = per_batch; per_batch += (long)t[q].s * r * nm * L;
    }
    if (!fits_int(per_batch) || !fits_int((long)L * nm * 32)) return fail(1, "block_matvec: intermediates exceed int addressing");
    double* ws = (double*)scratch_alloc(sizeof(double) * (size_t)per_batch * nbatch, st);
    if (!ws) return fail(5, "block_matvec: cannot allocate %ld doubles of scratch", per_batch * nbatch);
    int rc = 0;
    {   // ---- stage 1 ----
        CgParams p;
        p.nprob = nterms; p.nbatch = nbatch;
        for (int q = 0; q < nterms; ++q) {
            const MvTerm& T = t[q];
            const int S = T.S;
            CgSeg& s = p.seg[q];
            s.A = x + (long)T.in_blk * x_bs; s.a_batch = x_batch;
            s.aM = AxisMap{nm, (int)x_rs, (int)x_ns}; s.aK = AxisMap{TT_AX_BIG, 0, 1}; s.a_kfast = 1;
            s.B = T.P2; s.b_batch = 0; s.bK = AxisMap{TT_AX_BIG, 0, T.p2s[2]}; s.K = R;
            CgProb& pr = p.prob[q];
            pr.M = r * nm; pr.N = L * S; pr.seg0 = q; pr.nseg = 1;
            pr.C = ws + o1[q]; pr.c_batch = per_batch; pr.c_scale = 1.0; pr.sub = nullptr; pr.sub_scale = 0.0; pr.sumsq = nullptr;
            pr.sumsq_batch = 0; pr.sumsq_slots = 0; pr.ksplit = 1;
            pr.cM = AxisMap{nm, L * nm * S, S};                    // m = (rho, nu) -> row (rho, .), column block nu
            if (T.p2s[0] < T.p2s[1] && T.p2s[2] != 1) {
                // Lam is the fastest axis of this P2 (the transposed alias): order n = (sig', Lam) so loads coalesce
                s.bN = AxisMap{L, T.p2s[1], T.p2s[0]}; s.b_kfast = 0;
                pr.cN = AxisMap{L, 1, nm * S};
            } else {
                s.bN = AxisMap{S, T.p2s[0], T.p2s[1]}; s.b_kfast = T.p2s[2] == 1 ? 1 : 0;
                pr.cN = AxisMap{S, nm * S, 1};
            }
        }
        rc = cg_run(p, di.sms, st);
    }
    if (!rc) {   // ---- stage 2 ----
        CgParams p;
        p.nprob = nterms; p.nbatch = nbatch;
        for (int q = 0; q < nterms; ++q) {
            const MvTerm& T = t[q];
            const int s_ = T.s, S = T.S;
            CgSeg& s = p.seg[q];
            s.A = ws + o1[q]; s.a_batch = per_batch; s.aM = AxisMap{TT_AX_BIG, 0, nm * S}; s.aK = AxisMap{TT_AX_BIG, 0, 1};
            s.a_kfast = 1;
            s.B = T.A; s.b_batch = 0; s.bK = AxisMap{S, T.as_[2], T.as_[3]}; s.bN = AxisMap{nm, T.as_[0], T.as_[1]};
            s.b_kfast = T.as_[3] == 1 ? 1 : 0; s.K = nm * S;
            CgProb& pr = p.prob[q];
            pr.M = r * L; pr.N = s_ * nm; pr.seg0 = q; pr.nseg = 1;
            pr.C = ws + o2[q]; pr.c_batch = per_batch; pr.c_scale = T.alpha; pr.sub = nullptr; pr.sub_scale = 0.0;
            pr.sumsq = nullptr; pr.sumsq_batch = 0; pr.sumsq_slots = 0; pr.ksplit = 1;
            pr.cM = AxisMap{L, nm * L, 1};                         // m = (rho, Lam)
            pr.cN = AxisMap{nm, r * nm * L, L};                    // n = (sig, mu)
        }
        rc = cg_run(p, di.sms, st);
    }
    if (!rc) {   // ---- stage 3 ----
        CgParams p;
        p.nprob = nb_out; p.nbatch = nbatch;
        int nseg = 0;
        for (int i = 0; i < nb_out; ++i) {
            CgProb& pr = p.prob[i];
            pr.M = l; pr.N = nm * L; pr.seg0 = nseg; pr.nseg = 0;
            for (int q = 0; q < nterms; ++q) {
                if (t[q].out_blk != i) continue;
                const MvTerm& T = t[q];
                CgSeg& s = p.seg[nseg++];
                s.A = T.P1; s.a_batch = 0; s.aM = AxisMap{TT_AX_BIG, 0, T.p1s[0]}; s.aK = AxisMap{r, T.p1s[1], T.p1s[2]};
                s.a_kfast = T.p1s[2] == 1 ? 1 : 0; s.K = T.s * r;
                s.B = ws + o2[q]; s.b_batch = per_batch; s.bK = AxisMap{TT_AX_BIG, 0, nm * L}; s.bN = AxisMap{TT_AX_BIG, 0, 1};
                s.b_kfast = 0;
                pr.nseg++;
            }
            pr.C = y + (long)i * y_bs; pr.c_batch = y_batch; pr.c_scale = y_scale;
            pr.sub = sub ? sub + (long)i * y_bs : nullptr; pr.sub_scale = sub_scale;
            pr.sumsq = sumsq ? sumsq + (long)i * L : nullptr; pr.sumsq_batch = (long)nb_out * L; pr.sumsq_slots = L;
            pr.cM = AxisMap{TT_AX_BIG, 0, (int)y_rs};
            pr.cN = AxisMap{L, (int)y_ns, 1};
        }
        rc = cg_run(p, di.sms, st);
    }
    scratch_free(ws, st);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// K2 (interface update, reference src/tt_als.py:252-257) at large ranks: the same three-launch scheme.
//   forward : T1[l,s,N,R'] = Phi[l,s,r] V[r,N,R'];  T2[l,M,S,R'] = A[s,M,N,S] T1[l,s,N,R'] (batched over l);
//             out[L',S,R'] = U[l,M,L'] T2[l,M,S,R']
//   backward: T1[L,S,r,N] = Phi[L,S,R] V[r,N,R];    T2[L,s,M,r]  = A[s,M,N,S] T1[L,S,r,N]  (batched over L);
//             out[l,s,r]   = U[l,M,L] T2[L,s,M,r]
// ---------------------------------------------------------------------------------------------------------------
static void cg_clear(CgProb& pr) {
    pr.c_batch = 0; pr.c_scale = 1.0; pr.sub = nullptr; pr.sub_scale = 0.0; pr.sumsq = nullptr; pr.sumsq_batch = 0;
    pr.sumsq_slots = 0; pr.ksplit = 1; pr.nseg = 1;
}

bool phi_big_wanted(const PhiTermLite* t, int nterms, int forward, int ul, int uL, int vr, int vR, int nm) {
    if (nterms < 1 || nterms > CG_MAX_PROBS) return false;
    double flops = 0.0;
    for (int q = 0; q < nterms; ++q) {
        const double s = t[q].s, S = t[q].S;
        flops += forward ? 2.0 * ul * s * vr * nm * vR + 2.0 * ul * vR * s * nm * nm * S + 2.0 * ul * nm * uL * S * vR
                         : 2.0 * uL * S * vR * nm * vr + 2.0 * uL * vr * S * nm * nm * s + 2.0 * uL * nm * ul * s * vr;
    }
    return flops >= g_big_min_flops;
}

int phi_big(const PhiTermLite* t, int nterms, int forward, const double* U, int ul, int uL, const double* V, int vr, int vR,
            int nm, tt_stream_t st) {
    DevInfo di = dev_info();
    std::vector<long> o1(nterms), o2(nterms);
    long total = 0;
    for (int q = 0; q < nterms; ++q) {
        const long s = t[q].s, S = t[q].S;
        o1[q] = total; total += forward ? (long)ul * s * nm * vR : (long)uL * S * vr * nm;
        o2[q] = total; total += forward ? (long)ul * nm * S * vR : (long)uL * s * nm * vr;
    }
    if (!fits_int(total)) return fail(1, "phi_update: intermediates exceed int addressing");
    double* ws = (double*)scratch_alloc(sizeof(double) * (size_t)total, st);
    if (!ws) return fail(5, "phi_update: cannot allocate %ld doubles of scratch", total);
    int rc = 0;
    {   // ---- stage 1: Phi (rows (l,s) | (L,S)) times V ----
        CgParams p;
        p.nprob = nterms; p.nbatch = 1;
        for (int q = 0; q < nterms; ++q) {
            const int s = t[q].s, S = t[q].S;
            CgSeg& g = p.seg[q];
            CgProb& pr = p.prob[q];
            cg_clear(pr);
            pr.seg0 = q;
            g.A = t[q].Phi; g.a_batch = 0; g.b_batch = 0; g.B = V; g.a_kfast = 1;
            g.aK = AxisMap{TT_AX_BIG, 0, 1};
            if (forward) {
                g.aM = AxisMap{TT_AX_BIG, 0, vr}; g.K = vr;
                g.bK = AxisMap{TT_AX_BIG, 0, nm * vR}; g.bN = AxisMap{TT_AX_BIG, 0, 1}; g.b_kfast = 0;
                pr.M = ul * s; pr.N = nm * vR;
            } else {
                g.aM = AxisMap{TT_AX_BIG, 0, vR}; g.K = vR;
                g.bK = AxisMap{TT_AX_BIG, 0, 1}; g.bN = AxisMap{TT_AX_BIG, 0, vR}; g.b_kfast = 1;
                pr.M = uL * S; pr.N = vr * nm;
            }
            pr.C = ws + o1[q]; pr.cM = AxisMap{TT_AX_BIG, 0, pr.N}; pr.cN = AxisMap{TT_AX_BIG, 0, 1};
        }
        rc = cg_run(p, di.sms, st);
    }
    if (!rc) {   // ---- stage 2: operator core, batched over the leading interface index ----
        CgParams p;
        p.nprob = nterms; p.nbatch = forward ? ul : uL;
        for (int q = 0; q < nterms; ++q) {
            const int s = t[q].s, S = t[q].S;
            const int* as_ = t[q].as_;
            CgSeg& g = p.seg[q];
            CgProb& pr = p.prob[q];
            cg_clear(pr);
            pr.seg0 = q;
            g.A = t[q].A; g.a_batch = 0; g.B = ws + o1[q];
            if (forward) {
                // A(m = (M,S), k = (s,N)), B(k = (s,N), n = R') = T1[l][k][R']
                g.aM = AxisMap{S, as_[1], as_[3]}; g.aK = AxisMap{nm, as_[0], as_[2]}; g.a_kfast = 0; g.K = s * nm;
                g.bK = AxisMap{TT_AX_BIG, 0, vR}; g.bN = AxisMap{TT_AX_BIG, 0, 1}; g.b_kfast = 0; g.b_batch = (long)s * nm * vR;
                pr.M = nm * S; pr.N = vR; pr.c_batch = (long)nm * S * vR;
            } else {
                // A(m = (s,M), k = (S,N)), B(k = (S,N), n = r) = T1[L][S][r][N]
                g.aM = AxisMap{nm, as_[0], as_[1]}; g.aK = AxisMap{nm, as_[3], as_[2]}; g.a_kfast = 1; g.K = S * nm;
                g.bK = AxisMap{nm, vr * nm, 1}; g.bN = AxisMap{TT_AX_BIG, 0, nm}; g.b_kfast = 1; g.b_batch = (long)S * vr * nm;
                pr.M = s * nm; pr.N = vr; pr.c_batch = (long)s * nm * vr;
            }
            pr.C = ws + o2[q]; pr.cM = AxisMap{TT_AX_BIG, 0, pr.N}; pr.cN = AxisMap{TT_AX_BIG, 0, 1};
        }
        rc = cg_run(p, di.sms, st);
    }
    if (!rc) {   // ---- stage 3: the left core ----
        CgParams p;
        p.nprob = nterms; p.nbatch = 1;
        for (int q = 0; q < nterms; ++q) {
            const int s = t[q].s, S = t[q].S;
            CgSeg& g = p.seg[q];
            CgProb& pr = p.prob[q];
            cg_clear(pr);
            pr.seg0 = q;
            g.A = U; g.a_batch = 0; g.B = ws + o2[q]; g.b_batch = 0;
            if (forward) {
                // A(m = L', k = (l,M)) = U[k * uL + m];  B(k, n = (S,R')) = T2[k][n]
                g.aM = AxisMap{TT_AX_BIG, 0, 1}; g.aK = AxisMap{TT_AX_BIG, 0, uL}; g.a_kfast = 0; g.K = ul * nm;
                g.bK = AxisMap{TT_AX_BIG, 0, S * vR}; g.bN = AxisMap{TT_AX_BIG, 0, 1}; g.b_kfast = 0;
                pr.M = uL; pr.N = S * vR;
            } else {
                // A(m = l, k = (M,L)) = U[m * nm * uL + k];  B(k = (M,L), n = (s,r)) = T2[L][s][M][r]
                g.aM = AxisMap{TT_AX_BIG, 0, nm * uL}; g.aK = AxisMap{TT_AX_BIG, 0, 1}; g.a_kfast = 1; g.K = nm * uL;
                g.bK = AxisMap{uL, vr, s * nm * vr}; g.bN = AxisMap{vr, nm * vr, 1}; g.b_kfast = 0;
                pr.M = ul; pr.N = s * vr;
            }
            pr.C = t[q].out; pr.cM = AxisMap{TT_AX_BIG, 0, pr.N}; pr.cN = AxisMap{TT_AX_BIG, 0, 1};
        }
        rc = cg_run(p, di.sms, st);
    }
    scratch_free(ws, st);
    return rc;
}

}  // namespace ttipm

using namespace ttipm;

extern "C" int ttipm_cgemm_vector_loads(int on) {
    const int old = g_vec_enabled;
    if (on >= 0) g_vec_enabled = on ? 1 : 0;
    return old;
}

extern "C" int ttipm_cgemm_force_cfg(int cfg) {
    const int old = g_force_cfg;
    if (cfg >= -1) g_force_cfg = cfg;
    return old;
}

extern "C" int ttipm_cgemm_force_ksplit(int ksplit) {
    const int old = g_force_ksplit;
    if (ksplit >= 0) g_force_ksplit = ksplit;
    return old;
}

extern "C" double ttipm_matvec_big_min_flops(double min_flops) {
    const double old = g_big_min_flops;
    if (min_flops >= 0.0) g_big_min_flops = min_flops;
    return old;
}
