// Device-resident LGMRES on the Schur-reduced local KKT operator.
//
// Replaces, in one persistent cooperative kernel, what the reference does with PETSc KSP lgmres
// calling back into Python and cy_src/lgmres_cy.pyx for every matvec (reference
// src/tt_ipm.py:101-162, :249-254, :362-367; cy_src/lgmres_cy.pyx:291-331, :490-510):
//   operator (eq)  : [y; x]    -> [K00 y + K01 x ; K21 x - K22 (inv_I o K01^T y)]
//   operator (ineq): [y; x; t] -> [K00 y + K01 x ; K21 x - K22 (inv_I o K01^T y + t) ; K31 x + K33 t]
// Algorithm = PETSc's LGMRES(restart, augment) restated in oracle/lgmres_ref.py: zero initial guess,
// classical Gram-Schmidt (one pass), Givens-rotated Hessenberg, error-approximation augmentation.
// All Krylov vectors, the Hessenberg matrix and the control state live in device memory; the host is
// not involved until the solve has finished.  Vectors are block-major (blk, r, n, R).
#include "api_util.h"

namespace ttipm {

enum { LG_SOLVE = 0, LG_APPLY = 1 };
enum { R_NONE = 0, R_RTOL = 1, R_ATOL = 2, R_ITS = 3, R_DTOL = -1, R_BREAKDOWN = -2, R_NULL = -3, R_NAN = -4 };

struct LgParams {
    MvGeom g;
    int nblk, m, nv;          // unknown blocks (2|3), block length r*n*R, vector length
    int nA, nB;
    MvTerm tA[8];
    MvTerm tB[4];
    int nslotA;               // output slots of phase A
    int slotA[3];
    const double* inv_I;
    const double* b;
    double* x;
    double* V;                // (max_k + 2) x nv
    double* AUG;              // aug_dim x nv
    double* AAUG;             // aug_dim x nv
    double* upd;              // nv
    double* tmp;              // m
    double* part;             // G x (max_k + 3)
    double* hh;               // column-major (max_k + 2) x (max_k + 1)
    double* hes;
    double* small_;           // grs (max_k+2) | cc (max_k+1) | ss (max_k+1) | nrs (max_k+2) | avec (max_k+2)
    double* ctrl_d;           // [0]=res, [1]=grs at loc_it, [2]=final rnorm
    int* ctrl_i;              // [0]=null flag ; outputs [4]=its [5]=matvecs [6]=reason [7]=cycles
    unsigned* barrier;
    int max_k, aug_dim, max_it;
    double rtol, abstol, dtol, haptol;
    int mode;
    int ch;                   // chunk length per CTA
    int oW, oH, oScr;         // CGS shared-memory offsets (doubles)
    // single-CTA solves (tiny local systems, the small / medium regime): every operand of the reduced operator and the
    // vector being multiplied are staged in shared memory once per solve; together with the scalar small-block stages
    // of matvec.cuh a matvec then never waits for L2 (each of its ~15 dependent stages otherwise pays 1-4 L2 round trips)
    int stage;                // 1: stage operands (grid of one CTA and enough shared memory)
    int oStage;               // offset of the staging area (doubles): [term copies][operand data][vector nv]
    int oTiny;                // single-CTA matvec scratch: per term T1 (r R n S) and T2 (s r n R), then tmp (m)
    int tinyT1[12], tinyT2[12], tinyTmp;   // offsets inside that scratch, terms in the order tA[0..nA), tB[0..nB)
    int oHess;                // CTA 0: Givens cosines / sines, rotated rhs, working Hessenberg column, back-substitution
                              // vector (5 x (max_k + 2) doubles) and three control words in shared memory, outside the matvec / CGS scratch
};

#define LG_STAGE_TERM_DOUBLES ((int)((12 * sizeof(MvTerm) + 7) / 8))

struct LgCtx {
    const LgParams& p;
    double* smem;
    unsigned epoch;
    const MvTerm* tA;         // phase A / B term lists: the kernel parameters, or their shared-memory staged copies
    const MvTerm* tB;
    double* xs;               // staged input vector (nullptr when not staging)
    double* wcopy;            // single-CTA mode: the matvec also leaves its result here (the CGS work vector)
    TT_DEVM LgCtx(const LgParams& pp, double* s)
        : p(pp), smem(s), epoch(0), tA(pp.tA), tB(pp.tB), xs(nullptr), wcopy(nullptr) {}
    TT_DEVM void sync() { grid_sync(p.barrier, epoch); }
};

// one output slot of one phase for one L-tile
TT_DEV void lg_item(LgCtx& c, const MvTerm* terms, int nterms, int slot, int tile, const double* src, double* dst) {
    const LgParams& p = c.p;
    const MvGeom& g = p.g;
    const int L0 = tile * g.Lt, Ltc = imin(g.Lt, g.L - L0);
    mv_zero_tile(g, c.smem);
    for (int it = 0; it < nterms; ++it) {
        if (terms[it].out_blk != slot) continue;
        const double* xin = terms[it].in_blk == 3 ? p.tmp : src + (long)terms[it].in_blk * p.m;
        mv_accumulate_term(terms[it], xin, g.nm * g.R, g.R, g, L0, Ltc, c.smem);
    }
    const double* Ys = c.smem + g.oYs;
    const int cols = g.nm * Ltc;
    for (int i = threadIdx.x; i < g.l * cols; i += blockDim.x) {
        const int lam = i / cols, cc = i % cols, mu = cc / Ltc, lt = cc % Ltc;
        const int o = (lam * g.nm + mu) * g.L + L0 + lt;
        double v = Ys[lam * g.ldY + mu * Ltc + lt];
        if (slot == 3) {
            v *= p.inv_I[o];
            if (p.nblk == 3) v += src[2 * p.m + o];
            p.tmp[o] = v;
        } else {
            dst[(long)slot * p.m + o] = v;
        }
    }
    __syncthreads();
}

TT_DEV void lg_apply_tiny(LgCtx& c, const double* src, double* dst);

// dst = Op(src); src must be globally visible on entry, dst is globally visible on exit
TT_DEV void lg_apply(LgCtx& c, const double* src, double* dst) {
    const LgParams& p = c.p;
    const int nt = p.g.ntiles;
    if (c.xs) {                        // single CTA: the vector being multiplied goes to shared memory once
        __syncthreads();
        for (int e = threadIdx.x; e < p.nv; e += blockDim.x) c.xs[e] = src[e];
        __syncthreads();
        lg_apply_tiny(c, c.xs, dst);
        return;
    }
    for (int it = blockIdx.x; it < p.nslotA * nt; it += gridDim.x)
        lg_item(c, c.tA, p.nA, p.slotA[it / nt], it % nt, src, dst);
    c.sync();
    for (int it = blockIdx.x; it < nt; it += gridDim.x) lg_item(c, c.tB, p.nB, 1, it, src, dst);
    c.sync();
}

// Single-CTA matvec on staged operands: ncu showed the one-term-at-a-time form waiting at block barriers 58 % of the
// time (~30 barriers per matvec, each phase a short dependent chain on a few hundred outputs).  Here all terms of a
// phase advance together: stage 1 of every term, barrier, stage 2 of every term, barrier, stage 3 per output slot,
// barrier -- 6 barriers per matvec, each phase with 1000+ independent outputs for the 512 threads.
// Arithmetic per output = the same 3-GEMM chain as mv_accumulate_term_small (reference cy_src/lgmres_cy.pyx:146-153).
TT_DEV void lg_apply_tiny(LgCtx& c, const double* src, double* dst) {
    const LgParams& p = c.p;
    const int r = p.g.r, R = p.g.R, nm = p.g.nm, m = p.m, nth = blockDim.x, tid = threadIdx.x;
    double* tb = c.smem + p.oTiny;
    double* tmpS = tb + p.tinyTmp;
    for (int phase = 0; phase < 2; ++phase) {
        const MvTerm* T = phase ? c.tB : c.tA;
        const int nT = phase ? p.nB : p.nA, q0 = phase ? p.nA : 0;
        // ---- stage 1: T1[(rho, Lam), (nu, sig')] = sum_P x[rho, nu, P] P2[Lam, sig', P] ----
        int tot = 0;
        for (int q = 0; q < nT; ++q) tot += r * R * nm * T[q].S;
        for (int i = tid; i < tot; i += nth) {
            int q = 0, j = i;
            while (j >= r * R * nm * T[q].S) { j -= r * R * nm * T[q].S; ++q; }
            const int S = T[q].S;
            const int sp = j % S, nu = (j / S) % nm, Lam = (j / (S * nm)) % R, rho = j / (S * nm * R);
            const double* xin = T[q].in_blk == 3 ? tmpS : src + T[q].in_blk * m;
            const double* xp = xin + (rho * nm + nu) * R;
            const double* pp = T[q].P2 + (Lam * S + sp) * R;
            double acc = 0.0;
            for (int P = 0; P < R; ++P) acc += xp[P] * pp[P];
            tb[p.tinyT1[q0 + q] + j] = acc;               // j = ((rho * R + Lam) * nm + nu) * S + sp
        }
        __syncthreads();
        // ---- stage 2: T2[(sig, rho), (mu, Lam)] = sum_(nu, sig') T1[(rho, Lam), (nu, sig')] A[sig, mu, nu, sig'] ----
        tot = 0;
        for (int q = 0; q < nT; ++q) tot += T[q].s * r * nm * R;
        for (int i = tid; i < tot; i += nth) {
            int q = 0, j = i;
            while (j >= T[q].s * r * nm * R) { j -= T[q].s * r * nm * R; ++q; }
            const int S = T[q].S;
            const int Lam = j % R, mu = (j / R) % nm, rho = (j / (R * nm)) % r, sg = j / (R * nm * r);
            const double* tp = tb + p.tinyT1[q0 + q] + (rho * R + Lam) * nm * S;
            const double* ap = T[q].A + (sg * nm + mu) * nm * S;
            double acc = 0.0;
            for (int k = 0; k < nm * S; ++k) acc += tp[k] * ap[k];
            tb[p.tinyT2[q0 + q] + j] = acc;               // j = ((sig * r + rho) * nm + mu) * R + Lam
        }
        __syncthreads();
        // ---- stage 3 per output slot: y[lam, mu, Lam] = sum_terms alpha sum_(sig, rho) P1[lam, sig, rho] T2[...] ----
        const int nslot = phase ? 1 : p.nslotA;
        for (int i = tid; i < nslot * m; i += nth) {
            const int slot = phase ? 1 : p.slotA[i / m], o = i % m;
            const int c_ = o % (nm * R), lam = o / (nm * R);
            double v = 0.0;
            for (int q = 0; q < nT; ++q) {
                if (T[q].out_blk != slot) continue;
                const int sr = T[q].s * r;
                const double* pp = T[q].P1 + lam * sr;
                const double* t2 = tb + p.tinyT2[q0 + q] + c_;
                double acc = 0.0;
                for (int k = 0; k < sr; ++k) acc += pp[k] * t2[k * nm * R];
                v += T[q].alpha * acc;
            }
            if (slot == 3) {
                v *= p.inv_I[o];
                if (p.nblk == 3) v += src[2 * m + o];
                tmpS[o] = v;
            } else {
                dst[(long)slot * m + o] = v;
                if (c.wcopy) c.wcopy[slot * m + o] = v;
            }
        }
        __syncthreads();
    }
}

// copy the logical (d0, d1, d2[, d3]) tensor addressed by `strides` to a contiguous block of shared memory
TT_DEV void lg_stage_tensor(double* dst, const double* src, int d0, int d1, int d2, int d3, const int* st) {
    const int n = d0 * d1 * d2 * d3;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int t = i;
        const int i3 = t % d3; t /= d3;
        const int i2 = t % d2; t /= d2;
        const int i1 = t % d1; t /= d1;
        dst[i] = src[t * st[0] + i1 * st[1] + i2 * st[2] + (d3 > 1 ? i3 * st[3] : 0)];
    }
}

// stage every operand of the reduced operator (single-CTA solves); returns after a block barrier
TT_DEV void lg_stage_operands(LgCtx& c) {
    const LgParams& p = c.p;
    double* base = c.smem + p.oStage;
    MvTerm* tl = (MvTerm*)base;
    double* data = base + LG_STAGE_TERM_DOUBLES;
    const int r = p.g.r, R = p.g.R, nm = p.g.nm;
    int off = 0;
    for (int q = 0; q < p.nA + p.nB; ++q) {
        const MvTerm& t = q < p.nA ? p.tA[q] : p.tB[q - p.nA];
        const int n1 = r * t.s * r, nA = t.s * nm * nm * t.S, n2 = R * t.S * R;
        lg_stage_tensor(data + off, t.P1, r, t.s, r, 1, t.p1s);
        lg_stage_tensor(data + off + n1, t.A, t.s, nm, nm, t.S, t.as_);
        lg_stage_tensor(data + off + n1 + nA, t.P2, R, t.S, R, 1, t.p2s);
        if (threadIdx.x == 0) {
            MvTerm u = t;
            u.P1 = data + off; u.p1s[0] = t.s * r; u.p1s[1] = r; u.p1s[2] = 1;
            u.A = data + off + n1; u.as_[0] = nm * nm * t.S; u.as_[1] = nm * t.S; u.as_[2] = t.S; u.as_[3] = 1;
            u.P2 = data + off + n1 + nA; u.p2s[0] = t.S * R; u.p2s[1] = R; u.p2s[2] = 1;
            tl[q] = u;
        }
        off += n1 + nA + n2;
    }
    c.tA = tl;
    c.tB = tl + p.nA;
    c.xs = data + off;
    __syncthreads();
}

TT_DEV double lg_sum_partials(const LgParams& p, int col) {
    double t = 0.0;
    for (unsigned cta = 0; cta < gridDim.x; ++cta) t += ld_cg(&p.part[(long)cta * (p.max_k + 3) + col]);
    return t;
}

TT_DEV int lg_converged(const LgParams& p, int it, double rn, double& ttol, double& rnorm0) {
    if (it == 0) {
        ttol = fmax(p.rtol * rn, p.abstol);
        rnorm0 = rn;
    }
    if (!(rn == rn) || isinf(rn)) return R_NAN;
    if (rn <= ttol) return rn < p.abstol ? R_ATOL : R_RTOL;
    if (rn >= p.dtol * rnorm0) return R_DTOL;
    return R_NONE;
}

TT_GLOBAL void __launch_bounds__(512) k_lgmres(const LgParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    LgCtx c(p, smem);
    const int nv = p.nv, ldh = p.max_k + 2;
    const int e0 = imin(nv, (int)blockIdx.x * p.ch), e1 = imin(nv, e0 + p.ch);
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nth >> 5;
    double* wS = smem + p.oW;
    double* hS = smem + p.oH;
    double* scr = smem + p.oScr;
    double* mypart = p.part + (long)blockIdx.x * (p.max_k + 3);
    double* nrs = p.small_ + (p.max_k + 2) + 2 * (p.max_k + 1);
    double* avec = nrs + (p.max_k + 2);
    // the small dense recurrences of GMRES run on CTA 0 out of shared memory (one warp; the serial Givens chain on
    // lane 0): in global memory every dependent load of the chain costs an L2 round trip
    double* cc = smem + p.oHess;
    double* ss = cc + (p.max_k + 2);
    double* grs = ss + (p.max_k + 2);
    double* colS = grs + (p.max_k + 2);
    double* tS = colS + (p.max_k + 2);
    double* ctlS = tS + (p.max_k + 2);      // single-CTA mode: residual estimate, rotated rhs entry, null-pivot flag
    const bool lead = blockIdx.x == 0 && tid == 0;
    const bool lead_warp = blockIdx.x == 0 && wid == 0;
    if (p.stage) lg_stage_operands(c);
    const bool tiny = p.stage != 0;            // single CTA, operands staged
    if (tiny) c.wcopy = wS;

    if (p.mode == LG_APPLY) {
        lg_apply(c, p.b, p.x);
        return;
    }

    for (int e = e0 + tid; e < e1; e += nth) p.x[e] = 0.0;
    int its = 0, matvecs = 0, aug_ct = 0, reason = R_NONE, cycles = 0;
    int aug_order[16];
    for (int i = 0; i < 16; ++i) aug_order[i] = 0;
    double ttol = 0.0, rnorm0 = 0.0, res = 0.0;
    bool first = true;
    const int it_arnoldi = p.max_k - p.aug_dim;

    for (;;) {
        ++cycles;
        double* V0 = p.V;
        if (first) {
            for (int e = e0 + tid; e < e1; e += nth) V0[e] = p.b[e];
        } else {
            c.sync();                      // x complete everywhere
            lg_apply(c, p.x, V0);
            ++matvecs;
            for (int e = e0 + tid; e < e1; e += nth) V0[e] = p.b[e] - V0[e];
        }
        first = false;
        {
            double s2 = 0.0;
            for (int e = e0 + tid; e < e1; e += nth) s2 += V0[e] * V0[e];
            s2 = block_sum(s2, scr);
            if (tid == 0) mypart[0] = s2;
        }
        c.sync();
        const double res_norm = sqrt(lg_sum_partials(p, 0));
        res = res_norm;
        if (lead) {
            grs[0] = res_norm;
            p.ctrl_d[1] = res_norm;
        }
        __syncwarp();
        if (res == 0.0) {
            reason = R_ATOL;
            break;
        }
        {
            const double inv = 1.0 / res_norm;
            for (int e = e0 + tid; e < e1; e += nth) V0[e] *= inv;
        }
        const int it_total = it_arnoldi + aug_ct;
        reason = lg_converged(p, its, res, ttol, rnorm0);
        int loc_it = 0;
        bool hapend = false;
        double grs_loc = res_norm;
        c.sync();                          // V[0] complete everywhere

        while (reason == R_NONE && loc_it < it_total && its < p.max_it) {
            double* Vn = p.V + (long)(loc_it + 1) * nv;
            if (loc_it < it_arnoldi) {
                lg_apply(c, p.V + (long)loc_it * nv, Vn);
                ++matvecs;
            } else {
                const int order = loc_it - it_arnoldi + 1;
                int spot = 0;
                for (int ii = 0; ii < p.aug_dim; ++ii)
                    if (aug_order[ii] == order) {
                        spot = ii;
                        break;
                    }
                const double* src = p.AAUG + (long)spot * nv;
                for (int e = e0 + tid; e < e1; e += nth) Vn[e] = src[e];
            }
            // classical Gram-Schmidt, one pass: partial dots over the CTA's chunk.  A single-CTA solve (tiny) keeps every
            // partial result in shared memory: no global partial sums, no L2 round trips, half the barriers
            const bool from_matvec = tiny && loc_it < it_arnoldi;
            if (!from_matvec) {
                __syncthreads();
                for (int e = e0 + tid; e < e1; e += nth) wS[e - e0] = Vn[e];
                __syncthreads();
            }
            for (int i = wid; i <= loc_it; i += nw) {
                const double* Vi = p.V + (long)i * nv;
                double d = 0.0;
                for (int e = e0 + lane; e < e1; e += 32) d += Vi[e] * wS[e - e0];
                d = warp_sum(d);
                if (lane == 0) {
                    if (tiny) hS[i] = d;
                    else mypart[i] = d;
                }
            }
            if (tiny) {
                __syncthreads();
            } else {
                c.sync();
                for (int i = tid; i <= loc_it; i += nth) hS[i] = lg_sum_partials(p, i);
                __syncthreads();
            }
            double tt;
            {
                double s2 = 0.0;
                for (int e = e0 + tid; e < e1; e += nth) {
                    double w = wS[e - e0];
                    for (int i = 0; i <= loc_it; ++i) w -= hS[i] * p.V[(long)i * nv + e];
                    wS[e - e0] = w;
                    s2 += w * w;
                }
                s2 = block_sum(s2, scr);
                if (tiny) {
                    tt = sqrt(s2);
                } else {
                    if (tid == 0) mypart[p.max_k + 2] = s2;
                    c.sync();
                    tt = sqrt(lg_sum_partials(p, p.max_k + 2));
                }
            }
            double hapbnd = fabs(tt / grs_loc);
            if (hapbnd > p.haptol) hapbnd = p.haptol;
            if (tt > hapbnd) {
                const double inv = 1.0 / tt;
                for (int e = e0 + tid; e < e1; e += nth) Vn[e] = wS[e - e0] * inv;
            } else {
                for (int e = e0 + tid; e < e1; e += nth) Vn[e] = wS[e - e0];
                hapend = true;
            }
            if (lead_warp) {
                double* col = p.hh + (long)loc_it * ldh;
                double* hcol = p.hes + (long)loc_it * ldh;
                for (int i = lane; i <= loc_it + 1; i += 32) {
                    const double v = i <= loc_it ? hS[i] : tt;
                    colS[i] = v;
                    hcol[i] = v;
                }
                __syncwarp();
                if (lane == 0) {
                    for (int j = 0; j < loc_it; ++j) {
                        const double t0 = colS[j];
                        colS[j] = cc[j] * t0 + ss[j] * colS[j + 1];
                        colS[j + 1] = cc[j] * colS[j + 1] - ss[j] * t0;
                    }
                    double newres = 0.0;
                    ctlS[2] = 0.0;
                    if (!hapend) {
                        const double t0 = sqrt(colS[loc_it] * colS[loc_it] + colS[loc_it + 1] * colS[loc_it + 1]);
                        if (t0 == 0.0) {
                            p.ctrl_i[0] = 1;
                            ctlS[2] = 1.0;
                        } else {
                            cc[loc_it] = colS[loc_it] / t0;
                            ss[loc_it] = colS[loc_it + 1] / t0;
                            grs[loc_it + 1] = -(ss[loc_it] * grs[loc_it]);
                            grs[loc_it] = cc[loc_it] * grs[loc_it];
                            colS[loc_it] = cc[loc_it] * colS[loc_it] + ss[loc_it] * colS[loc_it + 1];
                            newres = fabs(grs[loc_it + 1]);
                        }
                    }
                    p.ctrl_d[0] = newres;
                    p.ctrl_d[1] = grs[loc_it + 1];
                    ctlS[0] = newres;               // single-CTA mode reads these back from shared memory
                    ctlS[1] = grs[loc_it + 1];
                }
                __syncwarp();
                for (int i = lane; i <= loc_it + 1; i += 32) col[i] = colS[i];
            }
            c.sync();
            if (tiny ? ctlS[2] != 0.0 : ld_cg_i(&p.ctrl_i[0]) != 0) {
                reason = R_NULL;
                break;
            }
            res = tiny ? ctlS[0] : ld_cg(&p.ctrl_d[0]);
            grs_loc = tiny ? ctlS[1] : ld_cg(&p.ctrl_d[1]);
            ++loc_it;
            ++its;
            reason = lg_converged(p, its, res, ttol, rnorm0);
            if (hapend && reason == R_NONE) {
                reason = R_BREAKDOWN;
                break;
            }
        }

        // ---- solution of this cycle ------------------------------------------------------
        const int it = loc_it - 1;
        int n_arn = 0, n_aug = 0;
        if (it >= 0) {
            if (it_arnoldi >= it + 1) {
                n_arn = it + 1;
            } else {
                n_arn = it_arnoldi;
                n_aug = it + 1 - it_arnoldi;
            }
            if (lead_warp) {
                // back-substitution R y = g, column oriented: one division per step on the diagonal, the update of the
                // remaining right-hand side spread over the lanes (columns of hh are contiguous)
                for (int i = lane; i <= it; i += 32) tS[i] = grs[i];
                __syncwarp();
                for (int k = it; k >= 0; --k) {
                    const double d = p.hh[(long)k * ldh + k];
                    const double yk = (k == it && d == 0.0) ? 0.0 : tS[k] / d;
                    if (lane == 0) nrs[k] = yk;
                    for (int i = lane; i < k; i += 32) tS[i] -= p.hh[(long)k * ldh + i] * yk;
                    __syncwarp();
                }
            }
            c.sync();
            for (int i = tid; i <= it; i += nth) hS[i] = ld_cg(&nrs[i]);
            __syncthreads();
            for (int e = e0 + tid; e < e1; e += nth) {
                double u = 0.0;
                for (int i = 0; i < n_arn; ++i) u += hS[i] * p.V[(long)i * nv + e];
                for (int ii = 0; ii < n_aug; ++ii) {
                    int spot = 0;
                    for (int jj = 0; jj < p.aug_dim; ++jj)
                        if (aug_order[jj] == ii + 1) {
                            spot = jj;
                            break;
                        }
                    u += hS[n_arn + ii] * p.AUG[(long)spot * nv + e];
                }
                p.upd[e] = u;
                p.x[e] += u;
            }
        }

        // ---- harvest the error approximation for the next cycle -----------------------------
        if (reason == R_NONE && its < p.max_it && p.aug_dim > 0) {
            int spot = 0;
            if (aug_ct == 0) {
                spot = 0;
                ++aug_ct;
            } else if (aug_ct < p.aug_dim) {
                spot = aug_ct;
                ++aug_ct;
            } else {
                for (int ii = 0; ii < p.aug_dim; ++ii)
                    if (aug_order[ii] == p.aug_dim) spot = ii;
            }
            {
                double s2 = 0.0;
                for (int e = e0 + tid; e < e1; e += nth) s2 += p.upd[e] * p.upd[e];
                s2 = block_sum(s2, scr);
                if (tid == 0) mypart[0] = s2;
            }
            if (lead_warp) {
                // avec = H y (the un-rotated Hessenberg matrix times the cycle's solution), one output entry per lane
                for (int jj = lane; jj <= it_total; jj += 32) {
                    double a = 0.0;
                    for (int ii = imax(0, jj - 1); ii < it_total; ++ii) a += p.hes[(long)ii * ldh + jj] * nrs[ii];
                    avec[jj] = a;
                }
            }
            c.sync();
            const double inv = 1.0 / sqrt(lg_sum_partials(p, 0));
            for (int i = tid; i <= it_total; i += nth) hS[i] = ld_cg(&avec[i]);
            __syncthreads();
            double* aug = p.AUG + (long)spot * nv;
            double* aaug = p.AAUG + (long)spot * nv;
            for (int e = e0 + tid; e < e1; e += nth) {
                aug[e] = p.upd[e] * inv;
                double u = 0.0;
                for (int i = 0; i <= it_total; ++i) u += hS[i] * p.V[(long)i * nv + e];
                aaug[e] = u * inv;
            }
            for (int ii = 0; ii < p.aug_dim; ++ii) aug_order[ii] += 1;
            aug_order[spot] = 1;
        }
        if (reason != R_NONE) break;
        if (its >= p.max_it) {
            reason = R_ITS;
            break;
        }
    }
    if (lead) {
        p.ctrl_i[4] = its;
        p.ctrl_i[5] = matvecs;
        p.ctrl_i[6] = reason;
        p.ctrl_i[7] = cycles;
        p.ctrl_d[2] = res;
    }
}

static int g_single_cta_threads = 512;

static int lg_setup(LgParams& p, int ineq, const ttipm_term* K00, const ttipm_term* K01, const ttipm_term* K21,
                    const ttipm_term* K22, const ttipm_term* K31, const ttipm_term* K33, const double* inv_I, int r,
                    int R, int nmode, int grid_hint, int* grid_out, int* smem_out, int max_k) {
    p.nblk = ineq ? 3 : 2;
    p.m = r * nmode * R;
    p.nv = p.nblk * p.m;
    p.inv_I = inv_I;
    const ttipm_term* src[7] = {K00, K01, K01, K21, K22, K31, K33};
    MvTerm t[7];
    int smax = 1, Smax = 1;
    for (int i = 0; i < (ineq ? 7 : 5); ++i) {
        if (!src[i]) return fail(1, "lgmres: missing operator block %d", i);
        if (convert_term(*src[i], t[i])) return fail(1, "lgmres: stride overflow");
        smax = imax(smax, imax(src[i]->s, src[i]->S));
        Smax = smax;
    }
    // phase A: K00 (0->0), K01 (1->0), K01^T (0->tmp), [K31 (1->2), K33 (2->2)]
    p.nA = 0;
    t[0].in_blk = 0; t[0].out_blk = 0; t[0].alpha = 1.0; p.tA[p.nA++] = t[0];
    t[1].in_blk = 1; t[1].out_blk = 0; t[1].alpha = 1.0; p.tA[p.nA++] = t[1];
    {   // transposed apply of K01: swap (l,r), (m,n), (L,R) strides and (s,S) stay
        MvTerm tt = t[2];
        int q;
        q = tt.p1s[0]; tt.p1s[0] = tt.p1s[2]; tt.p1s[2] = q;
        q = tt.as_[1]; tt.as_[1] = tt.as_[2]; tt.as_[2] = q;
        q = tt.p2s[0]; tt.p2s[0] = tt.p2s[2]; tt.p2s[2] = q;
        tt.in_blk = 0; tt.out_blk = 3; tt.alpha = 1.0;
        p.tA[p.nA++] = tt;
    }
    p.nslotA = 2; p.slotA[0] = 0; p.slotA[1] = 3; p.slotA[2] = 2;
    if (ineq) {
        t[5].in_blk = 1; t[5].out_blk = 2; t[5].alpha = 1.0; p.tA[p.nA++] = t[5];
        t[6].in_blk = 2; t[6].out_blk = 2; t[6].alpha = 1.0; p.tA[p.nA++] = t[6];
        p.nslotA = 3;
    }
    p.nB = 0;
    t[3].in_blk = 1; t[3].out_blk = 1; t[3].alpha = 1.0; p.tB[p.nB++] = t[3];
    t[4].in_blk = 3; t[4].out_blk = 1; t[4].alpha = -1.0; p.tB[p.nB++] = t[4];

    DevInfo di = dev_info();
    // grid: one CTA for tiny systems (block barriers only), otherwise fill the machine
    double flops = 0.0;
    for (int i = 0; i < (ineq ? 7 : 5); ++i)
        flops += 2.0 * r * nmode * R * R * src[i]->S + 2.0 * r * R * src[i]->s * nmode * nmode * src[i]->S +
                 2.0 * r * nmode * R * r * src[i]->s;
    int G = grid_hint > 0 ? grid_hint : (flops < 4e5 ? 1 : di.sms);
    if (G > di.sms) G = di.sms;
    // the Hessenberg / Givens arrays sit behind the matvec's carve-up: the matvec plan must leave room for them (a plan
    // that filled the shared memory to the last KB made the loop below alternate between G = items and 2 G forever:
    // graphm_3, IPM iteration 4, local block r = 150, R = 16, operator ranks 30)
    const int hess_bytes = (5 * (max_k + 2) + 4 + 2) * 8;
    int last_G = -1;
    for (;;) {
        if (mv_plan(p.g, r, R, r, R, nmode, smax, Smax, p.nslotA, G, di.smem_optin - 1024 - hess_bytes))
            return fail(4, "lgmres: local block r=%d R=%d s=%d does not fit in shared memory", r, R, smax);
        const int items = p.nslotA * p.g.ntiles;
        if (grid_hint <= 0 && G > items) G = items;
        if (G == last_G) return fail(4, "lgmres: no grid fits the shared memory for r=%d R=%d s=%d", r, R, smax);
        last_G = G;
        p.ch = (p.nv + G - 1) / G;
        p.oW = 0;
        p.oH = p.ch + (p.ch & 1);
        p.oScr = p.oH + max_k + 4;
        const int cgs_bytes = (p.oScr + 40) * 8;
        int bytes = imax(p.g.smem_bytes, cgs_bytes);
        p.oHess = (bytes + 15) / 16 * 2;
        bytes = (p.oHess + 5 * (max_k + 2) + 4) * 8;
        p.stage = 0;
        p.oStage = 0;
        if (G == 1) {
            long data = LG_STAGE_TERM_DOUBLES + p.nv;
            for (int i = 0; i < p.nA; ++i) data += (long)r * p.tA[i].s * r + (long)p.tA[i].s * nmode * nmode * p.tA[i].S + (long)R * p.tA[i].S * R;
            for (int i = 0; i < p.nB; ++i) data += (long)r * p.tB[i].s * r + (long)p.tB[i].s * nmode * nmode * p.tB[i].S + (long)R * p.tB[i].S * R;
            const long o = (bytes + 15) / 16 * 2;
            long tiny = 0;
            for (int i = 0; i < p.nA + p.nB; ++i) {
                const MvTerm& t = i < p.nA ? p.tA[i] : p.tB[i - p.nA];
                p.tinyT1[i] = (int)tiny; tiny += (long)r * R * nmode * t.S;
                p.tinyT2[i] = (int)tiny; tiny += (long)t.s * r * nmode * R;
            }
            p.tinyTmp = (int)tiny; tiny += p.m;
            if ((o + data + tiny) * 8 <= di.smem_optin) {
                p.stage = 1;
                p.oStage = (int)o;
                p.oTiny = (int)(o + data);
                bytes = (int)((o + data + tiny) * 8);
            }
        }
        if (bytes <= di.smem_optin) {
            *grid_out = G;
            *smem_out = bytes;
            return 0;
        }
        if (G >= di.sms) return fail(4, "lgmres: Krylov chunk of %d doubles does not fit in shared memory", p.ch);
        G = imin(di.sms, G * 2);
    }
}

// a single-CTA solve is bound by exposed instruction latency (two warps per scheduler at 256 threads): it runs 512
// threads; the multi-CTA grid keeps 256-thread CTAs
static int lg_threads(int G) {
#ifdef TTIPM_EMU
    (void)G;
    return block_threads();
#else
    return G == 1 ? g_single_cta_threads : 256;
#endif
}

struct LgInfoParams {
    const int* ci;
    const double* cd;
    double* out;
    int grid;
};
TT_GLOBAL void k_lg_info(const LgInfoParams p) {
    pdl_entry();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int i = 0; i < 4; ++i) p.out[i] = (double)p.ci[4 + i];
        p.out[4] = p.cd[2];
        p.out[5] = (double)p.grid;
    }
}
static int launch_info_copy(const int* ci, const double* cd, double* out, int grid, tt_stream_t st) {
    LgInfoParams ip{ci, cd, out, grid};
    return launch_kernel("k_lg_info", k_lg_info, dim3(1), dim3(32), 0, st, false, ip);
}

}  // namespace ttipm

using namespace ttipm;

// planning only: what ttipm_local_lgmres would launch for a block of this shape whose operator cores all have rank
// op_rank (no device access; the CPU tier pins the grid search against cycling with it)
extern "C" int ttipm_lgmres_plan(int ineq, int r, int R, int nmode, int op_rank, int restart, int* grid, int* smem_bytes) {
    if (r < 1 || R < 1 || nmode < 1 || op_rank < 1 || restart < 2) return fail(1, "lgmres_plan: bad arguments");
    ttipm_term t;
    memset(&t, 0, sizeof(t));
    t.s = op_rank; t.S = op_rank;
    t.p1_strides[0] = (int64_t)op_rank * r; t.p1_strides[1] = r; t.p1_strides[2] = 1;
    t.a_strides[0] = (int64_t)nmode * nmode * op_rank; t.a_strides[1] = (int64_t)nmode * op_rank; t.a_strides[2] = op_rank;
    t.a_strides[3] = 1;
    t.p2_strides[0] = (int64_t)op_rank * R; t.p2_strides[1] = R; t.p2_strides[2] = 1;
    t.alpha = 1.0;
    LgParams p;
    int G = 1, smem = 0;
    int rc = lg_setup(p, ineq, &t, &t, &t, &t, ineq ? &t : nullptr, ineq ? &t : nullptr, nullptr, r, R, nmode, 0, &G, &smem,
                      restart);
    if (rc) return rc;
    if (grid) *grid = G;
    if (smem_bytes) *smem_bytes = smem;
    return 0;
}

extern "C" int64_t ttipm_lgmres_workspace(int ineq, int r, int R, int nmode, int restart, int augment) {
    const int64_t nv = (int64_t)(ineq ? 3 : 2) * r * nmode * R, m = (int64_t)r * nmode * R;
    const int64_t mk = restart;
    DevInfo di = dev_info();
    int64_t n = (mk + 2) * nv + 2 * (int64_t)imax(augment, 1) * nv + nv + m;       // V, AUG, AAUG, upd, tmp
    n += (int64_t)di.sms * (mk + 3);                                                // part
    n += 2 * (mk + 2) * (mk + 1);                                                   // hh, hes
    n += 5 * (mk + 2) + 16 + 16;                                                    // small, ctrl_d, ctrl_i+barrier
    return n;
}

extern "C" int ttipm_local_lgmres(int ineq, const ttipm_term* K00, const ttipm_term* K01, const ttipm_term* K21,
                                  const ttipm_term* K22, const ttipm_term* K31, const ttipm_term* K33,
                                  const double* inv_I, int r, int R, int nmode, const double* rhs, double* x,
                                  double* workspace, int64_t ws_doubles, int restart, int augment, int max_it,
                                  double rtol, int apply_only, int grid_hint, double* info, void* stream) {
    if (restart < 2 || augment < 0 || augment >= restart || augment > 16)
        return fail(1, "lgmres: restart=%d augment=%d unsupported", restart, augment);
    if (ws_doubles < ttipm_lgmres_workspace(ineq, r, R, nmode, restart, augment))
        return fail(1, "lgmres: workspace too small");
    LgParams p;
    int G = 1, smem = 0;
    int rc = lg_setup(p, ineq, K00, K01, K21, K22, K31, K33, inv_I, r, R, nmode, grid_hint, &G, &smem, restart);
    if (rc) return rc;
    const int64_t nv = p.nv, mk = restart;
    DevInfo di = dev_info();
    double* w = workspace;
    p.V = w; w += (mk + 2) * nv;
    p.AUG = w; w += (int64_t)imax(augment, 1) * nv;
    p.AAUG = w; w += (int64_t)imax(augment, 1) * nv;
    p.upd = w; w += nv;
    p.tmp = w; w += p.m;
    p.part = w; w += (int64_t)di.sms * (mk + 3);
    p.hh = w; w += (mk + 2) * (mk + 1);
    p.hes = w; w += (mk + 2) * (mk + 1);
    p.small_ = w; w += 5 * (mk + 2);
    p.ctrl_d = w; w += 16;
    p.ctrl_i = (int*)w;
    p.barrier = (unsigned*)(p.ctrl_i + 16);
    p.b = rhs; p.x = x;
    p.max_k = restart; p.aug_dim = augment; p.max_it = max_it;
    p.rtol = rtol; p.abstol = 1e-50; p.dtol = 1e5; p.haptol = 1e-30;
    p.mode = apply_only ? LG_APPLY : LG_SOLVE;
    tt_stream_t st = (tt_stream_t)stream;
    if (dev_memset(p.ctrl_d, 0, 32 * 8, st)) return fail(5, "lgmres: memset failed");
    if (check_bound_device()) return 6;                 // the shared-memory opt-in below is per device
#ifndef TTIPM_EMU
    {   // the cooperative grid must be co-resident
        static int optin_done = 0;
        if (!optin_done) {
            cudaFuncSetAttribute((const void*)k_lgmres, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin);
            optin_done = 1;
        }
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lgmres, lg_threads(G), smem);
        if (per_sm < 1) return fail(4, "lgmres: kernel does not fit on an SM with %d B shared memory", smem);
        if (G > per_sm * di.sms) return fail(4, "lgmres: grid %d not co-resident", G);
    }
#endif
    rc = launch_kernel("k_lgmres", k_lgmres, dim3(G), dim3(lg_threads(G)), smem, st, true, p);
    if (rc) return rc;
    if (info) {
        // info[0..3] = its, matvecs, reason, cycles (as doubles), info[4] = final residual estimate, info[5] = grid
        return launch_info_copy(p.ctrl_i, p.ctrl_d, info, G, st);
    }
    return 0;
}
