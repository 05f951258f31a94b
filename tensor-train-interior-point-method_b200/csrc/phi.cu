// K2 interface updates, K3 right-hand-side contractions, K5 generic strided GEMM.
// Every contraction is a chain of tgemm calls with shared-memory intermediates.
#include "api_util.h"

namespace ttipm {

struct PhiTerm {
    const double* Phi;
    const double* A;
    double* out;
    int as_[4];
    int s, S;
};
struct PhiParams {
    int nterms;
    PhiTerm t[TTIPM_MAX_TERMS];
    const double* U;
    const double* V;
    int ul, uL, vr, vR, nm;
    int forward;
    int tile, ntiles;          // tile width over the free V index, tiles per term
    int ld1, ld2, ldA;
    int oT1, oT2, oAs, oOffs;  // in doubles
};

// forward : out[L',S,R'] = sum Phi[l,s,r] U[l,M,L'] A[s,M,N,S] V[r,N,R']   tile over R'
// backward: out[l,s,r]   = sum Phi[L,S,R] U[l,M,L]  A[s,M,N,S] V[r,N,R]    tile over r
TT_GLOBAL void __launch_bounds__(TT_MAX_THREADS) k_phi_update(const PhiParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    double* T1 = smem + p.oT1;
    double* T2 = smem + p.oT2;
    double* As = smem + p.oAs;
    int* offs = (int*)(smem + p.oOffs);
    const PhiTerm& t = p.t[blockIdx.x / p.ntiles];
    const int tl = blockIdx.x % p.ntiles;
    const int nm = p.nm, s = t.s, S = t.S, ul = p.ul, uL = p.uL, vr = p.vr, vR = p.vR;
    const int ld1 = p.ld1, ld2 = p.ld2, ldA = p.ldA;
    if (p.forward) {
        const int R0 = tl * p.tile, Rt = imin(p.tile, vR - R0);
        // As[(s,N),(M,S)] = A[s,M,N,S]
        for (int i = threadIdx.x; i < s * nm * nm * S; i += blockDim.x) {
            const int col = i % (nm * S), row = i / (nm * S);
            const int sg = row / nm, N = row % nm, M = col / S, Sp = col % S;
            As[row * ldA + col] = t.A[sg * t.as_[0] + M * t.as_[1] + N * t.as_[2] + Sp * t.as_[3]];
        }
        // T1[(l,rt),(s,N)] = sum_r Phi[(l,s),r] V[r,(N,R0+rt)]
        tgemm(ul * s, nm * Rt, vr, t.Phi, ax1(vr), ax1(1), p.V + R0, ax1(nm * vR), ax2(Rt, vR, 1),
              [&](int m, int n, double v) {
                  const int l = m / s, sg = m % s, N = n / Rt, rt = n % Rt;
                  T1[(l * Rt + rt) * ld1 + sg * nm + N] = v;
              },
              offs);
        // T2[(l,M),(S,rt)] = sum_(s,N) T1[(l,rt),(s,N)] As[(s,N),(M,S)]
        tgemm(ul * Rt, nm * S, s * nm, T1, ax1(ld1), ax1(1), As, ax1(ldA), ax1(1),
              [&](int m, int n, double v) {
                  const int l = m / Rt, rt = m % Rt, M = n / S, Sp = n % S;
                  T2[(l * nm + M) * ld2 + Sp * Rt + rt] = v;
              },
              offs);
        // out[L',(S,R0+rt)] = sum_(l,M) U[(l,M),L'] T2[(l,M),(S,rt)]
        double* out = t.out;
        tgemm(uL, S * Rt, ul * nm, p.U, ax1(1), ax1(uL), T2, ax1(ld2), ax1(1),
              [&](int m, int n, double v) {
                  const int Sp = n / Rt, rt = n % Rt;
                  out[((long)m * S + Sp) * vR + R0 + rt] = v;
              },
              offs);
    } else {
        const int r0 = tl * p.tile, rt_n = imin(p.tile, vr - r0);
        // As[(S,N),(s,M)] = A[s,M,N,S]
        for (int i = threadIdx.x; i < S * nm * s * nm; i += blockDim.x) {
            const int col = i % (s * nm), row = i / (s * nm);
            const int Sp = row / nm, N = row % nm, sg = col / nm, M = col % nm;
            As[row * ldA + col] = t.A[sg * t.as_[0] + M * t.as_[1] + N * t.as_[2] + Sp * t.as_[3]];
        }
        // T1[(L,rt),(S,N)] = sum_R Phi[(L,S),R] V[(r0+rt,N),R]
        tgemm(uL * S, rt_n * nm, vR, t.Phi, ax1(vR), ax1(1), p.V + (long)r0 * nm * vR, ax1(1), ax1(vR),
              [&](int m, int n, double v) {
                  const int L = m / S, Sp = m % S, rt = n / nm, N = n % nm;
                  T1[(L * rt_n + rt) * ld1 + Sp * nm + N] = v;
              },
              offs);
        // T2[(M,L),(s,rt)] = sum_(S,N) T1[(L,rt),(S,N)] As[(S,N),(s,M)]
        tgemm(uL * rt_n, s * nm, S * nm, T1, ax1(ld1), ax1(1), As, ax1(ldA), ax1(1),
              [&](int m, int n, double v) {
                  const int L = m / rt_n, rt = m % rt_n, sg = n / nm, M = n % nm;
                  T2[(M * uL + L) * ld2 + sg * rt_n + rt] = v;
              },
              offs);
        // out[l,(s,r0+rt)] = sum_(M,L) U[l,(M,L)] T2[(M,L),(s,rt)]
        double* out = t.out;
        tgemm(ul, s * rt_n, nm * uL, p.U, ax1(nm * uL), ax1(1), T2, ax1(ld2), ax1(1),
              [&](int m, int n, double v) {
                  const int sg = n / rt_n, rt = n % rt_n;
                  out[((long)m * s + sg) * vr + r0 + rt] = v;
              },
              offs);
    }
}

// ---------------------------------------------------------------------------------------
struct RhsTerm {
    const double* Xb1;
    const double* B;
    const double* Xb2;
    double* out;
    int b, Bp;
};
struct RhsParams {
    int nterms;
    RhsTerm t[TTIPM_MAX_TERMS];
    const double* core;
    int r, R, nm, mode;
    long out_rs;
    int tile, ntiles;
    int oT, oOffs;
};

TT_GLOBAL void __launch_bounds__(TT_MAX_THREADS) k_rhs_contract(const RhsParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    double* T = smem + p.oT;
    int* offs = (int*)(smem + p.oOffs);
    const RhsTerm& t = p.t[blockIdx.x / p.ntiles];
    const int tl = blockIdx.x % p.ntiles;
    const int nm = p.nm, b = t.b, Bp = t.Bp, r = p.r, R = p.R;
    if (p.mode == 0) {
        // rows rho in [r0, r0+rt): T[(rho,n),B'] = sum_b Xb1[b,rho] B[b,(n,B')] ; out[(rho,n),R] = T Xb2
        const int r0 = tl * p.tile, rt = imin(p.tile, r - r0);
        if (!t.B) {
            // "zero term": this output block has no right-hand-side core -- cleared here instead of by a memset node ahead of
            // the launch (a memset between two kernels breaks the programmatic dependent launch chain)
            for (int i = threadIdx.x; i < rt * nm * R; i += blockDim.x)
                t.out[(long)(r0 + i / (nm * R)) * p.out_rs + i % (nm * R)] = 0.0;
            return;
        }
        for (int i = threadIdx.x; i < rt * nm * Bp; i += blockDim.x) {
            const int bp = i % Bp, n = (i / Bp) % nm, rho = i / (Bp * nm);
            double acc = 0.0;
            for (int bb = 0; bb < b; ++bb) acc += t.Xb1[bb * r + r0 + rho] * t.B[(bb * nm + n) * Bp + bp];
            T[i] = acc;
        }
        __syncthreads();
        double* out = t.out;
        const long ors = p.out_rs;
        tgemm(rt * nm, R, Bp, T, ax1(Bp), ax1(1), t.Xb2, ax1(R), ax1(1),
              [&](int m, int n, double v) { out[(long)(r0 + m / nm) * ors + (m % nm) * R + n] = v; }, offs);
    } else if (p.mode == 1) {
        // T[(rho,n),B'] = sum_b Xb1[b,rho] B[b,n,B'] ; out[B',R] = sum_(rho,n) T[(rho,n),B'] core[(rho,n),R]
        for (int i = threadIdx.x; i < r * nm * Bp; i += blockDim.x) {
            const int bp = i % Bp, n = (i / Bp) % nm, rho = i / (Bp * nm);
            double acc = 0.0;
            for (int bb = 0; bb < b; ++bb) acc += t.Xb1[bb * r + rho] * t.B[(bb * nm + n) * Bp + bp];
            T[i] = acc;
        }
        double* out = t.out;
        tgemm(Bp, R, r * nm, T, ax1(1), ax1(Bp), p.core, ax1(R), ax1(1),
              [&](int m, int n, double v) { out[m * R + n] = v; }, offs);
    } else {
        // T[(b,n),Rho] = sum_B' B[(b,n),B'] Xb2[B',Rho] ; out[b,rho] = sum_(n,Rho) T[b,(n,Rho)] core[rho,(n,Rho)]
        tgemm(b * nm, R, Bp, t.B, ax1(Bp), ax1(1), t.Xb2, ax1(R), ax1(1),
              [&](int m, int n, double v) { T[m * R + n] = v; }, offs);
        double* out = t.out;
        tgemm(b, r, nm * R, T, ax1(nm * R), ax1(1), p.core, ax1(1), ax1(nm * R),
              [&](int m, int n, double v) { out[m * r + n] = v; }, offs);
    }
}

// ---------------------------------------------------------------------------------------
struct GemmParams {
    int M, N, K;
    double alpha, beta;
    const double* A;
    const double* B;
    double* C;
    long a_rs, a_cs, a_bs, b_rs, b_cs, b_bs, c_rs, c_cs, c_bs;
    int tm, tn, kc;
};

TT_GLOBAL void __launch_bounds__(TT_MAX_THREADS) k_gemm(const GemmParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    int* offs = (int*)smem_raw;
    const int tiles_n = (p.N + p.tn - 1) / p.tn;
    const int m0 = (blockIdx.x / tiles_n) * p.tm, n0 = (blockIdx.x % tiles_n) * p.tn;
    const int Mc = imin(p.tm, p.M - m0), Nc = imin(p.tn, p.N - n0);
    const double* A = p.A + blockIdx.y * p.a_bs + m0 * p.a_rs;
    const double* B = p.B + blockIdx.y * p.b_bs + n0 * p.b_cs;
    double* C = p.C + blockIdx.y * p.c_bs + m0 * p.c_rs + n0 * p.c_cs;
    const double alpha = p.alpha, beta = p.beta;
    const long crs = p.c_rs, ccs = p.c_cs;
    if (p.K == 0) {
        for (int i = threadIdx.x; i < Mc * Nc; i += blockDim.x) {
            double* c = C + (i / Nc) * crs + (i % Nc) * ccs;
            *c = beta == 0.0 ? 0.0 : beta * *c;
        }
        return;
    }
    for (int k0 = 0; k0 < p.K; k0 += p.kc) {
        const int Kc = imin(p.kc, p.K - k0);
        const bool first = k0 == 0;
        tgemm(Mc, Nc, Kc, A + k0 * p.a_cs, ax1((int)p.a_rs), ax1((int)p.a_cs), B + k0 * p.b_rs, ax1((int)p.b_rs),
              ax1((int)p.b_cs),
              [&](int m, int n, double v) {
                  double* c = C + m * crs + n * ccs;
                  if (first)
                      *c = (beta == 0.0 ? 0.0 : beta * *c) + alpha * v;
                  else
                      *c += alpha * v;
              },
              offs);
    }
}

}  // namespace ttipm

using namespace ttipm;

extern "C" int ttipm_phi_update(const ttipm_phi_term* terms, int nterms, int forward, const double* U, int ul,
                                int uL, const double* V, int vr, int vR, int nmode, void* stream) {
    if (nterms < 1 || nterms > TTIPM_MAX_TERMS) return fail(1, "phi_update: nterms=%d out of range", nterms);
    PhiParams p;
    p.nterms = nterms;
    int smax = 1, Smax = 1;
    for (int i = 0; i < nterms; ++i) {
        p.t[i].Phi = terms[i].Phi; p.t[i].A = terms[i].A; p.t[i].out = terms[i].out;
        for (int j = 0; j < 4; ++j) {
            if (!fits_int(terms[i].a_strides[j])) return fail(1, "phi_update: stride overflow");
            p.t[i].as_[j] = (int)terms[i].a_strides[j];
        }
        p.t[i].s = terms[i].s; p.t[i].S = terms[i].S;
        if (terms[i].s > smax) smax = terms[i].s;
        if (terms[i].S > Smax) Smax = terms[i].S;
    }
    p.U = U; p.V = V; p.ul = ul; p.uL = uL; p.vr = vr; p.vR = vR; p.nm = nmode; p.forward = forward;
    // large interfaces: three grouped contraction-GEMM launches (cgemm.cu); also the fallback when the fused kernel's
    // intermediates do not fit shared memory
    PhiTermLite lite[TTIPM_MAX_TERMS];
    for (int i = 0; i < nterms; ++i) {
        lite[i].Phi = p.t[i].Phi; lite[i].A = p.t[i].A; lite[i].out = p.t[i].out;
        for (int j = 0; j < 4; ++j) lite[i].as_[j] = p.t[i].as_[j];
        lite[i].s = p.t[i].s; lite[i].S = p.t[i].S;
    }
    if (phi_big_wanted(lite, nterms, forward, ul, uL, vr, vR, nmode))
        return phi_big(lite, nterms, forward, U, ul, uL, V, vr, vR, nmode, (tt_stream_t)stream);
    DevInfo di = dev_info();
    const int free_dim = forward ? vR : vr;
    int tile = (nterms * free_dim + 2 * di.sms - 1) / (2 * di.sms);
    if (tile < 1) tile = 1;
    for (;;) {
        p.tile = tile;
        p.ntiles = (free_dim + tile - 1) / tile;
        // forward: T1 (ul*tile) x (s*nm), T2 (ul*nm) x (S*tile), As (s*nm) x (nm*S)
        // backward: T1 (uL*tile) x (S*nm), T2 (nm*uL) x (s*tile), As (S*nm) x (s*nm)
        const int kk = (forward ? smax : Smax) * nmode;
        const int oo = (forward ? Smax : smax);
        const int rows1 = (forward ? ul : uL) * tile, rows2 = nmode * (forward ? ul : uL);
        p.ld1 = mv_pad(kk, 4, 8);
        p.ld2 = oo * tile + ((oo * tile) % 16 == 0 ? 8 : 0);
        p.ldA = nmode * oo + ((nmode * oo) % 16 == 0 ? 8 : 0);
        const int nT1 = rows1 * p.ld1, nT2 = rows2 * p.ld2, nAs = kk * p.ldA;
        int mM = imax(imax((forward ? ul * smax : uL * Smax), rows1), forward ? uL : ul);
        int mK = imax(imax(forward ? vr : vR, kk), rows2);
        int mN = imax(imax(nmode * tile, nmode * oo), oo * tile);
        p.oT1 = 0; p.oT2 = nT1; p.oAs = p.oT2 + nT2; p.oOffs = p.oAs + nAs;
        const int bytes = (p.oOffs + (mM + 2 * mK + mN + 1) / 2 + 8) * 8;
        if (bytes <= di.smem_optin) {
            return launch_kernel("k_phi_update", k_phi_update, dim3(nterms * p.ntiles), dim3(block_threads()), bytes,
                                 (tt_stream_t)stream, false, p);
        }
        if (tile == 1) return phi_big(lite, nterms, forward, U, ul, uL, V, vr, vR, nmode, (tt_stream_t)stream);
        tile /= 2;
    }
}

extern "C" int ttipm_rhs_contract(const ttipm_rhs_term* terms, int nterms, int mode, const double* core, int r, int R,
                                  int nmode, int64_t out_row_stride, void* stream) {
    if (nterms < 1 || nterms > TTIPM_MAX_TERMS) return fail(1, "rhs_contract: nterms=%d out of range", nterms);
    if (mode < 0 || mode > 2) return fail(1, "rhs_contract: mode=%d", mode);
    RhsParams p;
    p.nterms = nterms;
    int bmax = 1, Bmax = 1;
    for (int i = 0; i < nterms; ++i) {
        p.t[i].Xb1 = terms[i].Xb1; p.t[i].B = terms[i].B; p.t[i].Xb2 = terms[i].Xb2; p.t[i].out = terms[i].out;
        p.t[i].b = terms[i].b; p.t[i].Bp = terms[i].Bp;
        if (!terms[i].B) {
            if (mode != 0) return fail(1, "rhs_contract: zero terms (B == NULL) exist in mode 0 only");
            p.t[i].b = 0; p.t[i].Bp = 0;
            continue;
        }
        bmax = imax(bmax, terms[i].b); Bmax = imax(Bmax, terms[i].Bp);
    }
    p.core = core; p.r = r; p.R = R; p.nm = nmode; p.mode = mode; p.out_rs = out_row_stride;
    DevInfo di = dev_info();
    int tile = r, ntiles = 1;
    if (mode == 0) {
        tile = (nterms * r + 2 * di.sms - 1) / (2 * di.sms);
        if (tile < 1) tile = 1;
        ntiles = (r + tile - 1) / tile;
    }
    p.tile = tile; p.ntiles = ntiles;
    int nT, mM, mK, mN;
    if (mode == 0) { nT = tile * nmode * Bmax; mM = tile * nmode; mK = imax(bmax, Bmax); mN = imax(Bmax, R); }
    else if (mode == 1) { nT = r * nmode * Bmax; mM = Bmax; mK = r * nmode; mN = R; }
    else { nT = bmax * nmode * R; mM = bmax * nmode; mK = imax(Bmax, nmode * R); mN = imax(R, r); }
    p.oT = 0; p.oOffs = nT + (nT & 1);
    const int bytes = (p.oOffs + (mM + 2 * mK + mN + 1) / 2 + 8) * 8;
    if (bytes > di.smem_optin) return fail(4, "rhs_contract: shapes need %d B shared memory", bytes);
    return launch_kernel("k_rhs_contract", k_rhs_contract, dim3(nterms * ntiles), dim3(block_threads()), bytes,
                         (tt_stream_t)stream, false, p);
}

extern "C" int ttipm_gemm(int M, int N, int K, double alpha, const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs,
                          const double* B, int64_t b_rs, int64_t b_cs, int64_t b_bs, double beta, double* C,
                          int64_t c_rs, int64_t c_cs, int64_t c_bs, int nbatch, void* stream) {
    if (M < 1 || N < 1 || K < 0 || nbatch < 1) return fail(1, "gemm: bad dims %d %d %d batch %d", M, N, K, nbatch);
    if (!fits_int(a_rs * (int64_t)64) || !fits_int(a_cs * (int64_t)2048) || !fits_int(b_rs * (int64_t)2048) ||
        !fits_int(b_cs * (int64_t)64))
        return fail(1, "gemm: strides too large");
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.A = A; p.B = B; p.C = C;
    p.a_rs = a_rs; p.a_cs = a_cs; p.a_bs = a_bs; p.b_rs = b_rs; p.b_cs = b_cs; p.b_bs = b_bs;
    p.c_rs = c_rs; p.c_cs = c_cs; p.c_bs = c_bs;
    p.tm = 64; p.tn = 64; p.kc = 2048;
    // small problems: shrink the macro tile so that more CTAs share the work
    DevInfo di = dev_info();
    while (p.tm > 8 && ((M + p.tm - 1) / p.tm) * ((N + p.tn - 1) / p.tn) * nbatch < di.sms) {
        if (p.tm >= p.tn) p.tm /= 2; else p.tn /= 2;
        if (p.tn < 8) { p.tn = 8; break; }
    }
    const int tiles = ((M + p.tm - 1) / p.tm) * ((N + p.tn - 1) / p.tn);
    const int bytes = (p.tm + 2 * p.kc + p.tn + 8) * 4;
    return launch_kernel("k_gemm", k_gemm, dim3(tiles, nbatch), dim3(block_threads()), bytes, (tt_stream_t)stream,
                         false, p);
}
