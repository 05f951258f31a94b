// cgemm: grouped "contraction GEMMs" for the large-rank regime of the AMEn hot path.
//
// The fused kernels of matvec.cu / phi.cu keep both intermediates of the reference's 3-GEMM chain
// (cy_src/lgmres_cy.pyx:146-153, src/tt_als.py:190-257) in one CTA's shared memory, which is right while a local
// block is a few hundred KB.  At ranks r, R >= ~64 with operator ranks >= 8 every stage is a real GEMM
// (10^8..10^10 flop) and wants 2-D tiling over the whole machine.  This file provides that path:
//
//   * a problem = C(m, n) = c_scale * sum over segments  A_seg(m, k) B_seg(k, n)   (+ sub_scale * sub)
//     where every operand is addressed through composite-axis maps (AxisMap), so no operand of the TT
//     contraction is ever transposed or copied;
//   * a launch = a list of problems (all terms of one stage of the chain), one CTA per output tile;
//   * inside a tile: cp.async (LDGSTS) global -> shared through a 3-stage ring, operand tiles stored K-major
//     or M/N-major depending on which axis is contiguous in memory (coalesced loads either way) with row
//     paddings that make every DMMA fragment load bank-conflict free, fp64 tensor-core m8n8k4 warp tiles.
//
// tcgen05 / TMEM has no f64 kind, so DMMA is the fp64 tensor path on sm_100a; TMA tensor maps do not fit the
// composite (two-level) axes of the TT operands, hence per-element cp.async with precomputed offsets.
#pragma once
#include "common.cuh"

namespace ttipm {

#define CG_BK 16
#define CG_STAGES 3
#define CG_MAX_SEGS 40
#define CG_MAX_PROBS 40

struct CgSeg {
    const double* A;        // A(m, k) = A[axoff(aM, m) + axoff(aK, k) + batch * a_batch]
    const double* B;        // B(k, n) = B[axoff(bK, k) + axoff(bN, n) + batch * b_batch]
    AxisMap aM, aK, bK, bN;
    long a_batch, b_batch;
    int K;
    int a_kfast, b_kfast;   // 1: the K axis is the contiguous one in memory (tile stored [m][k]); 0: tile stored [k][m]
    int a_vec, b_vec;       // 1: the contiguous axis has unit stride and every row / column start is 16-byte aligned:
                            //    the loader moves two doubles per cp.async (set by cg_mark_vec on the host)
};

struct CgProb {
    int M, N;
    int seg0, nseg;
    double* C;              // C(m, n) at C[axoff(cM, m) + axoff(cN, n) + batch * c_batch]
    AxisMap cM, cN;
    long c_batch;
    double c_scale;
    const double* sub;      // optional, addressed like C without the batch stride
    double sub_scale;
    double* sumsq;          // optional: one partial sum of squares per tile at sumsq[batch * sumsq_batch + tile index]
    long sumsq_batch;
    int sumsq_slots;        // slots available per batch entry (tiles, or reduce CTAs, never exceed it)
    int tiles_m, tiles_n, tile0;
    // split-K (long K, few output tiles): split q of ksplit accumulates k-steps [q * steps_per, ...) and stores its raw
    // partial tile to part[(batch * ksplit + q) * M * N + m * N + n]; k_cg_reduce sums the splits in a fixed order and
    // applies the epilogue (deterministic, no atomics)
    int ksplit, steps_per;
    double* part;
};

struct CgParams {
    int nprob, jobs_per_batch, nbatch;
    int cfg;                // tile shape chosen by cg_plan: 0 = 128x128, 1 = 128x64, 2 = 64x64
    CgProb prob[CG_MAX_PROBS];
    CgSeg seg[CG_MAX_SEGS];
};

template <int BM, int BN, int WM, int WN>
struct CgCfg {
    static const int NT = WM * WN * 32;
    static const int MI = BM / WM / 8, NI = BN / WN / 8;
    static const int LDK = CG_BK + 4;                       // [x][k] layout: row stride = 4 (mod 16)
    static const int A_ELEMS = (BM * LDK > CG_BK * (BM + 4)) ? BM * LDK : CG_BK * (BM + 4);
    static const int B_ELEMS = (BN * LDK > CG_BK * (BN + 4)) ? BN * LDK : CG_BK * (BN + 4);
    static const int STAGE = A_ELEMS + B_ELEMS;
    static const int SMEM_DOUBLES = CG_STAGES * STAGE + CG_STAGES * CG_BK + 48;   // + K offsets (2 x 16 ints per stage) + scratch
};

#if defined(__CUDACC__) || defined(TTIPM_EMU)

#ifndef TTIPM_EMU
TT_DEV void cg_cp8(double* dst_smem, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
TT_DEV void cg_cp16(double* dst_smem, const double* src, int nvalid) {       // nvalid in {0, 1, 2} doubles, rest zero-filled
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int bytes = 8 * nvalid;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
TT_DEV void cg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
TT_DEV void cg_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#else
TT_DEV void cg_cp8(double* dst_smem, const double* src, bool valid) { *dst_smem = valid ? *src : 0.0; }
TT_DEV void cg_cp16(double* dst_smem, const double* src, int nvalid) {
    dst_smem[0] = nvalid > 0 ? src[0] : 0.0;
    dst_smem[1] = nvalid > 1 ? src[1] : 0.0;
}
TT_DEV void cg_commit() {}
template <int N>
TT_DEV void cg_wait() {}
#endif

// flattened K step -> (segment, k-tile) of a problem
TT_DEV void cg_locate(const CgParams& p, const CgProb& pr, int step, int& seg, int& kt) {
    seg = pr.seg0;
    for (int q = 0; q < pr.nseg; ++q) {
        const int nk = (p.seg[pr.seg0 + q].K + CG_BK - 1) / CG_BK;
        if (step < nk) {
            seg = pr.seg0 + q;
            kt = step;
            return;
        }
        step -= nk;
    }
    kt = 0;
}

// One output tile (m0.., n0..) of problem pr for batch entry `batch`.  All threads of the CTA call it.
template <int BM, int BN, int WM, int WN>
TT_DEV void cg_tile(const CgParams& p, const CgProb& pr, int batch, int m0, int n0, int tile_index, int split,
                    double* smem) {
    typedef CgCfg<BM, BN, WM, WN> Cfg;
    constexpr int NT = Cfg::NT, MI = Cfg::MI, NI = Cfg::NI, LDK = Cfg::LDK;
    constexpr int A_PASS = BM * CG_BK / NT, B_PASS = BN * CG_BK / NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = wid / WN, wn = wid % WN;
    int* koffs = (int*)(smem + CG_STAGES * Cfg::STAGE);              // [stage][2][16]
    double* scratch = smem + CG_STAGES * Cfg::STAGE + CG_STAGES * CG_BK;

    int nsteps = 0;
    for (int q = 0; q < pr.nseg; ++q) nsteps += (p.seg[pr.seg0 + q].K + CG_BK - 1) / CG_BK;
    const int step0 = split * pr.steps_per;                   // this CTA's range of flattened k-steps
    nsteps = imax(0, imin(nsteps - step0, pr.ksplit > 1 ? pr.steps_per : nsteps));

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // loader state of the segment currently being loaded
    int lseg = -1;
    long abase = 0, bbase = 0;       // batch offsets
    int aoff[A_PASS], boff[B_PASS];  // kfast: one row offset per pass; x-fast: [0] = this thread's row offset
    unsigned avalid = 0, bvalid = 0;

    auto prep = [&](int step) {      // threads < 32: K offsets of `step` into its ring slot
        int sg, kt;
        cg_locate(p, pr, step0 + step, sg, kt);
        const CgSeg& s = p.seg[sg];
        const int which = tid >> 4, kk = tid & 15, k = kt * CG_BK + kk;
        koffs[(step % CG_STAGES) * 32 + which * 16 + kk] = k < s.K ? axoff(which ? s.bK : s.aK, k) : 0;
    };
    auto issue = [&](int step) {
        int sg, kt;
        cg_locate(p, pr, step0 + step, sg, kt);
        const CgSeg& s = p.seg[sg];
        if (sg != lseg) {
            lseg = sg;
            abase = (long)batch * s.a_batch;
            bbase = (long)batch * s.b_batch;
            avalid = bvalid = 0;
            // row (column) offsets of this thread's copies; the thread <-> element mapping depends on the layout
            // (K-fast / M-fast) and on the copy width (one or two doubles)
            if (s.a_kfast) {
                const int rpp = s.a_vec ? NT / 8 : NT / 16, r0 = s.a_vec ? tid >> 3 : tid >> 4;
#pragma unroll
                for (int q = 0; q < A_PASS; ++q) {
                    const int m = m0 + r0 + q * rpp;
                    const bool ok = m < pr.M && r0 + q * rpp < BM;
                    aoff[q] = ok ? axoff(s.aM, m) : 0;
                    avalid |= ok ? (1u << q) : 0u;
                }
            } else {
                const int i = s.a_vec ? (tid % (BM / 2)) * 2 : tid % BM;
                const int left = pr.M - (m0 + i);
                aoff[0] = left > 0 ? axoff(s.aM, m0 + i) : 0;
                avalid = left > 1 ? 2u : (left > 0 ? 1u : 0u);          // valid elements of a 2-wide copy
            }
            if (s.b_kfast) {
                const int rpp = s.b_vec ? NT / 8 : NT / 16, r0 = s.b_vec ? tid >> 3 : tid >> 4;
#pragma unroll
                for (int q = 0; q < B_PASS; ++q) {
                    const int n = n0 + r0 + q * rpp;
                    const bool ok = n < pr.N && r0 + q * rpp < BN;
                    boff[q] = ok ? axoff(s.bN, n) : 0;
                    bvalid |= ok ? (1u << q) : 0u;
                }
            } else {
                const int i = s.b_vec ? (tid % (BN / 2)) * 2 : tid % BN;
                const int left = pr.N - (n0 + i);
                boff[0] = left > 0 ? axoff(s.bN, n0 + i) : 0;
                bvalid = left > 1 ? 2u : (left > 0 ? 1u : 0u);
            }
        }
        const int slot = step % CG_STAGES;
        double* As = smem + slot * Cfg::STAGE;
        double* Bs = As + Cfg::A_ELEMS;
        const int* ka = koffs + slot * 32;
        const int* kb = ka + 16;
        const int krem = s.K - kt * CG_BK;          // valid k's in this tile
        const double* Ab = s.A + abase;
        const double* Bb = s.B + bbase;
        if (s.a_kfast && s.a_vec) {
            const int kk = (tid & 7) * 2, i0 = tid >> 3;
            const int nv = imin(2, krem - kk);
            const int ko = ka[kk];
#pragma unroll
            for (int q = 0; q < A_PASS / 2; ++q) {
                const bool ok = nv > 0 && ((avalid >> q) & 1u);
                cg_cp16(As + (i0 + q * (NT / 8)) * LDK + kk, Ab + (ok ? aoff[q] + ko : 0), ok ? nv : 0);
            }
        } else if (s.a_kfast) {
            const int kk = tid & 15, i0 = tid >> 4;
            const bool kok = kk < krem;
            const int ko = ka[kk];
#pragma unroll
            for (int q = 0; q < A_PASS; ++q) {
                const bool ok = kok && ((avalid >> q) & 1u);
                cg_cp8(As + (i0 + q * (NT / 16)) * LDK + kk, Ab + (ok ? aoff[q] + ko : 0), ok);
            }
        } else if (s.a_vec) {
            const int i = (tid % (BM / 2)) * 2, kq = tid / (BM / 2);
#pragma unroll
            for (int q = 0; q < A_PASS / 2; ++q) {
                const int kk = kq + q * (NT / (BM / 2));
                const int nv = kk < krem ? (int)avalid : 0;
                cg_cp16(As + kk * (BM + 4) + i, Ab + (nv ? aoff[0] + ka[kk] : 0), nv);
            }
        } else {
            const int i = tid % BM, kq = tid / BM;
#pragma unroll
            for (int q = 0; q < A_PASS; ++q) {
                const int kk = kq + q * (NT / BM);
                const bool ok = avalid && kk < krem;
                cg_cp8(As + kk * (BM + 4) + i, Ab + (ok ? aoff[0] + ka[kk] : 0), ok);
            }
        }
        if (s.b_kfast && s.b_vec) {
            const int kk = (tid & 7) * 2, i0 = tid >> 3;
            const int nv = imin(2, krem - kk);
            const int ko = kb[kk];
#pragma unroll
            for (int q = 0; q < (B_PASS + 1) / 2; ++q) {
                const bool ok = nv > 0 && ((bvalid >> q) & 1u);
                if (i0 + q * (NT / 8) < BN) cg_cp16(Bs + (i0 + q * (NT / 8)) * LDK + kk, Bb + (ok ? boff[q] + ko : 0), ok ? nv : 0);
            }
        } else if (s.b_kfast) {
            const int kk = tid & 15, i0 = tid >> 4;
            const bool kok = kk < krem;
            const int ko = kb[kk];
#pragma unroll
            for (int q = 0; q < B_PASS; ++q) {
                const bool ok = kok && ((bvalid >> q) & 1u);
                cg_cp8(Bs + (i0 + q * (NT / 16)) * LDK + kk, Bb + (ok ? boff[q] + ko : 0), ok);
            }
        } else if (s.b_vec) {
            const int i = (tid % (BN / 2)) * 2, kq = tid / (BN / 2);
#pragma unroll
            for (int q = 0; q < (B_PASS + 1) / 2; ++q) {
                const int kk = kq + q * (NT / (BN / 2));
                const int nv = kk < krem ? (int)bvalid : 0;
                if (kk < CG_BK) cg_cp16(Bs + kk * (BN + 4) + i, Bb + (nv ? boff[0] + kb[kk] : 0), nv);
            }
        } else {
            const int i = tid % BN, kq = tid / BN;
#pragma unroll
            for (int q = 0; q < B_PASS; ++q) {
                const int kk = kq + q * (NT / BN);
                const bool ok = bvalid && kk < krem;
                cg_cp8(Bs + kk * (BN + 4) + i, Bb + (ok ? boff[0] + kb[kk] : 0), ok);
            }
        }
    };

    __syncthreads();                 // previous user of the shared memory is done
    if (tid < 32)
        for (int s_ = 0; s_ < CG_STAGES && s_ < nsteps; ++s_) prep(s_);
    __syncthreads();
    for (int s_ = 0; s_ < CG_STAGES - 1; ++s_) {
        if (s_ < nsteps) issue(s_);
        cg_commit();
    }
    for (int step = 0; step < nsteps; ++step) {
        cg_wait<CG_STAGES - 2>();
        __syncthreads();
        if (step + CG_STAGES - 1 < nsteps) issue(step + CG_STAGES - 1);
        cg_commit();
        if (tid < 32 && step + CG_STAGES < nsteps) prep(step + CG_STAGES);
        int sg, kt;
        cg_locate(p, pr, step0 + step, sg, kt);
        const CgSeg& s = p.seg[sg];
        const double* As = smem + (step % CG_STAGES) * Cfg::STAGE;
        const double* Bs = As + Cfg::A_ELEMS;
        const int sa_m = s.a_kfast ? LDK : 1, sa_k = s.a_kfast ? 1 : BM + 4;
        const int sb_n = s.b_kfast ? LDK : 1, sb_k = s.b_kfast ? 1 : BN + 4;
        const double* Aw = As + (wm * (BM / WM) + g) * sa_m + t * sa_k;
        const double* Bw = Bs + (wn * (BN / WN) + g) * sb_n + t * sb_k;
#pragma unroll
        for (int kk = 0; kk < CG_BK; kk += 4) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = Aw[8 * i * sa_m + kk * sa_k];
#pragma unroll
            for (int j = 0; j < NI; ++j) b[j] = Bw[8 * j * sb_n + kk * sb_k];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(a[i], b[j], acc[i][j][0], acc[i][j][1]);
        }
    }
    cg_wait<0>();

    if (pr.ksplit > 1) {               // raw partial tile, dense (M x N) per (batch, split)
        double* Pb = pr.part + ((long)batch * pr.ksplit + split) * pr.M * pr.N;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int m = m0 + wm * (BM / WM) + 8 * i + g;
            if (m >= pr.M) continue;
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int n = n0 + wn * (BN / WN) + 8 * j + 2 * t;
                if (n < pr.N) Pb[(long)m * pr.N + n] = acc[i][j][0];
                if (n + 1 < pr.N) Pb[(long)m * pr.N + n + 1] = acc[i][j][1];
            }
        }
        return;
    }
    // epilogue
    int om[MI], on[NI][2];
    bool mok[MI], nok[NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = m0 + wm * (BM / WM) + 8 * i + g;
        mok[i] = m < pr.M;
        om[i] = mok[i] ? axoff(pr.cM, m) : 0;
    }
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = n0 + wn * (BN / WN) + 8 * j + 2 * t + e;
            nok[j][e] = n < pr.N;
            on[j][e] = nok[j][e] ? axoff(pr.cN, n) : 0;
        }
    double* Cb = pr.C + (long)batch * pr.c_batch;
    double ss = 0.0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (mok[i] && nok[j][e]) {
                    const long o = (long)om[i] + on[j][e];
                    double v = pr.c_scale * acc[i][j][e];
                    if (pr.sub) v += pr.sub_scale * pr.sub[o];
                    Cb[o] = v;
                    ss += v * v;
                }
    if (pr.sumsq) {
        const double tot = block_sum(ss, scratch);
        if (tid == 0) pr.sumsq[(long)batch * pr.sumsq_batch + tile_index] = tot;
    }
}

template <int BM, int BN, int WM, int WN>
TT_GLOBAL void __launch_bounds__(WM * WN * 32) k_cgemm(const CgParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    const int batch = blockIdx.x / p.jobs_per_batch, job = blockIdx.x % p.jobs_per_batch;
    int q = 0;
    while (q + 1 < p.nprob && job >= p.prob[q + 1].tile0) ++q;
    const CgProb& pr = p.prob[q];
    const int ntiles = pr.tiles_m * pr.tiles_n;
    const int tl = (job - pr.tile0) % ntiles, split = (job - pr.tile0) / ntiles;
    cg_tile<BM, BN, WM, WN>(p, pr, batch, (tl / pr.tiles_n) * BM, (tl % pr.tiles_n) * BN, tl, split, smem);
}

// second pass of a split-K stage: C = c_scale * sum_splits part (+ sub_scale * sub), optional sum-of-squares partials
// (one per CTA of 256 outputs rows x ... -> slot = CTA index within the problem)
struct CgReduceParams {
    int nprob, nbatch;
    int cta0[CG_MAX_PROBS + 1];      // first CTA of each problem (per batch entry)
    CgProb prob[CG_MAX_PROBS];
};
TT_GLOBAL void __launch_bounds__(TT_MAX_THREADS) k_cg_reduce(const CgReduceParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* scratch = (double*)smem_raw;
    const int per_batch = p.cta0[p.nprob];
    const int batch = blockIdx.x / per_batch, cta = blockIdx.x % per_batch;
    int q = 0;
    while (q + 1 < p.nprob && cta >= p.cta0[q + 1]) ++q;
    const CgProb& pr = p.prob[q];
    const int local = cta - p.cta0[q], nctas = p.cta0[q + 1] - p.cta0[q];
    const long total = (long)pr.M * pr.N;
    double* Cb = pr.C + (long)batch * pr.c_batch;
    double ss = 0.0;
    for (long e = (long)local * blockDim.x + threadIdx.x; e < total; e += (long)nctas * blockDim.x) {
        const int m = (int)(e / pr.N), n = (int)(e % pr.N);
        double v = 0.0;
        for (int sp = 0; sp < pr.ksplit; ++sp) v += pr.part[((long)batch * pr.ksplit + sp) * total + e];
        const long o = (long)axoff(pr.cM, m) + axoff(pr.cN, n);
        v *= pr.c_scale;
        if (pr.sub) v += pr.sub_scale * pr.sub[o];
        Cb[o] = v;
        ss += v * v;
    }
    if (pr.sumsq) {
        const double tot = block_sum(ss, scratch);
        if (threadIdx.x == 0) pr.sumsq[(long)batch * pr.sumsq_batch + local] = tot;
    }
}
#endif

// host: mark the operands a loader may move two doubles at a time (unit stride along the contiguous axis, every other
// stride, the batch stride and the base address multiples of 16 bytes)
static inline bool cg_even(const AxisMap& a) { return (a.n1 == TT_AX_BIG || a.s0 % 2 == 0) && a.s1 % 2 == 0; }
static inline bool cg_unit(const AxisMap& a) { return a.n1 == TT_AX_BIG && a.s1 == 1; }
static inline void cg_mark_vec(CgSeg& s) {
    const bool a_al = ((uintptr_t)s.A % 16 == 0) && s.a_batch % 2 == 0, b_al = ((uintptr_t)s.B % 16 == 0) && s.b_batch % 2 == 0;
    s.a_vec = a_al && (s.a_kfast ? (cg_unit(s.aK) && cg_even(s.aM)) : (cg_unit(s.aM) && cg_even(s.aK)));
    s.b_vec = b_al && (s.b_kfast ? (cg_unit(s.bK) && cg_even(s.bN)) : (cg_unit(s.bN) && cg_even(s.bK)));
}

// host: tile counts / job ranges for a tile shape; returns the number of jobs per batch entry
static inline int cg_plan_tiles(CgParams& p, int BM, int BN) {
    int jobs = 0;
    for (int q = 0; q < p.nprob; ++q) {
        CgProb& pr = p.prob[q];
        pr.tiles_m = (pr.M + BM - 1) / BM;
        pr.tiles_n = (pr.N + BN - 1) / BN;
        pr.tile0 = jobs;
        if (pr.ksplit < 1) pr.ksplit = 1;
        jobs += pr.tiles_m * pr.tiles_n * pr.ksplit;
    }
    p.jobs_per_batch = jobs;
    return jobs;
}

// host: launch one grouped-GEMM stage (defined in cgemm.cu).  The tile shape (128x128 / 128x64 / 64x64) and the
// split-K factors are chosen from the problem list; `part` is scratch for split-K partial tiles (cg_part_doubles).
long cg_plan(CgParams& p, int sms);          // chooses p.cfg and the split-K factors; returns the doubles of `part` needed
int cg_launch(CgParams& p, double* part, tt_stream_t st);

}  // namespace ttipm
