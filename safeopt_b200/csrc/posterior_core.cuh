// Device building blocks shared by the posterior kernels (posterior.cu, posterior_tma.cuh) and the batched expander
// test (expander.cu): tile parameters, shared-memory layout, the kernel-row generators that write the k(x*, X) tile
// straight into DMMA B-fragment order, the DMMA K-segment loop and the tile-end reduction.
#pragma once
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kGridMaxDim = 6;   // grid fast path (larger d uses explicit rows)
constexpr int kMaxOut = 4;       // GPs evaluated by one launch when they share the factorisation (so_posterior_*_multi)
constexpr int kMaxPass = 8;      // passes of four block rows per warp: 8 warps x 8 x 4 = 256 block rows (N = 2048)
constexpr int kMaxSlots = 6;     // block rows per warp and pass: 4 (default), 2 (16-warp / two-CTA variants) or 6 (N = 281..384)

struct PostParams {
    int N, NB, d, RG, CG, T, TB, npass, kind;
    const double2* Afrag;        // L^-1 packed in DMMA A-fragment order (fit.cu: k_pack_afrag)
    const double* zvec;          // z = L^-1 y, zero padded to 8*NB: mean(x*) = (L^-1 k).z
    const double* Xs;            // training inputs scaled by 1/lengthscale, 8*NB x d
    double inv_ls[SO_MAX_DIM];
    double variance;
    const double* Xstar;         // explicit candidate rows (M x d) or nullptr
    int64_t M, row0, ntiles;
    int gd;                      // grid description (so_grid_define) for the index-generated rows
    int gn[kGridMaxDim];
    int goff[kGridMaxDim];
    int64_t gstride[kGridMaxDim];
    const double* E;             // per-axis factor tables, (sum n_j) x 8*NB
    double beta, fmin;
    double* mean;
    double* var;
    double* Q;
    int q_stride, q_col;
    uint8_t* S;
    int safe_mode;
    // Further GPs that share X, kernel and noise with this one (same L^-1, same kernel rows, same variance): only
    // z_o = L^-1 y_o differs, so they cost one more V.z_o per row instead of one more contraction.  Output 0 is described
    // by the fields above, outputs 1..n_out-1 by the arrays below.
    int n_out;
    const double* zvec_x[kMaxOut - 1];
    double fmin_x[kMaxOut - 1];
    double* mean_x[kMaxOut - 1];
    double* var_x[kMaxOut - 1];
    int q_col_x[kMaxOut - 1];
    // Block rows of every warp per pass when the eight warps split the rows (RG == 8): for NB a multiple of 32 the closed-form
    // pairing {g, 15-g, 16+g, 31-g} balances the triangular work exactly; for any other NB (a BO loop adds one observation
    // per iteration) the host assigns the rows longest-first to the least loaded warp (plan_rows) -- with the closed form,
    // NB = 33 put the whole extra block row on one warp in a second pass and a tile took 1.7x as long.
    int use_row_table;
    short row_table[kMaxPass][8][kMaxSlots];    // ascending per pass, -1 = unused slot (suffix); NB <= 256
};

struct SmemLayout {
    size_t k_bytes, xs_off, xt_off, ss_off, mean_off, meanx_off, exp_off, total;
};

// [ Kx tile | scaled training inputs | two tiles of candidate rows | |V|^2 partials x2 | mean partials x2 |
//   mean partials of the further outputs x2 | exp table ]
__host__ __device__ inline SmemLayout smem_layout(int NB, int T, int d, int RG, bool grid, int n_extra = 0) {
    SmemLayout L;
    const size_t Npad = 8 * (size_t)NB;
    L.k_bytes = Npad * T * sizeof(double);
    L.xs_off = L.k_bytes;
    L.xt_off = L.xs_off + (grid ? 0 : Npad * d * sizeof(double));
    L.ss_off = L.xt_off + (grid ? 0 : 2 * (size_t)T * d * sizeof(double));
    L.mean_off = L.ss_off + 2 * (size_t)RG * T * sizeof(double);
    L.meanx_off = L.mean_off + 2 * (size_t)RG * T * sizeof(double);
    L.exp_off = L.meanx_off + 2 * (size_t)n_extra * RG * T * sizeof(double);
    L.total = L.exp_off + 64 * sizeof(double);
    return L;
}

// ---------------------------------------------------------------- gen: explicit rows
// Lane l of the warp that owns column tile ct produces, for every k-block kb, the pair
// (k(x*_t, x_n0), k(x*_t, x_n0+1)) with t = 8 ct + l/4, n0 = 8 kb + 2 (l%4) -- exactly its slot of the B fragment.
// The generator is bound by the fp64 dependency chain (distance -> exp -> scale), not by the pipe: with 8 warps per SM
// only ILP hides the latency, so the body is branch-free and works on two k-blocks = four kernel values at a time, with
// the candidate row held in registers (D = compile-time dimension; D = 0 keeps the runtime loop for d > 6).
#ifndef SO_K2_GEN_SPLIT
#define SO_K2_GEN_SPLIT 1
#endif
template <int KIND, int D>
__device__ __forceinline__ void gen_rows_d(const PostParams& p, double2* __restrict__ sK, const double* __restrict__ sXs,
                                           const double* __restrict__ sXt, const double* __restrict__ sExpT, int warp, int lane) {
    const int d = D ? D : p.d, N = p.N, NB = p.NB, TB = p.TB;
    const int q = lane & 3, tl = lane >> 2;
    const double variance = p.variance;
    // work items = (column tile, part of the k range): with fewer column tiles than warps (TB = 2, 4, 6) the k-blocks of a
    // column tile are split over 4, 2, 4 warps so that all eight warps generate
#if SO_K2_GEN_SPLIT
    const int S = TB == 4 ? 2 : ((TB == 2 || TB == 6) ? 4 : 1);
#else
    const int S = 1;
#endif
    const int kb_part = (((NB + S - 1) / S) + 1) & ~1;          // even, so that the pairs (kb, kb + 1) never straddle parts
    for (int item = warp; item < TB * S; item += kWarps) {
        const int ct = item % TB, part = item / TB;
        const int kb_lo = part * kb_part, kb_hi = min(NB, kb_lo + kb_part);
        const double* xt = sXt + (ct * 8 + tl) * d;
        double xr[D ? D : 1];
#pragma unroll
        for (int j = 0; j < D; ++j) xr[j] = xt[j];
        for (int kb = kb_lo; kb < kb_hi; kb += 2) {
            const int kb1 = kb + 1 < NB ? kb + 1 : kb;        // odd NB: the last pass recomputes block kb (same store)
            const int n0 = 8 * kb + 2 * q, n1 = 8 * kb1 + 2 * q;
            const double* x0 = sXs + n0 * d;
            const double* x1 = sXs + n1 * d;
            double r00 = 0.0, r01 = 0.0, r10 = 0.0, r11 = 0.0;
            if (D) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const double a = xr[j] - x0[j], b = xr[j] - x0[D + j], c = xr[j] - x1[j], e = xr[j] - x1[D + j];
                    r00 = fma(a, a, r00); r01 = fma(b, b, r01); r10 = fma(c, c, r10); r11 = fma(e, e, r11);
                }
            } else {
                for (int j = 0; j < d; ++j) {
                    const double xv = xt[j];
                    const double a = xv - x0[j], b = xv - x0[d + j], c = xv - x1[j], e = xv - x1[d + j];
                    r00 = fma(a, a, r00); r01 = fma(b, b, r01); r10 = fma(c, c, r10); r11 = fma(e, e, r11);
                }
            }
            double k00 = kernel_of_r2_fast<KIND>(r00, variance, sExpT);
            double k01 = kernel_of_r2_fast<KIND>(r01, variance, sExpT);
            double k10 = kernel_of_r2_fast<KIND>(r10, variance, sExpT);
            double k11 = kernel_of_r2_fast<KIND>(r11, variance, sExpT);
            k00 = n0 < N ? k00 : 0.0;                          // zero padding of the last k-block
            k01 = n0 + 1 < N ? k01 : 0.0;
            k10 = n1 < N ? k10 : 0.0;
            k11 = n1 + 1 < N ? k11 : 0.0;
            sK[(kb * TB + ct) * 32 + lane] = make_double2(k00, k01);
            sK[(kb1 * TB + ct) * 32 + lane] = make_double2(k10, k11);
        }
    }
}

template <int KIND>
__device__ __forceinline__ void gen_rows(const PostParams& p, double2* __restrict__ sK, const double* __restrict__ sXs,
                                         const double* __restrict__ sXt, const double* __restrict__ sExpT, int warp, int lane) {
    switch (p.d) {
        case 1: gen_rows_d<KIND, 1>(p, sK, sXs, sXt, sExpT, warp, lane); break;
        case 2: gen_rows_d<KIND, 2>(p, sK, sXs, sXt, sExpT, warp, lane); break;
        case 3: gen_rows_d<KIND, 3>(p, sK, sXs, sXt, sExpT, warp, lane); break;
        case 4: gen_rows_d<KIND, 4>(p, sK, sXs, sXt, sExpT, warp, lane); break;
        case 5: gen_rows_d<KIND, 5>(p, sK, sXs, sXt, sExpT, warp, lane); break;
        case 6: gen_rows_d<KIND, 6>(p, sK, sXs, sXt, sExpT, warp, lane); break;
        default: gen_rows_d<KIND, 0>(p, sK, sXs, sXt, sExpT, warp, lane); break;
    }
}

// ---------------------------------------------------------------- gen: separable RBF on a grid (per-axis tables)
// k(x*, x_n) = prod_j E_j[idx_j(row)][n]; the tables (sum_j n_j rows of 8*NB doubles, built once per fit by
// k_grid_tables) replace N fp64 exp() per row by (d-1) multiplies -- exp costs ~21 FMA slots of the one FP64 pipe the
// contraction also needs (profiles/r01_fp64_rates_b200.jsonl).  Used by the expander kernel and as the fallback of the
// grid path when the scaled-operand tables of the TMA kernel would not fit.
__device__ __forceinline__ void gen_grid(const PostParams& p, double2* __restrict__ sK, int64_t tile_row0, int warp, int lane) {
    const int NB = p.NB, TB = p.TB, Npad = 8 * p.NB;
    const int q = lane & 3, tl = lane >> 2;
    const int64_t last = p.row0 + p.M - 1;
    for (int ct = warp; ct < TB; ct += kWarps) {
        int64_t row = tile_row0 + ct * 8 + tl;
        if (row > last) row = last;
        const double2* e[kGridMaxDim];
#pragma unroll
        for (int j = 0; j < kGridMaxDim; ++j) {
            if (j < p.gd) {
                const int idx = (int)((row / p.gstride[j]) % p.gn[j]);
                e[j] = reinterpret_cast<const double2*>(p.E + (size_t)(p.goff[j] + idx) * Npad) + q;
            } else {
                e[j] = nullptr;
            }
        }
        for (int kb = 0; kb < NB; ++kb) {
            double2 v = __ldg(e[0] + 4 * kb);
#pragma unroll
            for (int j = 1; j < kGridMaxDim; ++j) {
                if (j < p.gd) {
                    const double2 w = __ldg(e[j] + 4 * kb);
                    v.x *= w.x;
                    v.y *= w.y;
                }
            }
            sK[(kb * TB + ct) * 32 + lane] = v;
        }
    }
}

// ---------------------------------------------------------------- contraction
// One K segment: accumulator slots FIRST..NS-1 are active.  `a` holds the fragments of the current k-block on entry and
// those of k-block kb_hi+1 on exit (software prefetch into registers, distance one block; the packed operand has slack
// blocks behind it).  The fragments of k-block kb + SO_K2_A_PF_L1 are pulled from L2 into L1 meanwhile (CCTL.PF1, no
// registers held): the A stream has no reuse, so without it every register prefetch pays a full L2 round trip and with
// two to four warps per scheduler those waits coincide often enough to idle the FP64 pipe (measured: 14.0 -> 13.4 ms).
#ifndef SO_K2_A_PF_L1
#define SO_K2_A_PF_L1 2
#endif
// Where the B fragments of a k-block come from.  PlainB: the whole tile is resident in shared memory (bulk kernel, TMA
// double-buffer kernel).  The ring kernel (posterior_ring.cuh) streams k-block chunks and synchronises in acquire/release.
// A source hands out B in spans of consecutive k-blocks that need no synchronisation inside (kSpan = whole range for PlainB,
// one ring stage for RingB): the k-block loop is split into an outer loop over spans (acquire / release, possibly blocking)
// and a plain inner loop that ptxas can software-pipeline.
struct PlainB {
    static constexpr bool kStreaming = false;
    static constexpr int kSpan = 1 << 20;
    const double2* sB;      // tile base + this warp's column group + lane
    int TB;
    __device__ __forceinline__ void enter(int) const {}                      // first k-block of a span is about to be read
    __device__ __forceinline__ const double2* at(int kb) const { return sB + (size_t)kb * TB * 32; }
    __device__ __forceinline__ void leave(int) const {}                      // last k-block of a span has been read
    __device__ __forceinline__ void drain(int) const {}
};

// One k-block of a segment: prefetch the next block's A fragments (registers one block ahead, L1 two blocks ahead), then
// 2 * BT * (NS - FIRST) DMMAs on the current one.
template <int BT, int NS, int FIRST>
__device__ __forceinline__ void mma_kblock(double (&acc)[NS][BT][2], double2 (&a)[NS], const double2* __restrict__ Afrag,
                                           const size_t (&abase)[NS], const double2* __restrict__ bp, int kb) {
    double2 an[NS];
#pragma unroll
    for (int s = FIRST; s < NS; ++s) an[s] = __ldg(Afrag + abase[s] + (size_t)(kb + 1) * 32);
#if SO_K2_A_PF_L1 > 0
#pragma unroll
    for (int s = FIRST; s < NS; ++s)
        asm volatile("prefetch.global.L1 [%0];\n" ::"l"(Afrag + abase[s] + (size_t)(kb + SO_K2_A_PF_L1) * 32));
#endif
#pragma unroll
    for (int c = 0; c < BT; ++c) {
        const double2 b = bp[c * 32];
#pragma unroll
        for (int s = FIRST; s < NS; ++s) {
            dmma884(acc[s][c][0], acc[s][c][1], a[s].x, b.x);
            dmma884(acc[s][c][0], acc[s][c][1], a[s].y, b.y);
        }
    }
#pragma unroll
    for (int s = FIRST; s < NS; ++s) a[s] = an[s];
}

template <int BT, int NS, int FIRST, typename BS>
__device__ __forceinline__ void mma_segment(double (&acc)[NS][BT][2], double2 (&a)[NS], const double2* __restrict__ Afrag,
                                            const size_t (&abase)[NS], BS& bs, int kb_lo, int kb_hi) {
    if (!BS::kStreaming) {
        // resident tile: one plain loop (ptxas software-pipelines the B loads across k-blocks)
        for (int kb = kb_lo; kb <= kb_hi; ++kb) mma_kblock<BT, NS, FIRST>(acc, a, Afrag, abase, bs.at(kb), kb);
        return;
    }
    // streamed B: outer loop over the spans of k-blocks inside one ring stage (enter / leave may block), plain inner loop
    for (int kb0 = kb_lo; kb0 <= kb_hi;) {
        const int span_end = (kb0 / BS::kSpan) * BS::kSpan + BS::kSpan - 1;
        const int kb1 = span_end < kb_hi ? span_end : kb_hi;
        if (kb0 % BS::kSpan == 0) bs.enter(kb0);
        for (int kb = kb0; kb <= kb1; ++kb) mma_kblock<BT, NS, FIRST>(acc, a, Afrag, abase, bs.at(kb), kb);
        bs.leave(kb1);
        kb0 = kb1 + 1;
    }
}

template <int BT, int NS, int FIRST, typename BS>
struct SegmentChain {       // segments FIRST..NS-1 in order; segment s covers k-blocks ext[s-1]+1 .. ext[s]
    static __device__ __forceinline__ void run(double (&acc)[NS][BT][2], double2 (&a)[NS], const double2* __restrict__ Afrag,
                                               const size_t (&abase)[NS], BS& bs, const int (&ext)[NS]) {
        mma_segment<BT, NS, FIRST, BS>(acc, a, Afrag, abase, bs, FIRST == 0 ? 0 : ext[FIRST > 0 ? FIRST - 1 : 0] + 1, ext[FIRST]);
        SegmentChain<BT, NS, FIRST + 1, BS>::run(acc, a, Afrag, abase, bs, ext);
    }
};
template <int BT, int NS, typename BS>
struct SegmentChain<BT, NS, NS, BS> {
    static __device__ __forceinline__ void run(double (&)[NS][BT][2], double2 (&)[NS], const double2* __restrict__,
                                               const size_t (&)[NS], BS&, const int (&)[NS]) {}
};

// Block rows of a warp in one pass, ascending: {g, 2RG-1-g} (NS = 2) or {g, 2RG-1-g, 2RG+g, 4RG-1-g} (NS = 4) above
// base = NS * RG * pass -- a pairing that gives every warp the same triangular work.
template <int NS>
__device__ __forceinline__ int warp_row(int base, int RG, int g, int i) {
    const int r0 = base + g, r1 = base + 2 * RG - 1 - g;
    if (NS == 2) return i == 0 ? r0 : r1;
    const int r2 = base + 2 * RG + g, r3 = base + 4 * RG - 1 - g;
    return i == 0 ? r0 : (i == 1 ? r1 : (i == 2 ? r2 : r3));
}

// Recursive-halving reduction of K values per lane over the 8 lanes that hold the rows of one 8x8 block (lane bits
// 4,3,2): after the three stages every lane owns K/8 fully reduced values, at 7K/8 shuffle+add pairs per lane instead of
// 3K for a butterfly.  Scalar fp64 instructions are precious here: they share the one FP64 pipe with DMMA and are served
// behind it (profiles/r01_k2_variants.md).  Lane (b4,b3,b2) ends up with original indices b4*K/2 + b3*K/4 + b2*K/8 + [0, K/8).
template <int K>
__device__ __forceinline__ void halving_reduce(double (&v)[K], int lane) {
    static_assert(K % 8 == 0, "K must be a multiple of 8");
#pragma unroll
    for (int i = 0; i < K / 2; ++i) {
        const bool up = (lane & 16) != 0;
        const double send = up ? v[i] : v[i + K / 2];
        const double keep = up ? v[i + K / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < K / 4; ++i) {
        const bool up = (lane & 8) != 0;
        const double send = up ? v[i] : v[i + K / 4];
        const double keep = up ? v[i + K / 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < K / 8; ++i) {
        const bool up = (lane & 4) != 0;
        const double send = up ? v[i] : v[i + K / 8];
        const double keep = up ? v[i + K / 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
}

// The whole contraction of one tile for one warp: V = A.B over the warp's NS block rows per pass (warp_row) and its BT
// column tiles, then per column
//     |V|^2 partial (sum of squares of this warp's rows)  and  V.z partial (the mean's share, z = L^-1 y)
// taken from the accumulators in one pass, reduced over the 8 row lanes by recursive halving and left in this row
// group's slot of sSS / sMean (fixed order everywhere => bit-reproducible).  NS = 4 with 8 warps per CTA, NS = 2 with 16
// (half the accumulators per warp, twice the warps per scheduler to cover each other's waits).
template <int BT, int NS = 4, typename BS = PlainB>
__device__ __forceinline__ void contract_tile(const PostParams& p, const double2* __restrict__ Afrag_lane, BS& bs,
                                              double* __restrict__ sSST, double* __restrict__ sMeanT,
                                              double* __restrict__ sMeanXT, int g, int cg, int lane) {
    static_assert(BT % 2 == 0, "BT must be even");
    static_assert(NS == 2 || NS == 4 || NS == 6, "two, four or six block rows per warp and pass");
    const int RG = p.RG, NB = p.NB, T = p.T;
    for (int pass = 0; pass < p.npass; ++pass) {
        const int base = NS * RG * pass;
        // rows are ascending; the ones beyond NB (inactive) form a suffix.  Slots are ordered by K extent, inactive
        // slots (extent -1) first, so that "slots FIRST..NS-1 active" holds in every segment.
        int rows[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            int r;
            if constexpr (NS == 6) r = (int)p.row_table[pass][g][i];           // six rows per warp: always planned by the host
            else r = (NS == 4 && p.use_row_table) ? (int)p.row_table[pass][g][i & 3] : warp_row<NS>(base, RG, g, i);
            rows[i] = r < NB ? r : -1;
        }
        int na = 0;
#pragma unroll
        for (int i = 0; i < NS; ++i) na += rows[i] >= 0;
        int ext[NS];
        size_t abase[NS];
        double zs[NS];             // z at this lane's row of each slot's block (0 for inactive slots)
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int src = s - (NS - na);
            int r = -1;
#pragma unroll
            for (int i = 0; i < NS; ++i) r = (i == src) ? rows[i] : r;
            ext[s] = r;
            abase[s] = r >= 0 ? (size_t)r * (r + 1) / 2 * 32 : 0;
            zs[s] = r >= 0 ? __ldg(p.zvec + 8 * r + (lane >> 2)) : 0.0;
        }
        double acc[NS][BT][2];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int c = 0; c < BT; ++c) { acc[s][c][0] = 0.0; acc[s][c][1] = 0.0; }
        double2 a[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) a[s] = __ldg(Afrag_lane + abase[s]);
        SegmentChain<BT, NS, 0, BS>::run(acc, a, Afrag_lane, abase, bs, ext);
        bs.drain(ext[NS - 1] + 1);      // streamed B: the k-blocks this warp has no rows for are still handed back
        // red[c*2+h] = sum of squares, red[2BT + c*2+h] = mean share, for column 8c + 2(lane%4) + h of this column group
        double red[4 * BT];
#pragma unroll
        for (int c = 0; c < BT; ++c)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                double q2 = 0.0, mz = 0.0;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    q2 = fma(acc[s][c][hh], acc[s][c][hh], q2);
                    mz = fma(acc[s][c][hh], zs[s], mz);
                }
                red[c * 2 + hh] = q2;
                red[2 * BT + c * 2 + hh] = mz;
            }
        halving_reduce<4 * BT>(red, lane);
        // lane (b4,b3,b2) owns original indices first + [0, BT/2)
        const int first = ((lane >> 4) & 1) * (2 * BT) + ((lane >> 3) & 1) * BT + ((lane >> 2) & 1) * (BT / 2);
#pragma unroll
        for (int i = 0; i < BT / 2; ++i) {
            const int idx = first + i;
            const bool is_mean = idx >= 2 * BT;
            const int ch = is_mean ? idx - 2 * BT : idx;           // c*2 + h
            double* dst = (is_mean ? sMeanT : sSST) + (size_t)g * T + (size_t)(cg * BT + (ch >> 1)) * 8 + 2 * (lane & 3) + (ch & 1);
            *dst = pass == 0 ? red[i] : *dst + red[i];
        }
        // further outputs sharing this factorisation: V.z_o from the same accumulators, two outputs per reduction
        // (plane o-1 of sMeanXT holds output o's partials, laid out like sMeanT)
        for (int o = 1; o < p.n_out; o += 2) {
            const bool two = o + 1 < p.n_out;
            double za[NS], zb[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int r = ext[s];
                za[s] = r >= 0 ? __ldg(p.zvec_x[o - 1] + 8 * r + (lane >> 2)) : 0.0;
                zb[s] = (two && r >= 0) ? __ldg(p.zvec_x[o] + 8 * r + (lane >> 2)) : 0.0;
            }
            double red2[4 * BT];
#pragma unroll
            for (int c = 0; c < BT; ++c)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    double ma = 0.0, mb = 0.0;
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        ma = fma(acc[s][c][hh], za[s], ma);
                        mb = fma(acc[s][c][hh], zb[s], mb);
                    }
                    red2[c * 2 + hh] = ma;
                    red2[2 * BT + c * 2 + hh] = mb;
                }
            halving_reduce<4 * BT>(red2, lane);
#pragma unroll
            for (int i = 0; i < BT / 2; ++i) {
                const int idx = first + i;
                const bool second = idx >= 2 * BT;
                const int ch = second ? idx - 2 * BT : idx;
                if (second && !two) continue;
                double* dst = sMeanXT + (size_t)(o - 1 + (second ? 1 : 0)) * p.RG * T + (size_t)g * T +
                              (size_t)(cg * BT + (ch >> 1)) * 8 + 2 * (lane & 3) + (ch & 1);
                *dst = pass == 0 ? red2[i] : *dst + red2[i];
            }
        }
    }
}

// Finalise one row from the per-row-group partials: var = max(k** - |V|^2, 1e-15), l/u = mean -/+ beta sqrt(var) with
// separate multiply and add roundings (NumPy does not contract, gp_opt.py:475-476), S bit (strict >, gp_opt.py:481).
__device__ __forceinline__ void finalize_row(const PostParams& p, const double* __restrict__ sSST, const double* __restrict__ sMeanT,
                                             const double* __restrict__ sMeanXT, int t, int64_t row) {
    const int T = p.T, RG = p.RG;
    double sumsq = 0.0, mu = 0.0;
    for (int g = 0; g < RG; ++g) { sumsq += sSST[(size_t)g * T + t]; mu += sMeanT[(size_t)g * T + t]; }
    double v = p.variance - sumsq;
    v = v > SO_VAR_FLOOR ? v : SO_VAR_FLOOR;
    const double sd = sqrt(v);
    const double bs = __dmul_rn(p.beta, sd);
    const double lo = __dsub_rn(mu, bs), up = __dadd_rn(mu, bs);
    if (p.mean) p.mean[row] = mu;
    if (p.var) p.var[row] = v;
    if (p.Q) {
        double* qp = p.Q + (size_t)row * p.q_stride + p.q_col;
        if ((p.q_stride & 1) == 0 && (p.q_col & 1) == 0) *reinterpret_cast<double2*>(qp) = make_double2(lo, up);
        else { qp[0] = lo; qp[1] = up; }
    }
    uint8_t safe = lo > p.fmin ? 1 : 0;
    for (int o = 1; o < p.n_out; ++o) {                  // same variance, own mean / bounds / threshold
        double mo = 0.0;
        for (int g = 0; g < RG; ++g) mo += sMeanXT[((size_t)(o - 1) * RG + g) * T + t];
        const double lo_o = __dsub_rn(mo, bs), up_o = __dadd_rn(mo, bs);
        if (p.mean_x[o - 1]) p.mean_x[o - 1][row] = mo;
        if (p.var_x[o - 1]) p.var_x[o - 1][row] = v;
        if (p.Q) {
            double* qp = p.Q + (size_t)row * p.q_stride + p.q_col_x[o - 1];
            if ((p.q_stride & 1) == 0 && (p.q_col_x[o - 1] & 1) == 0) *reinterpret_cast<double2*>(qp) = make_double2(lo_o, up_o);
            else { qp[0] = lo_o; qp[1] = up_o; }
        }
        safe &= lo_o > p.fmin_x[o - 1] ? 1 : 0;
    }
    if (p.safe_mode != SO_SAFE_NONE && p.S)
        p.S[row] = p.safe_mode == SO_SAFE_WRITE ? safe : (uint8_t)(p.S[row] & safe);
}

__device__ __forceinline__ void load_tile_rows(const PostParams& p, double* __restrict__ sXt, int64_t tile_local0) {
    const int d = p.d, T = p.T;
    for (int e = threadIdx.x; e < T * d; e += kThreads) {
        const int t = e / d, j = e - t * d;
        int64_t row = tile_local0 + t;
        if (row >= p.M) row = p.M - 1;
        sXt[e] = p.Xstar[(size_t)row * d + j] * p.inv_ls[j];
    }
}

}  // namespace
