// K2, warp-specialised form.  One persistent 384-thread CTA per SM:
//
//   warps 0-3  (producers, 104 registers) : build the k(x*, X) tile of the NEXT candidates k-block by
//                                           k-block, straight into DMMA fragment order, in a shared-memory
//                                           ring of "groups" (4 k-blocks = 32 training points x T rows);
//                                           they also accumulate the mean k.alpha.
//   warps 4-11 (consumers, 200 registers) : the triangular contraction V = L^-1 k on the fp64 tensor
//                                           pipe (DMMA.8x8x4), column sums of squares, epilogue.
//
// Ring slots are handed over with mbarriers (full: 128 producer-lane arrivals, empty: 256 consumer-
// lane arrivals), so the tensor pipe never waits for kernel-row generation and there is no CTA-wide
// barrier in steady state (the r01 profile of the bulk-synchronous version showed 23% of the time in
// the generation phase, profiles/r01_k_posterior_v1_ncu.txt).  Registers are re-split with setmaxnreg.
//
// Grid path: k(x*, x_n) = Pslow[row / F][n] * Pfast[row % F][n] with two product tables built once per
// fit (k_grid_tables2): the F fastest-varying grid rows and the M/F slow combinations.  Two 16-byte
// loads and two multiplies per pair of kernel values instead of two fp64 exp (~21 FMA slots each on the
// one FP64 pipe DMMA also uses).
#pragma once
#include "posterior_core.cuh"

namespace {

constexpr int kProducerWarps = 4;
constexpr int kGroupK = 4;            // k-blocks per ring slot
// Two shapes of the consumer side:
//   CW = 8  consumer warps x 4 block rows x 8 column tiles (128 accumulator registers, 200 regs/thread) -- used when the
//           producers need registers themselves (explicit rows / multiply mode) or N is small;
//   CW = 16 consumer warps x 2 block rows x 8 column tiles (64 accumulator registers, 104 regs/thread) -- TMA mode: four
//           consumer warps per SM sub-partition hide each other's non-DMMA time (A-fragment latency in the one- and
//           two-row tail segments, loop control, reductions): the r01 profile of the 8-warp shape had every consumer warp
//           outside the DMMA stream 44% of the time and the pipe 72% busy.
// setmaxnreg moves registers inside the pool the CTA was LAUNCHED with (threads x launch registers), not the whole SM
// file: 128 x producer + 32 CW x consumer must not exceed it or the last consumer warp spins in USETMAXREG.TRY_ALLOC
// forever (that was the first version's hang with 96/208 on a 384 x 168 pool).
template <int CW> struct WsShape;
template <> struct WsShape<8> {
    static constexpr int kThreads = 384, kLaunchRegs = 168, kProducerRegs = 104, kConsumerRegs = 200, kRows = 4;
};
template <> struct WsShape<16> {
    static constexpr int kThreads = 640, kLaunchRegs = 96, kProducerRegs = 56, kConsumerRegs = 104, kRows = 2;
};
static_assert(128 * WsShape<8>::kProducerRegs + 256 * WsShape<8>::kConsumerRegs <= WsShape<8>::kThreads * WsShape<8>::kLaunchRegs, "");
static_assert(128 * WsShape<16>::kProducerRegs + 512 * WsShape<16>::kConsumerRegs <= WsShape<16>::kThreads * WsShape<16>::kLaunchRegs, "");

// Producer modes.
//   kModeRows : explicit candidate rows; producers evaluate the kernel (distance + exp / Matern profile).
//   kModeGrid : product grid, producers multiply two table entries per kernel value (fallback when the
//               scaled-operand tables of kModeTma would be too large).
//   kModeTma  : product grid, NO fp64 work in tile generation: the slow-axis factor is folded into the A operand
//               (A'(s) = L^-1 diag(Pslow[s]), one packed matrix per slow index, L2-resident), so the B tile is a
//               contiguous 16 KB slice per group of the fragment-ordered fast table and one thread moves it with
//               cp.async.bulk (TMA) straight into the ring; the producer warps only accumulate the mean from the
//               ring.  Motivation: profiles/r01_k_posterior_ws_grid_ncu.txt -- with DMMA saturating the single
//               FP64 pipe, producers that multiply were starved (math-throttle) and consumers waited 28% of the time.
constexpr int kModeRows = 0, kModeGrid = 1, kModeTma = 2;

struct WsParams {
    PostParams p;
    const double* Pfast;              // fast_rows x Npad
    const double* Pslow;              // slow_rows x Npad (carries the signal variance)
    int64_t fast_rows;
    int gpt;                          // groups per tile = ceil(NB / 4)
    int Rg;                           // ring depth in groups
    // kModeTma
    const double2* PfFrag;            // [tile-in-block][group][kb][col tile][lane], zero padded
    const double2* Aprime;            // slow_rows packed scaled operands, a_stride double2 apart
    const double* Wslow;              // slow_rows x Npad: Pslow[s][n] * alpha[n]
    size_t a_stride;
    int tpb;                          // tiles per slow block = ceil(fast_rows / T)
    int64_t first_tile;               // global tile index of this shard's first tile
};

struct WsSmem {
    size_t ring_off, alpha_off, xs_off, xt_off, meanp_off, ss_off, bar_off, total;
};

__host__ __device__ inline WsSmem ws_smem(int NB, int T, int d, int RG, int Rg, bool grid, bool tma = false) {
    WsSmem L;
    const size_t Npad = 8 * (size_t)NB;
    L.ring_off = 0;
    size_t off = (size_t)Rg * kGroupK * 8 * T * sizeof(double);
    L.alpha_off = off; off += Npad * sizeof(double);
    L.xs_off = off; off += grid ? 0 : Npad * d * sizeof(double);
    L.xt_off = off; off += grid ? 0 : 2 * (size_t)T * d * sizeof(double);
    // mean partials: per producer warp (compute modes) or per consumer row group (TMA mode, written by the consumers)
    L.meanp_off = off; off += 2 * (size_t)(tma ? RG : kProducerWarps) * T * sizeof(double);
    L.ss_off = off; off += 2 * (size_t)RG * T * sizeof(double);
    L.bar_off = off; off += (2 * (size_t)Rg + 4) * sizeof(unsigned long long);
    L.total = off;
    return L;
}

// ---- mbarrier / register-split primitives (PTX; SASS: SYNCS.*, USETMAXREG) ----------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int REGS> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS)); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory"); }

// TMA bulk copy global -> shared with byte-count completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// streaming 16-byte load that does not allocate in L1 (A fragments and fast-table rows are used once per tile per SM)
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// ---------------------------------------------------------------- producer
template <int KIND, bool GRID>
__device__ __forceinline__ void ws_producer(const WsParams& wp, double2* sRing, const double* sAlpha, const double* sXs,
                                            double* sXt, double* sMeanP, unsigned long long* full, unsigned long long* empty,
                                            unsigned long long* meanfull, unsigned long long* meanempty, int pw, int lane) {
    const PostParams& p = wp.p;
    const int NB = p.NB, TB = p.TB, T = p.T, N = p.N, d = p.d, Npad = 8 * p.NB;
    const int q = lane & 3, tl = lane >> 2;
    const double2* sA2 = reinterpret_cast<const double2*>(sAlpha);
    const int64_t last = p.row0 + p.M - 1;
    const int ptid = pw * 32 + lane;
    int pslot = 0;                 // ring position of the next group this producer fills
    unsigned pwrap = 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int par = it & 1;
        const int64_t tile_local0 = tile * T;
        unsigned foff[8], soff[8];    // offsets (in double2) of this lane's fast / slow table rows
        double m[8];
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) { m[ct] = 0.0; foff[ct] = 0; soff[ct] = 0; }
        if (GRID) {
#pragma unroll
            for (int ct = 0; ct < 8; ++ct) {
                if (ct < TB) {
                    int64_t row = p.row0 + tile_local0 + ct * 8 + tl;
                    if (row > last) row = last;
                    const int64_t si = row / wp.fast_rows, fi = row - si * wp.fast_rows;
                    foff[ct] = (unsigned)(fi * (Npad / 2)) + q;
                    soff[ct] = (unsigned)(si * (Npad / 2)) + q;
                }
            }
        } else {
            // stage the tile's candidate rows (scaled by 1/lengthscale); producers only
            double* xt = sXt + (size_t)par * T * d;
            for (int e = ptid; e < T * d; e += kProducerWarps * 32) {
                const int t = e / d, j = e - t * d;
                int64_t row = tile_local0 + t;
                if (row >= p.M) row = p.M - 1;
                xt[e] = p.Xstar[(size_t)row * d + j] * p.inv_ls[j];
            }
            named_bar_sync(2, kProducerWarps * 32);
        }
        const double* xt = sXt + (size_t)par * T * d;
        for (int gi = 0; gi < wp.gpt; ++gi) {
            const int slot = pslot;
            mbar_wait(&empty[slot], (pwrap & 1u) ^ 1u);
            if (++pslot == wp.Rg) { pslot = 0; ++pwrap; }
            const int kb = gi * kGroupK + pw;
            if (kb < NB) {
                double2* dst = sRing + ((size_t)slot * kGroupK + pw) * TB * 32 + lane;
                const double2 a = sA2[4 * kb + q];
                if (GRID) {
                    const double2* Pf = reinterpret_cast<const double2*>(wp.Pfast) + 4 * kb;
                    const double2* Ps = reinterpret_cast<const double2*>(wp.Pslow) + 4 * kb;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double2 vf[4], vs[4];
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const int ct = 4 * h + c4;
                            if (ct < TB) { vf[c4] = ldg_stream(Pf + foff[ct]); vs[c4] = __ldg(Ps + soff[ct]); }
                        }
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const int ct = 4 * h + c4;
                            if (ct < TB) {
                                double2 v;
                                v.x = vf[c4].x * vs[c4].x;
                                v.y = vf[c4].y * vs[c4].y;
                                m[ct] = fma(v.x, a.x, m[ct]);
                                m[ct] = fma(v.y, a.y, m[ct]);
                                dst[ct * 32] = v;
                            }
                        }
                    }
                } else {
                    const int n0 = 8 * kb + 2 * q;
                    const double* x0 = sXs + (size_t)n0 * d;
#pragma unroll
                    for (int ct = 0; ct < 8; ++ct)
                        if (ct < TB) {
                            const double* xr = xt + (size_t)(ct * 8 + tl) * d;
                            double r0 = 0.0, r1 = 0.0;
                            for (int j = 0; j < d; ++j) {
                                const double xv = xr[j];
                                const double t0 = xv - x0[j], t1 = xv - x0[d + j];
                                r0 = fma(t0, t0, r0);
                                r1 = fma(t1, t1, r1);
                            }
                            double2 v;
                            v.x = n0 < N ? kernel_of_r2<KIND>(r0, p.variance) : 0.0;
                            v.y = n0 + 1 < N ? kernel_of_r2<KIND>(r1, p.variance) : 0.0;
                            m[ct] = fma(v.x, a.x, m[ct]);
                            m[ct] = fma(v.y, a.y, m[ct]);
                            dst[ct * 32] = v;
                        }
                }
            }
            mbar_arrive(&full[slot]);
        }
        // partial means of this producer warp (its k-blocks only); the epilogue adds the four partials in order
        mbar_wait(&meanempty[par], (((unsigned)(it >> 1)) & 1u) ^ 1u);
#pragma unroll
        for (int ct = 0; ct < 8; ++ct)
            if (ct < TB) {
                double v = m[ct];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (q == 0) sMeanP[((size_t)par * kProducerWarps + pw) * T + ct * 8 + tl] = v;
            }
        mbar_arrive(&meanfull[par]);
    }
}

// ---------------------------------------------------------------- producer group, kModeTma
// warp 0 (one lane): walks the tiles ahead of the consumers and moves each 16 KB group of the fragment-ordered fast table
// into the ring with one bulk copy.  warps 1-2: epilogue (one row per lane).  warp 3: idle (the register split needs a
// full warpgroup).
__device__ __forceinline__ void ws_tma_issuer(const WsParams& wp, double2* sRing, unsigned long long* full, unsigned long long* empty) {
    const PostParams& p = wp.p;
    const unsigned group_bytes = (unsigned)(kGroupK * p.TB * 512);
    const size_t group_elems = (size_t)kGroupK * p.TB * 32;
    int pslot = 0;
    unsigned pwrap = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int64_t gt = wp.first_tile + tile;
        const int j = (int)(gt % wp.tpb);
        const double2* src = wp.PfFrag + (size_t)j * wp.gpt * group_elems;
        for (int gi = 0; gi < wp.gpt; ++gi) {
            mbar_wait(&empty[pslot], (pwrap & 1u) ^ 1u);
            mbar_expect_tx(&full[pslot], group_bytes);
            tma_bulk_g2s(sRing + (size_t)pslot * group_elems, src + (size_t)gi * group_elems, group_bytes, &full[pslot]);
            if (++pslot == wp.Rg) { pslot = 0; ++pwrap; }
        }
    }
}

// Finalises a tile from the consumers' partials: |V|^2 summed over the RG row groups, mean summed over the NB k-blocks
// (both in fixed order => bit-reproducible), then var, l/u (separate multiply and add roundings like NumPy) and the S bit.
__device__ __forceinline__ void ws_epilogue_warp(const WsParams& wp, const double* sMeanG, const double* sSS,
                                                 unsigned long long* tilefull, unsigned long long* tileempty, int elane) {
    const PostParams& p = wp.p;
    const int T = p.T, RG = p.RG;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int par = it & 1;
        const int64_t gt = wp.first_tile + tile;
        const int64_t si = gt / wp.tpb;
        const int j = (int)(gt - si * wp.tpb);
        const int64_t tile_row0 = si * wp.fast_rows + (int64_t)j * T - p.row0;
        const int64_t left = wp.fast_rows - (int64_t)j * T;
        const int valid_cols = left < T ? (int)left : T;
        mbar_wait(&tilefull[par], ((unsigned)(it >> 1)) & 1u);
        const double* ssp = sSS + (size_t)par * RG * T;
        const double* mgp = sMeanG + (size_t)par * RG * T;
        for (int t = elane; t < T; t += 64) {
            const int64_t row = tile_row0 + t;
            if (t >= valid_cols || row < 0 || row >= p.M) continue;
            double sumsq = 0.0, mu = 0.0;
            for (int g = 0; g < RG; ++g) { sumsq += ssp[(size_t)g * T + t]; mu += mgp[(size_t)g * T + t]; }
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
            if (p.safe_mode != SO_SAFE_NONE && p.S) {
                const uint8_t safe = lo > p.fmin ? 1 : 0;
                p.S[row] = p.safe_mode == SO_SAFE_WRITE ? safe : (uint8_t)(p.S[row] & safe);
            }
        }
        mbar_arrive(&tileempty[par]);
    }
}

// ---------------------------------------------------------------- consumer
// 32-bit shared-memory addresses and 32-bit fragment offsets keep the scalar state of the contraction
// loop small: the accumulators alone take 128 of the 200 registers.
__device__ __forceinline__ void mbar_arrive_u32(unsigned bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

struct RingCursor {           // position of group 0 of the current tile in the ring
    int slot0;
    unsigned wrap0;
    int Rg;
    unsigned full0, empty0;   // shared addresses of full[0], empty[0]
    unsigned ring0;           // shared address of the ring (+ this lane's fragment offset)
    unsigned group_bytes, kb_bytes;
    __device__ __forceinline__ void locate(int gi, int& slot, unsigned& parity) const {
        int s = slot0 + gi;
        unsigned w = wrap0;
        while (s >= Rg) { s -= Rg; ++w; }
        slot = s;
        parity = w & 1u;
    }
};

struct MeanSink {              // TMA mode: where a consumer warp accumulates the mean contribution of its own k-blocks
    const double2* w2;         // Wslow[s] as double2, + (lane & 3)
    unsigned base;             // shared address of sMeanG[par][g][this warp's first column + lane/4]
    bool first;                // no contribution stored yet in this tile
};

template <int BT, int ROWS, int FIRST, bool MEAN>
__device__ __forceinline__ void ws_segment(double (&acc)[ROWS][BT][2], double2 (&a)[ROWS], const double2* __restrict__ Afrag,
                                           const unsigned (&abase)[ROWS], const RingCursor& rc, int kb_lo, int kb_hi,
                                           bool release, int NB, int& next_release, MeanSink& ms, int lane) {
    int slot = 0;
    unsigned parity = 0;
    for (int kb = kb_lo; kb <= kb_hi; ++kb) {
        const int gi = kb >> 2;
        if ((kb & 3) == 0 || kb == kb_lo) {
            rc.locate(gi, slot, parity);
            mbar_wait_u32(rc.full0 + 8u * slot, parity);
        }
        double2 an[ROWS];
#pragma unroll
        for (int s = FIRST; s < ROWS; ++s) an[s] = ldg_stream(Afrag + abase[s] + (unsigned)(kb + 1) * 32u);
        const unsigned bp = rc.ring0 + (unsigned)slot * rc.group_bytes + (unsigned)(kb & 3) * rc.kb_bytes;
#pragma unroll
        for (int c = 0; c < BT; ++c) {
            const double2 b = lds_f64x2(bp + c * 512u);
#pragma unroll
            for (int s = FIRST; s < ROWS; ++s) {
                dmma884(acc[s][c][0], acc[s][c][1], a[s].x, b.x);
                dmma884(acc[s][c][0], acc[s][c][1], a[s].y, b.y);
            }
        }
#pragma unroll
        for (int s = FIRST; s < ROWS; ++s) a[s] = an[s];
        if (MEAN && kb == kb_hi) {
            // kb_hi is the block row that slot FIRST owns: every k-block index is owned by exactly one (warp, pass, slot),
            // so the mean k.w is assembled from per-k-block partials without any cross-warp accumulation
            const double2 wv = __ldg(ms.w2 + 4 * kb);
#pragma unroll
            for (int c = 0; c < BT; ++c) {
                const double2 b = lds_f64x2(bp + c * 512u);
                double v = fma(b.y, wv.y, b.x * wv.x);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if ((lane & 3) == 0) {
                    const unsigned addr = ms.base + c * 64u;
                    if (!ms.first) {
                        double prev;
                        asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(prev) : "r"(addr) : "memory");
                        v += prev;
                    }
                    asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
                }
            }
            ms.first = false;
        }
        if (release && ((kb & 3) == 3 || kb == NB - 1)) {
            mbar_arrive_u32(rc.empty0 + 8u * slot);
            next_release = gi + 1;
        }
    }
}

template <int BT, int CW, bool TMA>
__device__ __forceinline__ void ws_consumer(const WsParams& wp, const double2* sRing, double* sMeanP, double* sSS,
                                            unsigned long long* full, unsigned long long* empty, unsigned long long* meanfull,
                                            unsigned long long* meanempty, int cw, int lane) {
    const PostParams& p = wp.p;
    const int RG = p.RG, NB = p.NB, T = p.T;
    const int g = cw % RG, cg = cw / RG;
    const double2* Afrag = p.Afrag + lane;
    const int ctid = cw * 32 + lane;
    RingCursor rc;
    rc.Rg = wp.Rg;
    rc.full0 = smem_u32(full);
    rc.empty0 = smem_u32(empty);
    rc.ring0 = smem_u32(sRing) + (unsigned)(cg * BT * 32 + lane) * 16u;
    rc.kb_bytes = (unsigned)p.TB * 512u;
    rc.group_bytes = rc.kb_bytes * kGroupK;
    rc.slot0 = 0;
    rc.wrap0 = 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int par = it & 1;
        double* sSST = sSS + (size_t)par * RG * T;
        int next_release = 0;
        int64_t tile_row0 = tile * T;         // local row of column 0 of this tile
        int valid_cols = T;
        MeanSink ms;
        ms.w2 = nullptr; ms.base = 0; ms.first = true;
        if (TMA) {
            const int64_t gt = wp.first_tile + tile;
            const int64_t si = gt / wp.tpb;
            const int j = (int)(gt - si * wp.tpb);
            Afrag = wp.Aprime + (size_t)si * wp.a_stride + lane;
            tile_row0 = si * wp.fast_rows + (int64_t)j * T - p.row0;
            const int64_t left = wp.fast_rows - (int64_t)j * T;
            valid_cols = left < T ? (int)left : T;
            // the epilogue warp must have consumed the buffers of tile it-2 before they are written again
            mbar_wait(&meanempty[par], (((unsigned)(it >> 1)) & 1u) ^ 1u);
            ms.w2 = reinterpret_cast<const double2*>(wp.Wslow + (size_t)si * 8 * NB) + (lane & 3);
            ms.base = smem_u32(sMeanP + ((size_t)par * RG + g) * T + (size_t)(cg * BT) * 8 + (lane >> 2));
            ms.first = true;
        }
        for (int pass = 0; pass < p.npass; ++pass) {
            const bool release = pass == p.npass - 1;
            constexpr int ROWS = WsShape<CW>::kRows;
            const int base = ROWS * RG * pass;
            // ROWS = 4: block rows {g, 2RG-1-g, 2RG+g, 4RG-1-g}; ROWS = 2: {g, 2RG-1-g} -- pairings that equalise the
            // triangular work.  Rows are ascending; the ones beyond NB (inactive) form a suffix.  Slots are ordered by K
            // extent, inactive slots (extent -1) first, so that "slots FIRST..ROWS-1 active" holds in every segment.
            const int r0 = base + g, r1 = base + 2 * RG - 1 - g;
            const int r2 = ROWS == 4 ? base + 2 * RG + g : (1 << 28), r3 = ROWS == 4 ? base + 4 * RG - 1 - g : (1 << 28);
            const int na = (r0 < NB) + (r1 < NB) + (r2 < NB) + (r3 < NB);
            int ext[ROWS];
            unsigned abase[ROWS];
#pragma unroll
            for (int s = 0; s < ROWS; ++s) {
                const int src = s - (ROWS - na);
                const int r = src >= 0 ? pick4(r0, r1, r2, r3, src) : -1;
                ext[s] = r;
                abase[s] = r >= 0 ? (unsigned)(r * (r + 1) / 2) * 32u : 0u;
            }
            double acc[ROWS][BT][2];
#pragma unroll
            for (int s = 0; s < ROWS; ++s)
#pragma unroll
                for (int c = 0; c < BT; ++c) { acc[s][c][0] = 0.0; acc[s][c][1] = 0.0; }
            double2 a[ROWS];
#pragma unroll
            for (int s = 0; s < ROWS; ++s) a[s] = ldg_stream(Afrag + abase[s]);
            ws_segment<BT, ROWS, 0, TMA>(acc, a, Afrag, abase, rc, 0, ext[0], release, NB, next_release, ms, lane);
            ws_segment<BT, ROWS, 1, TMA>(acc, a, Afrag, abase, rc, ext[0] + 1, ext[1], release, NB, next_release, ms, lane);
            if (ROWS == 4) {
                ws_segment<BT, ROWS, (ROWS == 4 ? 2 : 1), TMA>(acc, a, Afrag, abase, rc, ext[1] + 1, ext[ROWS == 4 ? 2 : 1], release, NB, next_release, ms, lane);
                ws_segment<BT, ROWS, (ROWS == 4 ? 3 : 1), TMA>(acc, a, Afrag, abase, rc, ext[ROWS == 4 ? 2 : 1] + 1, ext[ROWS - 1], release, NB, next_release, ms, lane);
            }
            // column sums of squares of this pass: over the 4 slots, then over the 8 rows of a block
            // (lane bits 2..4, fixed tree => deterministic); accumulated across passes in this warp's own
            // shared-memory slots so that no registers stay live across the contraction loop
#pragma unroll
            for (int c = 0; c < BT; ++c) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int s = 0; s < WsShape<CW>::kRows; ++s) {
                    s0 = fma(acc[s][c][0], acc[s][c][0], s0);
                    s1 = fma(acc[s][c][1], acc[s][c][1], s1);
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                if (lane < 4) {
                    double2* dst = reinterpret_cast<double2*>(sSST + (size_t)g * T + (cg * BT + c) * 8 + 2 * lane);
                    if (pass == 0) *dst = make_double2(s0, s1);
                    else { const double2 prev = *dst; *dst = make_double2(prev.x + s0, prev.y + s1); }
                }
            }
        }
        // groups this warp never touched in its last pass still have to be handed back
        for (int gi = next_release; gi < wp.gpt; ++gi) {
            int slot;
            unsigned parity;
            rc.locate(gi, slot, parity);
            mbar_wait_u32(rc.full0 + 8u * slot, parity);
            mbar_arrive_u32(rc.empty0 + 8u * slot);
        }
        // advance the ring cursor by one tile
        rc.slot0 += wp.gpt;
        while (rc.slot0 >= rc.Rg) { rc.slot0 -= rc.Rg; ++rc.wrap0; }
        if (TMA) {
            // no consumer-side barrier or epilogue: hand the partial sums to the epilogue warp and move on
            mbar_arrive(&meanfull[par]);
            continue;
        }
        named_bar_sync(1, CW * 32);
        if (ctid < T) {
            mbar_wait(&meanfull[par], ((unsigned)(it >> 1)) & 1u);
            const int64_t row = tile_row0 + ctid;
            if (ctid < valid_cols && row >= 0 && row < p.M) {
                double sumsq = 0.0;
                for (int gg = 0; gg < RG; ++gg) sumsq += sSST[(size_t)gg * T + ctid];
                const double* mp = sMeanP + (size_t)par * kProducerWarps * T + ctid;
                const double mu = ((mp[0] + mp[T]) + mp[2 * T]) + mp[3 * T];
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
                if (p.safe_mode != SO_SAFE_NONE && p.S) {
                    const uint8_t safe = lo > p.fmin ? 1 : 0;
                    p.S[row] = p.safe_mode == SO_SAFE_WRITE ? safe : (uint8_t)(p.S[row] & safe);
                }
            }
        }
        mbar_arrive(&meanempty[par]);
    }
}

template <int BT, int KIND, int MODE, int CW>
__global__ void __launch_bounds__(WsShape<CW>::kThreads, 1) k_posterior_ws(const __grid_constant__ WsParams wp) {
    constexpr int kWsThreads = WsShape<CW>::kThreads;
    constexpr bool GRID = MODE != kModeRows;
    constexpr bool TMA = MODE == kModeTma;
    const PostParams& p = wp.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const WsSmem L = ws_smem(p.NB, p.T, p.d, p.RG, wp.Rg, GRID, TMA);
    double2* sRing = reinterpret_cast<double2*>(smem_raw + L.ring_off);
    double* sAlpha = reinterpret_cast<double*>(smem_raw + L.alpha_off);
    double* sXs = reinterpret_cast<double*>(smem_raw + L.xs_off);
    double* sXt = reinterpret_cast<double*>(smem_raw + L.xt_off);
    double* sMeanP = reinterpret_cast<double*>(smem_raw + L.meanp_off);
    double* sSS = reinterpret_cast<double*>(smem_raw + L.ss_off);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + L.bar_off);
    unsigned long long* full = bars;
    unsigned long long* empty = bars + wp.Rg;
    unsigned long long* meanfull = bars + 2 * wp.Rg;
    unsigned long long* meanempty = meanfull + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Npad = 8 * p.NB;

    for (int i = threadIdx.x; i < Npad; i += kWsThreads) sAlpha[i] = p.alpha[i];
    if (!GRID)
        for (int i = threadIdx.x; i < Npad * p.d; i += kWsThreads) sXs[i] = p.Xs[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < wp.Rg; ++s) {
            // kModeTma: one expect_tx arrival + the TMA byte count complete `full`; the producer warps also read the
            // slot (mean) and therefore take part in `empty`
            // kModeTma: one expect_tx arrival + the TMA byte count complete `full`
            mbar_init(&full[s], TMA ? 1 : kProducerWarps * 32);
            mbar_init(&empty[s], CW * 32);
        }
        for (int s = 0; s < 2; ++s) {
            // compute modes: producers publish the mean (meanfull), consumers hand the buffer back (meanempty);
            // kModeTma: consumers publish their partials (tilefull = meanfull), the epilogue warp hands them back
            mbar_init(&meanfull[s], TMA ? CW * 32 : kProducerWarps * 32);
            mbar_init(&meanempty[s], TMA ? 64 : CW * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp < kProducerWarps) {
        reg_dealloc<WsShape<CW>::kProducerRegs>();
        if (TMA) {
            if (warp == 0 && lane == 0) ws_tma_issuer(wp, sRing, full, empty);
            else if (warp == 1 || warp == 2) ws_epilogue_warp(wp, sMeanP, sSS, meanfull, meanempty, (warp - 1) * 32 + lane);
        } else ws_producer<KIND, GRID>(wp, sRing, sAlpha, sXs, sXt, sMeanP, full, empty, meanfull, meanempty, warp, lane);
    } else {
        reg_alloc<WsShape<CW>::kConsumerRegs>();
        ws_consumer<BT, CW, TMA>(wp, sRing, sMeanP, sSS, full, empty, meanfull, meanempty, warp - kProducerWarps, lane);
    }
}

// Product tables of the grid path.  Row r of the fast table is the product over the fast axes of
// exp(-0.5 ((x_j - X_nj)/l_j)^2) for grid row r (< fast_rows); row s of the slow table the same over the
// slow axes for grid row s * fast_rows, times the signal variance.  Padding columns n >= N are zero.
struct TableSpec {
    int d;
    int n[kGridMaxDim];
    int off[kGridMaxDim];
    int64_t stride[kGridMaxDim];
    int in_fast[kGridMaxDim];
    int64_t fast_rows, slow_rows;
};

__global__ void k_grid_tables2(TableSpec ts, const double* __restrict__ axis, const double* __restrict__ Xs,
                               double* __restrict__ Pfast, double* __restrict__ Pslow, int N, int Npad, int d,
                               double variance, const double* __restrict__ inv_ls_d) {
    const int64_t r = blockIdx.x;             // one table row per block
    const bool fast = r < ts.fast_rows;
    const int64_t tr = fast ? r : r - ts.fast_rows;
    const int64_t grow = fast ? tr : tr * ts.fast_rows;
    for (int n = threadIdx.x; n < Npad; n += blockDim.x) {
        double v = 0.0;
        if (n < N) {
            v = fast ? 1.0 : variance;
            for (int j = 0; j < ts.d; ++j) {
                if ((ts.in_fast[j] != 0) != fast) continue;
                const int idx = (int)((grow / ts.stride[j]) % ts.n[j]);
                const double t = axis[ts.off[j] + idx] * inv_ls_d[j] - Xs[(size_t)n * d + j];
                v *= exp(-0.5 * (t * t));
            }
        }
        (fast ? Pfast : Pslow)[(size_t)tr * Npad + n] = v;
    }
}

// kModeTma tables.  PfFrag: the fast table re-ordered so that the 16 KB a ring slot needs for (tile j of a slow block,
// group g) are contiguous: [j][g][kb in group][col tile][lane] double2, lane l -> rows j*T + 8ct + l/4, training
// points 8kb + 2(l%4) + {0,1}; zero beyond fast_rows / N.
__global__ void k_pffrag(const double* __restrict__ Pfast, double2* __restrict__ PfFrag, int64_t fast_rows, int N, int Npad,
                         int T, int TB, int gpt, int tpb) {
    const size_t total = (size_t)tpb * gpt * kGroupK * TB * 32;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(e & 31);
        size_t r = e >> 5;
        const int ct = (int)(r % TB); r /= TB;
        const int kbi = (int)(r % kGroupK); r /= kGroupK;
        const int g = (int)(r % gpt);
        const int j = (int)(r / gpt);
        const int64_t row = (int64_t)j * T + ct * 8 + (lane >> 2);
        const int n0 = 8 * (g * kGroupK + kbi) + 2 * (lane & 3);
        double2 v = make_double2(0.0, 0.0);
        if (row < fast_rows) {
            if (n0 < N) v.x = Pfast[(size_t)row * Npad + n0];
            if (n0 + 1 < N) v.y = Pfast[(size_t)row * Npad + n0 + 1];
        }
        PfFrag[e] = v;
    }
}

// A'(s) = L^-1 diag(Pslow[s]) in the packed fragment order of Afrag, and Wslow[s] = Pslow[s] * alpha.
__global__ void k_aprime(const double2* __restrict__ Afrag, const double* __restrict__ Pslow, const double* __restrict__ alpha,
                         double2* __restrict__ Aprime, double* __restrict__ Wslow, int NB, size_t a_stride) {
    const int64_t si = blockIdx.y;
    const int Npad = 8 * NB;
    const double* ps = Pslow + (size_t)si * Npad;
    const size_t nfrag = tri_blocks(NB) * 32;
    double2* dst = Aprime + (size_t)si * a_stride;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < a_stride; e += (size_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (e < nfrag) {
            const size_t blk = e >> 5;
            const int lane = (int)(e & 31);
            // block index -> (i, kb) with blk = i(i+1)/2 + kb
            int i = (int)((sqrt(8.0 * (double)blk + 1.0) - 1.0) * 0.5);
            while ((size_t)(i + 1) * (i + 2) / 2 <= blk) ++i;
            while ((size_t)i * (i + 1) / 2 > blk) --i;
            const int kb = (int)(blk - (size_t)i * (i + 1) / 2);
            const int c0 = 8 * kb + 2 * (lane & 3);
            const double2 a = Afrag[e];
            v.x = a.x * ps[c0];
            v.y = a.y * ps[c0 + 1];
        }
        dst[e] = v;
    }
    if (blockIdx.x == 0)
        for (int n = threadIdx.x; n < Npad; n += blockDim.x) Wslow[(size_t)si * Npad + n] = ps[n] * alpha[n];
}

}  // namespace
