// K2, grid path, final form ("tma"): the bulk-synchronous contraction loop of posterior.cu -- whose MMA phase was
// measured at 92% of the DMMA peak -- with the kernel-row generation phase removed altogether:
//
//   * the slow-axis factor of the separable RBF kernel is folded into the A operand, A'(s) = L^-1 diag(Pslow[s])
//     (one packed 270 KB matrix per slow index, L2-resident while its 2500 rows are being processed), so the B tile of
//     a row block is a plain contiguous slice of the fragment-ordered fast table;
//   * one thread moves the NEXT tile's slice (T = 48 rows x N: 96 KB) with cp.async.bulk (TMA) into the second half of a
//     double buffer while all eight warps contract the current one; completion is an mbarrier transaction count;
//   * the mean k.(Pslow*alpha) is picked up inside the contraction loop: warp (g, pass, slot) adds the contribution of
//     k-block kb == its own block row, so every k-block is counted exactly once with ~1% extra fp64 work.
//
// Per tile: one mbarrier wait, the contraction, one __syncthreads, a 48-thread epilogue.  The warp-specialised
// producer/consumer variants (posterior_ws.cuh) stay available for comparison (SO_K2_VARIANT=ws); they lose ~20% in
// their consumer loop to per-group barrier traffic and the smaller register budget
// (profiles/r01_k2_variants.md).
#pragma once
#include "posterior_ws.cuh"

namespace {

struct TmaParams {
    PostParams p;
    const double2* PfFrag;            // [tile in slow block][k-block][col tile][lane] double2, zero padded
    const double2* Aprime;            // slow_rows packed scaled operands, a_stride double2 apart
    const double* zvec;               // Npad: z = L^-1 y
    size_t a_stride;
    int64_t fast_rows;
    int64_t first_tile;
    int tpb;                          // tiles per slow block
    int kb_pad;                       // k-blocks per tile in the table (4 * ceil(NB / 4))
};

struct TmaSmem { size_t buf_bytes, ss_off, mean_off, bar_off, total; };

__host__ __device__ inline TmaSmem tma_smem(int kb_pad, int TB, int RG, int T) {
    TmaSmem L;
    L.buf_bytes = (size_t)kb_pad * TB * 512;
    L.ss_off = 2 * L.buf_bytes;
    L.mean_off = L.ss_off + 2 * (size_t)RG * T * sizeof(double);
    L.bar_off = L.mean_off + 2 * (size_t)RG * T * sizeof(double);
    L.total = L.bar_off + 64;
    return L;
}

// Recursive-halving reduction of K values per lane over the 8 lanes that hold the rows of one 8x8 block (lane bits
// 4,3,2): after the three stages every lane owns K/8 fully reduced values, at 7K/8 shuffle+add pairs per lane instead of
// 3K for a butterfly.  Scalar fp64 instructions are precious here: they share the one FP64 pipe with DMMA and are
// served behind it (profiles/r01_k2_variants.md).  Lane (b4,b3,b2) ends up with original indices
// b4*K/2 + b3*K/4 + b2*K/8 + [0, K/8).
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

template <int BT, int FIRST>
__device__ __forceinline__ void tma_segment(double (&acc)[4][BT][2], double2 (&a)[4], const double2* __restrict__ Afrag,
                                            const size_t (&abase)[4], const double2* __restrict__ sB, int TB, int kb_lo,
                                            int kb_hi) {
    for (int kb = kb_lo; kb <= kb_hi; ++kb) {
        double2 an[4];
#pragma unroll
        for (int s = FIRST; s < 4; ++s) an[s] = __ldg(Afrag + abase[s] + (size_t)(kb + 1) * 32);
        const double2* bp = sB + (size_t)kb * TB * 32;
#pragma unroll
        for (int c = 0; c < BT; ++c) {
            const double2 b = bp[c * 32];
#pragma unroll
            for (int s = FIRST; s < 4; ++s) {
                dmma884(acc[s][c][0], acc[s][c][1], a[s].x, b.x);
                dmma884(acc[s][c][0], acc[s][c][1], a[s].y, b.y);
            }
        }
#pragma unroll
        for (int s = FIRST; s < 4; ++s) a[s] = an[s];
    }
}

template <int BT>
__global__ void __launch_bounds__(kThreads, 1) k_posterior_tma(const __grid_constant__ TmaParams tp) {
    const PostParams& p = tp.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TmaSmem L = tma_smem(tp.kb_pad, p.TB, p.RG, p.T);
    double2* sBuf = reinterpret_cast<double2*>(smem_raw);
    double* sSS = reinterpret_cast<double*>(smem_raw + L.ss_off);
    double* sMeanG = reinterpret_cast<double*>(smem_raw + L.mean_off);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + L.bar_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RG = p.RG, NB = p.NB, TB = p.TB, T = p.T, Npad = 8 * p.NB;
    const int g = warp % RG, cg = warp / RG;
    const size_t buf_elems = L.buf_bytes / sizeof(double2);
    const unsigned chunk_bytes = (unsigned)(kGroupK * TB * 512);
    const int nchunks = tp.kb_pad / kGroupK;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int b) {
        const int64_t gt = tp.first_tile + tile;
        const int j = (int)(gt % tp.tpb);
        const double2* src = tp.PfFrag + (size_t)j * buf_elems;
        double2* dst = sBuf + (size_t)b * buf_elems;
        mbar_expect_tx(&full[b], (unsigned)L.buf_bytes);
        for (int c = 0; c < nchunks; ++c)
            tma_bulk_g2s(dst + (size_t)c * (chunk_bytes / 16), src + (size_t)c * (chunk_bytes / 16), chunk_bytes, &full[b]);
    };

    if (threadIdx.x == 0 && (int64_t)blockIdx.x < p.ntiles) issue(blockIdx.x, 0);

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const int64_t next = tile + gridDim.x;
        // buffer b^1 was read by the contraction of the previous tile, which every thread left through the
        // __syncthreads below before this point
        if (threadIdx.x == 0 && next < p.ntiles) issue(next, b ^ 1);

        const int64_t gt = tp.first_tile + tile;
        const int64_t si = gt / tp.tpb;
        const int j = (int)(gt - si * tp.tpb);
        const double2* Afrag = tp.Aprime + (size_t)si * tp.a_stride + lane;
        const double2* sB = sBuf + (size_t)b * buf_elems + (size_t)(cg * BT) * 32 + lane;
        double* sSST = sSS + (size_t)b * RG * T;
        double* sMeanT = sMeanG + (size_t)b * RG * T;

        mbar_wait(&full[b], ((unsigned)(it >> 1)) & 1u);

        for (int pass = 0; pass < p.npass; ++pass) {
            const int base = 4 * RG * pass;
            const int r0 = base + g, r1 = base + 2 * RG - 1 - g, r2 = base + 2 * RG + g, r3 = base + 4 * RG - 1 - g;
            const int na = (r0 < NB) + (r1 < NB) + (r2 < NB) + (r3 < NB);
            int ext[4];
            size_t abase[4];
            double zs[4];              // z = L^-1 y at this lane's row of each slot's block (0 for inactive slots)
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int src = s - (4 - na);
                const int r = src >= 0 ? pick4(r0, r1, r2, r3, src) : -1;
                ext[s] = r;
                abase[s] = r >= 0 ? (size_t)r * (r + 1) / 2 * 32 : 0;
                zs[s] = r >= 0 ? __ldg(tp.zvec + 8 * r + (lane >> 2)) : 0.0;
            }
            double acc[4][BT][2];
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int c = 0; c < BT; ++c) { acc[s][c][0] = 0.0; acc[s][c][1] = 0.0; }
            double2 a[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) a[s] = __ldg(Afrag + abase[s]);
            tma_segment<BT, 0>(acc, a, Afrag, abase, sB, TB, 0, ext[0]);
            tma_segment<BT, 1>(acc, a, Afrag, abase, sB, TB, ext[0] + 1, ext[1]);
            tma_segment<BT, 2>(acc, a, Afrag, abase, sB, TB, ext[1] + 1, ext[2]);
            tma_segment<BT, 3>(acc, a, Afrag, abase, sB, TB, ext[2] + 1, ext[3]);
            // |V|^2 and V.z of this warp's rows: red[c*2+h] = sum of squares, red[2BT + c*2+h] = mean share, for column
            // 8c + 2(lane%4) + h of this warp's column group; then across the 8 row lanes by recursive halving
            double red[4 * BT];
#pragma unroll
            for (int c = 0; c < BT; ++c)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    double q2 = 0.0, mz = 0.0;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
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
        }
        __syncthreads();

        // epilogue: tiles are aligned to the slow blocks of the grid, rows outside this rank's shard are masked
        const int64_t tile_row0 = si * tp.fast_rows + (int64_t)j * T - p.row0;
        const int64_t left = tp.fast_rows - (int64_t)j * T;
        const int valid_cols = left < T ? (int)left : T;
        for (int t = threadIdx.x; t < T; t += kThreads) {
            const int64_t row = tile_row0 + t;
            if (t >= valid_cols || row < 0 || row >= p.M) continue;
            double sumsq = 0.0, mu = 0.0;
            const double* mg = sMeanG + (size_t)b * RG * T;
            for (int gg = 0; gg < RG; ++gg) { sumsq += sSST[(size_t)gg * T + t]; mu += mg[(size_t)gg * T + t]; }
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
}

}  // namespace
