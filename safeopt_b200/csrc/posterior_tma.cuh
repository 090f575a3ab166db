// K2, grid path ("tma" kernel): the contraction loop of posterior_core.cuh with the kernel-row generation phase removed.
//
//   * The slow-axis factor of the separable RBF kernel is folded into the A operand, A'(s) = L^-1 diag(Pslow[s]) -- one
//     packed 270 KB matrix per slow index (N = 256), L2-resident while its block of fast rows is processed -- so the B
//     tile of a row block is a plain contiguous slice of the fragment-ordered fast table: V = L^-1 k = A'(s) Pfast.
//   * One thread moves the NEXT tile's slice (T = 48 rows x N: 96 KB) with cp.async.bulk (TMA, SASS UBLKCP) into the
//     other half of a shared-memory double buffer while all eight warps contract the current one; completion is an
//     mbarrier transaction count.  No fp64 instruction is spent on generating kernel rows: DMMA saturates the one FP64
//     pipe and scalar fp64 work is served behind it (profiles/r01_k2_variants.md).
//   * Mean and variance both come out of the accumulators: |V|^2 and V.z with z = L^-1 y.
//
// Per tile: one mbarrier wait, the contraction, one __syncthreads, a T-thread epilogue.  Tiles are aligned to the slow
// blocks of the grid (tile = (row / F) * tpb + (row % F) / T); rows outside the rank's shard are masked in the epilogue.
// Algorithmic HBM bytes per row: 33 written (mean, var, l, u, S); the A' table adds slow_rows * 270 KB of reads per
// launch (676 MB at C4 = 0.1 ms at 6.5 TB/s against a 14 ms kernel).
#pragma once
#include "posterior_core.cuh"

namespace {

constexpr int kChunkK = 4;            // k-blocks per TMA bulk copy

// ---- mbarrier / TMA primitives (PTX; SASS: SYNCS.*, UBLKCP) -----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct TmaParams {
    PostParams p;
    const double2* PfFrag;            // [tile in slow block][k-block][col tile][lane] double2, zero padded
    const double2* Aprime;            // slow_rows packed scaled operands, a_stride double2 apart
    size_t a_stride;
    int64_t s0;                       // slow index of the first operand resident in Aprime
    int64_t fast_rows;
    int64_t first_tile;
    int tpb;                          // tiles per slow block
    int kb_pad;                       // k-blocks per tile in the table (= NB: the last bulk copy of a tile may be short)
};

struct TmaSmem { size_t buf_bytes, ss_off, mean_off, meanx_off, bar_off, total; };

__host__ __device__ inline TmaSmem tma_smem(int kb_pad, int TB, int RG, int T, int n_extra = 0) {
    TmaSmem L;
    L.buf_bytes = (size_t)kb_pad * TB * 512;
    L.ss_off = 2 * L.buf_bytes;
    L.mean_off = L.ss_off + 2 * (size_t)RG * T * sizeof(double);
    L.meanx_off = L.mean_off + 2 * (size_t)RG * T * sizeof(double);
    L.bar_off = L.meanx_off + 2 * (size_t)n_extra * RG * T * sizeof(double);
    L.total = L.bar_off + 64;
    return L;
}

// WARPS = 8: four block rows per warp and pass (default); WARPS = 16: two (contract_tile<BT, 2>), i.e. four warps per
// scheduler instead of two -- kept for A/B measurements (see tma_warps() in posterior.cu).
// NS2X = true (8 warps, TWO block rows per warp, two CTAs per SM): for fits with at most 16 block rows (N <= 128) a tile has so
// few k-steps that its tile-end reduction, barrier and epilogue are a sizeable share of it; with half the accumulators the kernel
// needs ~110 registers and ~100 KB of shared memory, so two CTAs share an SM and one contracts while the other finishes a tile.
// NSV = 6 (8 warps, SIX block rows per warp, 32-row tiles): for 281 <= N <= 384 the 48-row tile no longer fits twice; with four rows
// per warp the 32-row tile needs two passes over its B tile (the second with one or two rows per warp); six rows per warp keep
// the 24 accumulator pairs of the N = 256 configuration and cover up to 48 block rows in ONE pass.
template <int BT, int WARPS, int NSV = 0>
__global__ void __launch_bounds__(WARPS * 32, NSV == 2 ? 2 : 1) k_posterior_tma(const __grid_constant__ TmaParams tp) {
    constexpr bool NS2X = NSV == 2;
    constexpr int NS = NSV != 0 ? NSV : (WARPS == 16 ? 2 : 4);
    (void)NS2X;
    constexpr int kCtaThreads = WARPS * 32;
    const PostParams& p = tp.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n_extra = p.n_out - 1;
    const TmaSmem L = tma_smem(tp.kb_pad, p.TB, p.RG, p.T, n_extra);
    double2* sBuf = reinterpret_cast<double2*>(smem_raw);
    double* sSS = reinterpret_cast<double*>(smem_raw + L.ss_off);
    double* sMean = reinterpret_cast<double*>(smem_raw + L.mean_off);
    double* sMeanX = reinterpret_cast<double*>(smem_raw + L.meanx_off);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + L.bar_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RG = p.RG, T = p.T, TB = p.TB;
    const int g = warp % RG, cg = warp / RG;
    const size_t buf_elems = L.buf_bytes / sizeof(double2);
    const unsigned chunk_bytes = (unsigned)(kChunkK * TB * 512);
    const int nchunks = (tp.kb_pad + kChunkK - 1) / kChunkK;
    const unsigned last_bytes = (unsigned)((tp.kb_pad - (nchunks - 1) * kChunkK) * TB * 512);

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
            tma_bulk_g2s(dst + (size_t)c * (chunk_bytes / 16), src + (size_t)c * (chunk_bytes / 16),
                         c + 1 < nchunks ? chunk_bytes : last_bytes, &full[b]);
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
        const double2* Afrag = tp.Aprime + (size_t)(si - tp.s0) * tp.a_stride + lane;
        const double2* sB = sBuf + (size_t)b * buf_elems + (size_t)(cg * BT) * 32 + lane;
        double* sSST = sSS + (size_t)b * RG * T;
        double* sMeanT = sMean + (size_t)b * RG * T;
        double* sMeanXT = sMeanX + (size_t)b * n_extra * RG * T;

        const int64_t left = tp.fast_rows - (int64_t)j * T;
        mbar_wait(&full[b], ((unsigned)(it >> 1)) & 1u);
        // the last tile of a slow block is usually short: contract only the column tiles that hold rows
        PlainB bsrc{sB, TB};
        if (BT > 2 && p.CG == 1 && left <= 16) contract_tile<2, NS>(p, Afrag, bsrc, sSST, sMeanT, sMeanXT, g, cg, lane);
        else if (BT > 4 && p.CG == 1 && left <= 32) contract_tile<4, NS>(p, Afrag, bsrc, sSST, sMeanT, sMeanXT, g, cg, lane);
        else contract_tile<BT, NS>(p, Afrag, bsrc, sSST, sMeanT, sMeanXT, g, cg, lane);
        __syncthreads();

        const int64_t tile_row0 = si * tp.fast_rows + (int64_t)j * T - p.row0;
        const int valid_cols = left < T ? (int)left : T;
        for (int t = threadIdx.x; t < T; t += kCtaThreads) {
            const int64_t row = tile_row0 + t;
            if (t < valid_cols && row >= 0 && row < p.M) finalize_row(p, sSST, sMeanT, sMeanXT, t, row);
        }
    }
}

// ---------------------------------------------------------------- tables of the grid path
// Product tables: row r of the fast table is the product over the fast (low-order) axes of exp(-0.5 ((x_j - X_nj)/l_j)^2)
// for grid row r (< fast_rows); row s of the slow table the same over the slow axes for grid row s * fast_rows, times the
// signal variance.  Padding columns n >= N are zero.
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
                               double variance, const double* __restrict__ inv_ls_d, int64_t slow0) {
    const int64_t r = blockIdx.x;             // one table row per block: the fast rows, then the slow rows slow0, slow0 + 1, ...
    const bool fast = r < ts.fast_rows;
    const int64_t tr = fast ? r : r - ts.fast_rows + slow0;
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

// PfFrag: the fast table re-ordered so that the bytes a tile needs are contiguous and already in B-fragment order:
// [tile j of a slow block][k-block][col tile][lane] double2, lane l -> rows j*T + 8ct + l/4, training points
// 8kb + 2(l%4) + {0,1}; zero beyond fast_rows / N.
__global__ void k_pffrag(const double* __restrict__ Pfast, double2* __restrict__ PfFrag, int64_t fast_rows, int N, int Npad,
                         int T, int TB, int kb_pad, int tpb) {
    const size_t total = (size_t)tpb * kb_pad * TB * 32;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(e & 31);
        size_t r = e >> 5;
        const int ct = (int)(r % TB); r /= TB;
        const int kb = (int)(r % kb_pad);
        const int j = (int)(r / kb_pad);
        const int64_t row = (int64_t)j * T + ct * 8 + (lane >> 2);
        const int n0 = 8 * kb + 2 * (lane & 3);
        double2 v = make_double2(0.0, 0.0);
        if (row < fast_rows) {
            if (n0 < N) v.x = Pfast[(size_t)row * Npad + n0];
            if (n0 + 1 < N) v.y = Pfast[(size_t)row * Npad + n0 + 1];
        }
        PfFrag[e] = v;
    }
}

// A'(s) = L^-1 diag(Pslow[s]) in the packed fragment order of Afrag.
__global__ void k_aprime(const double2* __restrict__ Afrag, const double* __restrict__ Pslow, double2* __restrict__ Aprime,
                         int NB, size_t a_stride, int64_t s0) {
    const int64_t si = s0 + blockIdx.y;
    const int Npad = 8 * NB;
    const double* ps = Pslow + (size_t)si * Npad;
    const size_t nfrag = tri_blocks(NB) * 32;
    double2* dst = Aprime + (size_t)blockIdx.y * a_stride;
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
}

}  // namespace
