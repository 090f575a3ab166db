// K2, grid path, "ring" kernel: k_posterior_tma for fits whose B tile no longer fits twice in shared memory (N > 280).
//
// k_posterior_tma keeps the whole B tile of a row block (T rows x all NB k-blocks) resident and double-buffers it, so the
// tile shrinks as N grows (48 rows up to N = 280, 32 up to 416, 16 beyond) and with it the reuse of every A fragment.  Here
// B streams through a shared-memory ring of k-block chunks (kRingKC k-blocks x TB column tiles, 12 KB at TB = 6) filled by a
// producer lane with cp.async.bulk; the tile stays 48 rows wide for any N, shared memory no longer depends on N.
//   * consumers: the eight warps of contract_tile (block rows over the warps, row table or closed-form pairing); a warp
//     waits for a chunk's `full` mbarrier when it enters it and arrives on its `empty` mbarrier when it leaves it; warps whose
//     rows end early still walk the remaining chunks (drain) so that every chunk sees eight arrivals;
//   * a pass re-walks k from 0, so the producer streams the tile's B slice once per pass (L2 traffic, the table stays
//     L2-resident);
//   * the producer is lane 0 of warp 0 and never blocks (RingProducer); one __syncthreads per tile, then the T-thread
//     epilogue, as in k_posterior_tma.
#pragma once
#include "posterior_tma.cuh"

namespace {

constexpr int kRingKC = 8;             // k-blocks per ring stage

struct RingSmem { size_t stage_bytes, ss_off, mean_off, meanx_off, bar_off, total; int stages; };

__host__ __device__ inline RingSmem ring_layout(int TB, int RG, int T, int n_extra, int stages) {
    RingSmem L;
    L.stage_bytes = (size_t)kRingKC * TB * 512;
    L.stages = stages;
    L.ss_off = (size_t)stages * L.stage_bytes;
    L.mean_off = L.ss_off + 2 * (size_t)RG * T * sizeof(double);
    L.meanx_off = L.mean_off + 2 * (size_t)RG * T * sizeof(double);
    L.bar_off = L.meanx_off + 2 * (size_t)n_extra * RG * T * sizeof(double);
    L.total = L.bar_off + 2 * 32 * sizeof(unsigned long long);
    return L;
}

// As many stages as fit (at most 16).
__host__ inline RingSmem ring_smem(int TB, int RG, int T, int n_extra, size_t limit) {
    int st = 16;
    while (st > 0 && ring_layout(TB, RG, T, n_extra, st).total > limit) --st;
    return ring_layout(TB, RG, T, n_extra, st);
}

__device__ __forceinline__ void mbar_arrive1(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// The producer is lane 0 of warp 0, never blocking: whenever warp 0 hands a chunk back, and while it waits for one, it
// refills every ring stage that all eight warps have released (a ninth warp would cost registers: a 288-thread CTA is
// allocated like a 384-thread one, 168 registers per thread).  The ring is deep (two tiles' worth at N = 256), so a refill
// that waits for warp 0's next visit is never late.
struct RingProducer {
    const TmaParams* tp;
    unsigned char* smem_raw;
    unsigned long long* full;
    unsigned long long* empty;
    size_t stage_bytes;
    int stages, nchunks;
    int st;
    unsigned ph;
    int64_t tile;           // next chunk to issue: (tile, pass, chunk)
    int pass, chunk;
    __device__ __forceinline__ void pump() {
        const PostParams& p = tp->p;
        while (tile < p.ntiles && mbar_test(&empty[st], ph ^ 1u)) {
            const int NB = p.NB, TB = p.TB;
            const int64_t gt = tp->first_tile + tile;
            const int j = (int)(gt % tp->tpb);
            const double2* src = tp->PfFrag + (size_t)j * NB * TB * 32 + (size_t)chunk * kRingKC * TB * 32;
            const int kbs = NB - chunk * kRingKC < kRingKC ? NB - chunk * kRingKC : kRingKC;
            const unsigned bytes = (unsigned)kbs * TB * 512u;
            mbar_expect_tx(&full[st], bytes);
            tma_bulk_g2s(smem_raw + (size_t)st * stage_bytes, src, bytes, &full[st]);
            if (++st == stages) { st = 0; ph ^= 1u; }
            if (++chunk == nchunks) {
                chunk = 0;
                if (++pass == p.npass) { pass = 0; tile += gridDim.x; }
            }
        }
    }
};

// B fragments of k-block kb from the ring (see PlainB for the interface contract_tile expects).
struct RingB {
    const double2* ring;        // ring base + this warp's column group + lane
    unsigned long long* full;
    unsigned long long* empty;
    size_t stage_elems;
    int TB, NB, stages, lane;
    int st;                     // stage of the chunk being walked
    unsigned ph;
    RingProducer* prod;         // non-null in warp 0
    static constexpr bool kStreaming = true;
    static constexpr int kSpan = kRingKC;
    __device__ __forceinline__ void enter(int) {
        if (prod) {
            while (true) {
                const unsigned ok = __shfl_sync(0xffffffffu, mbar_test(&full[st], ph) ? 1u : 0u, 0);
                if (ok) break;
                if (lane == 0) prod->pump();
                __syncwarp();
            }
        } else {
            mbar_wait(&full[st], ph);
        }
    }
    __device__ __forceinline__ const double2* at(int kb) const {
        return ring + (size_t)st * stage_elems + (size_t)(kb % kRingKC) * TB * 32;
    }
    __device__ __forceinline__ void leave(int kb) {
        if ((kb % kRingKC) == kRingKC - 1 || kb == NB - 1) {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive1(&empty[st]);
                if (prod) prod->pump();
            }
            __syncwarp();
            if (++st == stages) { st = 0; ph ^= 1u; }
        }
    }
    __device__ __forceinline__ void drain(int kb_from) {
        for (int kb = kb_from; kb < NB; ++kb) {
            if (kb % kRingKC == 0) enter(kb);
            leave(kb);
        }
    }
};

template <int BT>
__global__ void __launch_bounds__(kThreads, 1) k_posterior_ring(const __grid_constant__ TmaParams tp, int stages) {
    const PostParams& p = tp.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n_extra = p.n_out - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RG = p.RG, T = p.T, TB = p.TB, NB = p.NB;
    const RingSmem L = ring_layout(TB, RG, T, n_extra, stages);
    double* sSS = reinterpret_cast<double*>(smem_raw + L.ss_off);
    double* sMean = reinterpret_cast<double*>(smem_raw + L.mean_off);
    double* sMeanX = reinterpret_cast<double*>(smem_raw + L.meanx_off);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + L.bar_off);
    unsigned long long* empty = full + 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    RingProducer prod;
    prod.tp = &tp; prod.smem_raw = smem_raw; prod.full = full; prod.empty = empty; prod.stage_bytes = L.stage_bytes;
    prod.stages = stages; prod.nchunks = (NB + kRingKC - 1) / kRingKC; prod.st = 0; prod.ph = 0;
    prod.tile = blockIdx.x; prod.pass = 0; prod.chunk = 0;
    if (threadIdx.x == 0) prod.pump();          // first lap: every stage is free

    const int g = warp % RG, cg = warp / RG;
    RingB bs;
    bs.ring = reinterpret_cast<const double2*>(smem_raw) + (size_t)(cg * BT) * 32 + lane;
    bs.full = full; bs.empty = empty; bs.stage_elems = L.stage_bytes / sizeof(double2);
    bs.TB = TB; bs.NB = NB; bs.stages = stages; bs.lane = lane; bs.st = 0; bs.ph = 0;
    bs.prod = warp == 0 ? &prod : nullptr;

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const int64_t gt = tp.first_tile + tile;
        const int64_t si = gt / tp.tpb;
        const int j = (int)(gt - si * tp.tpb);
        const double2* Afrag = tp.Aprime + (size_t)(si - tp.s0) * tp.a_stride + lane;
        double* sSST = sSS + (size_t)b * RG * T;
        double* sMeanT = sMean + (size_t)b * RG * T;
        double* sMeanXT = sMeanX + (size_t)b * n_extra * RG * T;
        const int64_t left = tp.fast_rows - (int64_t)j * T;

        if (BT > 2 && p.CG == 1 && left <= 16) contract_tile<2, 4, RingB>(p, Afrag, bs, sSST, sMeanT, sMeanXT, g, cg, lane);
        else if (BT > 4 && p.CG == 1 && left <= 32) contract_tile<4, 4, RingB>(p, Afrag, bs, sSST, sMeanT, sMeanXT, g, cg, lane);
        else contract_tile<BT, 4, RingB>(p, Afrag, bs, sSST, sMeanT, sMeanXT, g, cg, lane);
        __syncthreads();

        const int64_t tile_row0 = si * tp.fast_rows + (int64_t)j * T - p.row0;
        const int valid_cols = left < T ? (int)left : T;
        for (int t = threadIdx.x; t < T; t += kThreads) {
            const int64_t row = tile_row0 + t;
            if (t < valid_cols && row >= 0 && row < p.M) finalize_row(p, sSST, sMeanT, sMeanXT, t, row);
        }
    }
}

}  // namespace
