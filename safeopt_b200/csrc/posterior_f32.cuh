// K2, grid path, fp32 arithmetic mode (BASELINE config 3: "fp32", tolerance 1e-4 relative): the contraction V = A'(s) Pfast on
// the 5th-generation tensor cores -- tcgen05.mma kind::tf32 issued by one thread, operands staged in shared memory by TMA bulk
// copies, fp32 accumulators in TMEM, read back with tcgen05.ld for the epilogue.
//
// Precision: one TF32 product (10-bit mantissas) is far outside the tolerance (|d var| ~ 1e-2), so every operand is split
// into two TF32 terms x = hi + lo (both exactly representable, |x - hi - lo| <= 2^-22 |x|) and each algorithmic product is
// three tensor-core products lo*hi + hi*lo + hi*hi accumulated in fp32 ("3xTF32"): the result is as good as an fp32 SGEMM
// (|d var| ~ 1e-5 sigma_f^2 at N = 128).  Only the variance goes this way -- it is the N^2/2 part of the work.  The mean
// k . alpha costs N multiply-adds per row and GP (2 % of the work) and suffers most from fp32 accumulation (alpha ~ y / noise:
// terms of size 100 cancel to O(1); the tensor core's fp32 adder left |d mean| ~ 2.6e-4 at N = 128, outside the tolerance),
// so it is computed in fp64 by k_mean_grid below, which also keeps the safe-set masks within the variance error of the fp64
// path.  The fit (Cholesky, L^-1, alpha) stays fp64, the epilogue (|V|^2, bounds, safe bit) runs in fp64 on the fp32
// accumulators, and everything downstream (set logic, expander test) is the fp64 code.
//
// Shape.  D[128 candidates x Np] (TMEM: lane = candidate, column = training row j) = A[128 x K] . B[Np x K]^T, K = training
// index n.  A = the tile's rows of the fast product table, B = A'(s) = L^-1 diag(Pslow[s]) -- the same factorisation of the
// separable RBF kernel as the fp64 grid kernel (posterior_tma.cuh).  A'(s) is lower triangular, so the K slab
// n in [32k, 32k+32) only touches columns j >= 32k: its MMAs are issued with N = Np - 32k at TMEM column 32k (62 % of the
// square at N = 128).  Both operands are pre-packed (k_f32_pack_a / k_f32_pack_b) in the no-swizzle K-major canonical layout
// of the UMMA shared-memory descriptor, [K chunk of 16 bytes][row][4 floats]: a slab is one contiguous bulk copy per operand.
//
// Work split: CTA b owns ONE fast tile j = b % tpb (128 candidate rows of every slow block) and walks the slow indices
// s_lo + b / tpb, + lanes, ...: its A operand never changes, so when it fits (N <= 128: 128 KB for hi + lo) it is loaded once and
// stays in shared memory, and only the B slabs of A'(s) stream through the ring (80 KB per 128 rows at N = 128).
// Roles (192 threads): warp 0 lane 0 = TMA producer (ring of kStages slabs, full/empty mbarriers), warp 1 lane 0 = MMA issuer
// (tcgen05.commit releases a slab / publishes an accumulator), warps 2..5 = epilogue, one thread per candidate row = TMEM
// lane.  Two accumulators of up to 256 columns: the MMAs of tile t+1 run under the epilogue of tile t.
#pragma once
#include "posterior_tma.cuh"

namespace {

constexpr int kF32Threads = 192;
constexpr int kF32TileRows = 128;
constexpr int kF32SlabK = 32;                   // training points per K slab = 8 chunks of 16 bytes
constexpr int kF32MaxNp = 256;                  // TMEM: two accumulators of 256 fp32 columns
constexpr int kF32ASlabBytes = 2 * 8 * kF32TileRows * 16;     // hi + lo planes: 32 KB

__host__ __device__ inline int f32_nslab(int N) { return (N + kF32SlabK - 1) / kF32SlabK; }
__host__ __device__ inline size_t f32_b_slab_bytes(int Np, int k) { return (size_t)2 * 8 * (Np - kF32SlabK * k) * 16; }
__host__ __device__ inline size_t f32_b_slab_offset(int Np, int k) {        // bytes before slab k inside one slow index
    // sum_{k' < k} 256 (Np - 32 k')
    return (size_t)256 * ((size_t)k * Np - (size_t)kF32SlabK * k * (k - 1) / 2);
}
__host__ __device__ inline size_t f32_b_bytes(int Np) { return f32_b_slab_offset(Np, Np / kF32SlabK); }

struct F32Params {
    PostParams p;                   // N, outputs, beta / fmin, z vectors (fp64), row range; n_out GPs share the contraction
    const unsigned char* Aop;       // [tile in slow block][slab][plane][chunk][row][4 floats]
    const unsigned char* Bop;       // [slow index - s0][slab (trimmed rows)][plane][chunk][row][4 floats]
    size_t b_stride;                // bytes per slow index
    int64_t s0, fast_rows;
    int64_t s_lo, s_hi;             // slow indices this launch covers
    int tpb, Np, nslab, stages;
    int lanes;                      // CTA b owns the fast tile j = b % tpb for the slow indices s_lo + b / tpb, + lanes, + 2 lanes, ...
    int a_resident;                 // the CTA's A operand (its fast tile, all slabs, hi + lo) stays in shared memory for the whole launch
    int mc;                         // > 1: the tpb CTAs that share a slow index form a cluster of this size and MULTICAST the B slabs:
                                    // each fetches 1 / mc of a slab from L2 and TMA delivers it to all of them
    const double* mean_in[kMaxOut]; // per output: the fp64 means of these rows (written by k_mean_grid just before)
};

// ---- tcgen05 / TMEM primitives (PTX; SASS: UTCHMMA / UTCQMMA family, LDTM, UTCBAR) ---------------------------------------
__device__ __forceinline__ void tmem_alloc_512(unsigned* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(unsigned taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(512u) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, TF32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                            unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes stored contiguously (128 B);
// SBO = byte distance between 8-row groups, LBO = byte distance between the two 16-byte K chunks of one MMA (K = 8 TF32).
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | ((unsigned long long)(lbo_bytes >> 4) << 16) |
           ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor: D fp32, A / B TF32, both K-major, M = 128, N = n.
__device__ __forceinline__ unsigned umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(kF32TileRows >> 4) << 24);
}

// Multicast variants (thread-block cluster): the bulk copy lands at the same shared-memory offset of every CTA in `mask` and
// completes bytes on the mbarrier at the same offset in each of them; the commit arrives on that barrier in every CTA of `mask`.
__device__ __forceinline__ void tma_bulk_g2s_mc(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar,
                                                unsigned short mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_mc(unsigned long long* bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void f32_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_plain(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}

struct F32Smem { size_t a_bytes, stage_bytes, bar_off, total; };
__host__ __device__ inline F32Smem f32_smem(int Np, int stages, bool a_resident) {
    F32Smem L;
    L.a_bytes = a_resident ? (size_t)(Np / kF32SlabK) * kF32ASlabBytes : 0;
    L.stage_bytes = (a_resident ? 0 : (size_t)kF32ASlabBytes) + f32_b_slab_bytes(Np, 0);
    L.bar_off = L.a_bytes + (size_t)stages * L.stage_bytes;
    L.total = L.bar_off + 16 * sizeof(unsigned long long) + 16;
    return L;
}

__global__ void __launch_bounds__(kF32Threads, 1) k_posterior_f32(const __grid_constant__ F32Params fp) {
    const PostParams& p = fp.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Np = fp.Np, nslab = fp.nslab, stages = fp.stages, n_out = p.n_out;
    const bool resident = fp.a_resident != 0;
    const F32Smem L = f32_smem(Np, stages, resident);
    unsigned char* sStage = smem_raw + L.a_bytes;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + L.bar_off);      // [stages] (<= 4)
    unsigned long long* empty = full + 4;
    unsigned long long* acc_full = full + 8;                                                      // [2]
    unsigned long long* acc_empty = full + 10;                                                    // [2]
    unsigned long long* a_full = full + 12;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(full + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x % fp.tpb;                   // this CTA's fast tile
    const int64_t s_first = fp.s_lo + blockIdx.x / fp.tpb;

    const int mc = fp.mc;
    const unsigned short mc_mask = (unsigned short)((1u << mc) - 1u);
    if (threadIdx.x == 0) {
        // with multicast a stage is refilled by every CTA of the cluster, so it is free only when all of them have released it
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], mc > 1 ? mc : 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) tmem_alloc_512(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (mc > 1) f32_cluster_sync();                      // every CTA's barriers exist before a peer's copy or commit can reach them
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;
    const unsigned char* a_src = fp.Aop + (size_t)j * nslab * kF32ASlabBytes;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (resident && s_first < fp.s_hi) {
                mbar_expect_tx(a_full, (unsigned)L.a_bytes);
                for (int k = 0; k < nslab; ++k)
                    tma_bulk_g2s(smem_raw + (size_t)k * kF32ASlabBytes, a_src + (size_t)k * kF32ASlabBytes, (unsigned)kF32ASlabBytes, a_full);
            }
            int st = 0;
            unsigned ph = 0;
            for (int64_t si = s_first; si < fp.s_hi; si += fp.lanes) {
                const unsigned char* b_src = fp.Bop + (size_t)(si - fp.s0) * fp.b_stride;
                for (int k = 0; k < nslab; ++k) {
                    mbar_wait(&empty[st], ph ^ 1u);
                    unsigned char* dst = sStage + (size_t)st * L.stage_bytes;
                    const unsigned bbytes = (unsigned)f32_b_slab_bytes(Np, k);
                    mbar_expect_tx(&full[st], (resident ? 0u : (unsigned)kF32ASlabBytes) + bbytes);
                    if (!resident) {
                        tma_bulk_g2s(dst, a_src + (size_t)k * kF32ASlabBytes, (unsigned)kF32ASlabBytes, &full[st]);
                        dst += kF32ASlabBytes;
                    }
                    if (mc > 1) {
                        const unsigned part = bbytes / (unsigned)mc;          // 256 Nk bytes, Nk a multiple of 32: divisible by 2, 4, 8
                        tma_bulk_g2s_mc(dst + (size_t)j * part, b_src + f32_b_slab_offset(Np, k) + (size_t)j * part, part, &full[st], mc_mask);
                    } else {
                        tma_bulk_g2s(dst, b_src + f32_b_slab_offset(Np, k), bbytes, &full[st]);
                    }
                    if (++st == stages) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int st = 0;
            unsigned ph = 0;
            int it = 0;
            if (resident && s_first < fp.s_hi) { mbar_wait(a_full, 0); tc_fence_after(); }
            for (int64_t si = s_first; si < fp.s_hi; si += fp.lanes, ++it) {
                const int acc = it & 1;
                mbar_wait(&acc_empty[acc], (((unsigned)it >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
                tc_fence_after();
                for (int k = 0; k < nslab; ++k) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const int Nk = Np - kF32SlabK * k;
                    const unsigned stage_base = smem_u32(sStage + (size_t)st * L.stage_bytes);
                    const unsigned a_base = resident ? smem_u32(smem_raw + (size_t)k * kF32ASlabBytes) : stage_base;
                    const unsigned b_base = resident ? stage_base : stage_base + kF32ASlabBytes;
                    const unsigned a_lbo = kF32TileRows * 16, b_lbo = (unsigned)Nk * 16;
                    const unsigned a_plane = 8 * a_lbo, b_plane = 8 * b_lbo;
                    const unsigned idesc = umma_idesc_tf32(Nk);
                    const unsigned d = tmem_base + (unsigned)(acc * kF32MaxNp + kF32SlabK * k);
#pragma unroll
                    for (int ks = 0; ks < kF32SlabK / 8; ++ks) {
                        const unsigned long long a_hi = umma_desc(a_base + 2 * ks * a_lbo, a_lbo, 128);
                        const unsigned long long a_lo = umma_desc(a_base + a_plane + 2 * ks * a_lbo, a_lbo, 128);
                        const unsigned long long b_hi = umma_desc(b_base + 2 * ks * b_lbo, b_lbo, 128);
                        const unsigned long long b_lo = umma_desc(b_base + b_plane + 2 * ks * b_lbo, b_lbo, 128);
                        // small terms first; the very first MMA of a tile overwrites all Np columns
                        tc_mma_tf32(d, a_lo, b_hi, idesc, (k | ks) != 0 ? 1u : 0u);
                        tc_mma_tf32(d, a_hi, b_lo, idesc, 1u);
                        tc_mma_tf32(d, a_hi, b_hi, idesc, 1u);
                    }
                    if (mc > 1) tc_commit_mc(&empty[st], mc_mask);       // tell every CTA of the cluster: this one has read the slab
                    else tc_commit(&empty[st]);              // the slab may be refilled once these MMAs have read it
                    if (++st == stages) { st = 0; ph ^= 1u; }
                }
                tc_commit(&acc_full[acc]);
            }
        }
    } else {
        // ===== epilogue: one thread per candidate row (TMEM lane); a warp reaches the lane quarter warp % 4 =====
        const int q = warp & 3;
        const int trow = 32 * q + lane;
        const int64_t frow = (int64_t)j * kF32TileRows + trow;                     // row inside the slow block
        int it = 0;
        for (int64_t si = s_first; si < fp.s_hi; si += fp.lanes, ++it) {
            const int acc = it & 1;
            mbar_wait(&acc_full[acc], ((unsigned)it >> 1) & 1u);
            tc_fence_after();
            // |V|^2 in fp64 from the fp32 accumulators, four independent chains
            double ss0 = 0.0, ss1 = 0.0, ss2 = 0.0, ss3 = 0.0;
            const unsigned taddr = tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)(acc * kF32MaxNp);
            for (int c0 = 0; c0 < Np; c0 += 32) {
                unsigned r[32];
                tmem_ld32(taddr + (unsigned)c0, r);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const double v0 = (double)__uint_as_float(r[i]), v1 = (double)__uint_as_float(r[i + 1]);
                    const double v2 = (double)__uint_as_float(r[i + 2]), v3 = (double)__uint_as_float(r[i + 3]);
                    ss0 = fma(v0, v0, ss0); ss1 = fma(v1, v1, ss1); ss2 = fma(v2, v2, ss2); ss3 = fma(v3, v3, ss3);
                }
            }
            const double sumsq = (ss0 + ss1) + (ss2 + ss3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_plain(&acc_empty[acc]);

            const int64_t row = si * fp.fast_rows + frow - p.row0;                  // local output row
            if (frow < fp.fast_rows && row >= 0 && row < p.M) {
                double v = p.variance - sumsq;
                v = v > SO_VAR_FLOOR ? v : SO_VAR_FLOOR;
                const double bs = __dmul_rn(p.beta, sqrt(v));
                uint8_t safe = 1;
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o) {
                    if (o >= n_out) break;
                    const double mu = fp.mean_in[o][row];     // fp64 mean of k_mean_grid (same stream, launched before)
                    const double lo = __dsub_rn(mu, bs), up = __dadd_rn(mu, bs);
                    double* vp = o == 0 ? p.var : p.var_x[o - 1];
                    const int qc = o == 0 ? p.q_col : p.q_col_x[o - 1];
                    const double fm = o == 0 ? p.fmin : p.fmin_x[o - 1];
                    if (vp) vp[row] = v;
                    if (p.Q) {
                        double* qp = p.Q + (size_t)row * p.q_stride + qc;
                        if ((p.q_stride & 1) == 0 && (qc & 1) == 0) *reinterpret_cast<double2*>(qp) = make_double2(lo, up);
                        else { qp[0] = lo; qp[1] = up; }
                    }
                    safe &= lo > fm ? 1 : 0;
                }
                if (p.safe_mode != SO_SAFE_NONE && p.S)
                    p.S[row] = p.safe_mode == SO_SAFE_WRITE ? safe : (uint8_t)(p.S[row] & safe);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (mc > 1) f32_cluster_sync();                      // no CTA leaves while a peer's multicast or commit may still target it
    if (warp == 1) tmem_dealloc_512(tmem_base);
}

// ---------------------------------------------------------------- operand packing
// Two TF32 terms of x: hi = x rounded to TF32 (10 explicit mantissa bits), lo = (x - hi) rounded to TF32.
__device__ __forceinline__ float tf32_round(float f) {
    unsigned u = __float_as_uint(f);
    u = (u + 0x1000u) & 0xFFFFE000u;
    return __uint_as_float(u);
}
__device__ __forceinline__ void tf32_split(double x, float& hi, float& lo) {
    hi = tf32_round((float)x);
    lo = tf32_round((float)(x - (double)hi));
}

// A operand: for tile t of a slow block, slab k: [plane][chunk c][row r][4 floats] = Pfast[128 t + r][32 k + 4 c + i]
__global__ void k_f32_pack_a(const double* __restrict__ Pfast, float* __restrict__ Aop, int64_t fast_rows, int N, int ldp, int nslab,
                             int tpb) {
    const size_t per_slab = 8 * kF32TileRows * 4;                 // floats per plane
    const size_t total = (size_t)tpb * nslab * per_slab;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e & 3);
        size_t r = e >> 2;
        const int row = (int)(r % kF32TileRows); r /= kF32TileRows;
        const int c = (int)(r & 7); r >>= 3;
        const int k = (int)(r % nslab);
        const int t = (int)(r / nslab);
        const int64_t fr = (int64_t)t * kF32TileRows + row;
        const int n = kF32SlabK * k + 4 * c + i;
        const double x = (fr < fast_rows && n < N) ? Pfast[(size_t)fr * ldp + n] : 0.0;
        float hi, lo;
        tf32_split(x, hi, lo);
        float* slab = Aop + ((size_t)t * nslab + k) * 2 * per_slab;
        const size_t off = ((size_t)c * kF32TileRows + row) * 4 + i;
        slab[off] = hi;
        slab[per_slab + off] = lo;
    }
}

// B operand: for slow index s, slab k (rows j >= 32 k only): [plane][chunk c][jj][4 floats] = L^-1[j][n] Pslow[s][n],
// j = 32 k + jj, n = 32 k + 4 c + i (zero above the diagonal and in the padding).
__global__ void k_f32_pack_b(const double* __restrict__ Linv, int ld, const double* __restrict__ Pslow, int ldp,
                             float* __restrict__ Bop, size_t b_stride_floats, int64_t s0, int N, int Np) {
    const int64_t si = s0 + blockIdx.y;
    const double* ps = Pslow + (size_t)si * ldp;
    float* dst = Bop + (size_t)blockIdx.y * b_stride_floats;
    const int nslab = Np / kF32SlabK;
    for (int k = 0; k < nslab; ++k) {
        const int Nk = Np - kF32SlabK * k;
        float* slab = dst + f32_b_slab_offset(Np, k) / 4;
        const size_t plane = (size_t)8 * Nk * 4;
        for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
            const int i = (int)(e & 3);
            size_t r = e >> 2;
            const int jj = (int)(r % Nk);
            const int c = (int)(r / Nk);
            const int j = kF32SlabK * k + jj, n = kF32SlabK * k + 4 * c + i;
            const double x = (j < N && n <= j) ? Linv[(size_t)j * ld + n] * ps[n] : 0.0;
            float hi, lo;
            tf32_split(x, hi, lo);
            slab[e] = hi;
            slab[plane + e] = lo;
        }
    }
}

// ---------------------------------------------------------------- fp64 mean of the fp32 mode
// mean_g(row) = sum_n Pfast[f][n] (Pslow[s][n] alpha_g[n]), f = row % F, s = row / F: N multiply-adds per row and GP.  For a fast
// tile (TR rows) and a batch of SB slow indices this is a small GEMM, C[TR x (SB G)] = Pfast_tile[TR x N] . W[N x (SB G)] with
// W[n][(s, g)] = Pslow[s][n] alpha_g[n], and runs on the fp64 tensor pipe (DMMA m8n8k4).  CTA (jt, y) owns the rows
// [TR jt, TR jt + TR) of every slow block: its slice of the fast table (TR x N x 8 bytes <= 128 KB) is loaded ONCE into shared
// memory, in A-fragment order ([n / 4][row][n % 4]: a warp's fragment load is 32 consecutive doubles), and stays for the whole
// loop over the slow-index batches; W is rebuilt per batch in B-fragment order.  Each warp computes 2 row blocks x 2 column
// blocks per k-step: four 256-byte shared-memory reads feed four DMMAs.
constexpr int kMeanThreads = 256;
constexpr int kMeanCols = 16;           // (slow index, GP) columns per batch: two 8-wide column blocks
struct MeanParams {
    const double* PfastT;       // [n][Fpad]
    const double* Pslow;        // [s][ldp]
    const double* alpha[kMaxOut];
    double* mean[kMaxOut];
    int n_out, N, N4, ldp, TR, SB;
    int64_t Fpad, fast_rows, s_lo, s_hi, row0, M;
};
__host__ __device__ inline size_t mean_smem_bytes(int N4, int TR) {
    return ((size_t)N4 * TR + (size_t)N4 * kMeanCols) * sizeof(double);
}

__global__ void __launch_bounds__(kMeanThreads, 1) k_mean_grid(const __grid_constant__ MeanParams mp) {
    extern __shared__ __align__(16) unsigned char mean_smem[];
    const int TR = mp.TR, SB = mp.SB, N = mp.N, N4 = mp.N4, n_out = mp.n_out;
    double* sT = reinterpret_cast<double*>(mean_smem);              // [n / 4][TR][n % 4]
    double* sW = sT + (size_t)N4 * TR;                              // [n / 4][kMeanCols][n % 4]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f0 = (int64_t)blockIdx.x * TR;
    for (int e = threadIdx.x; e < N4 * TR; e += kMeanThreads) {
        const int n = e / TR, r = e - n * TR;                       // consecutive threads read consecutive rows of the transposed table
        const double v = (n < N && f0 + r < mp.Fpad) ? mp.PfastT[(size_t)n * mp.Fpad + f0 + r] : 0.0;
        sT[((size_t)(n >> 2) * TR + r) * 4 + (n & 3)] = v;
    }
    const int nrb = TR / 8;                                         // row blocks of the tile: 16 or 8
    for (int64_t sb = mp.s_lo + (int64_t)blockIdx.y * SB; sb < mp.s_hi; sb += (int64_t)gridDim.y * SB) {
        __syncthreads();                                            // the tile is loaded / the previous batch's readers of sW are done
        for (int e = threadIdx.x; e < N4 * kMeanCols; e += kMeanThreads) {
            const int n = e / kMeanCols, col = e - n * kMeanCols;
            const int si = col / n_out, o = col - si * n_out;
            double v = 0.0;
            if (n < N && si < SB && sb + si < mp.s_hi) v = mp.Pslow[(size_t)(sb + si) * mp.ldp + n] * mp.alpha[o][n];
            sW[((size_t)(n >> 2) * kMeanCols + col) * 4 + (n & 3)] = v;
        }
        __syncthreads();
        for (int rb0 = 2 * warp; rb0 < nrb; rb0 += 2 * (kMeanThreads / 32)) {
            double c[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
            const double* ap = sT + (size_t)rb0 * 32 + lane;            // fragment of row block rb at k-step ks: sT[(ks TR + 8 rb) 4 + lane]
            const double* bp = sW + lane;
#pragma unroll 4
            for (int ks = 0; ks < N4 / 4; ++ks) {
                const double a0 = ap[(size_t)ks * TR * 4], a1 = ap[(size_t)ks * TR * 4 + 32];
                const double b0 = bp[(size_t)ks * kMeanCols * 4], b1 = bp[(size_t)ks * kMeanCols * 4 + 32];
                dmma884(c[0][0][0], c[0][0][1], a0, b0);
                dmma884(c[0][1][0], c[0][1][1], a0, b1);
                dmma884(c[1][0][0], c[1][0][1], a1, b0);
                dmma884(c[1][1][0], c[1][1][1], a1, b1);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int col = 8 * j + 2 * (lane & 3) + hh;
                        const int si = col / n_out, o = col - si * n_out;
                        const int64_t frow = f0 + 8 * (rb0 + i) + (lane >> 2);
                        const int64_t row = (sb + si) * mp.fast_rows + frow - mp.row0;
                        if (si < SB && sb + si < mp.s_hi && frow < mp.fast_rows && row >= 0 && row < mp.M) mp.mean[o][row] = c[i][j][hh];
                    }
        }
    }
}

// PfastT[n][f] = Pfast[f][n] (zero for f >= fast_rows)
__global__ void k_transpose_pfast(const double* __restrict__ Pfast, double* __restrict__ PfastT, int64_t fast_rows, int64_t Fpad,
                                  int N, int ldp) {
    __shared__ double tile[32][33];
    const int64_t f0 = (int64_t)blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int64_t f = f0 + i;
        const int n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (f < fast_rows && n < N) ? Pfast[(size_t)f * ldp + n] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int n = n0 + i;
        const int64_t f = f0 + threadIdx.x;
        if (n < N && f < Fpad) PfastT[(size_t)n * Fpad + f] = tile[threadIdx.x][i];
    }
}

}  // namespace
