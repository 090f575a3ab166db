// K1 in ONE launch (N <= 512): the whole training-side fit -- scaled inputs, Ky = K(X,X) + (noise + 1e-8) I, blocked
// right-looking Cholesky, the triangular inverse W = L^-1 (forward substitution on the identity, done right-looking
// alongside the factorisation), z = W y, alpha = W^T z and the DMMA fragment packing -- by one thread-block CLUSTER whose
// CTAs synchronise with the hardware cluster barrier (~0.2 us) instead of kernel boundaries (the multi-kernel version below
// spends 2 launches per 32-wide panel; at N = 256 that was 21 launches and 0.44 ms, almost all of it launch latency and
// instruction fetch of fully unrolled single-use code).
//
// Per panel P (32 columns at p0), with R = the rows below it:
//   1. every CTA factors the 32x32 diagonal block in its own shared memory (256 threads, one barrier per column: the
//      trailing update uses a_ik -= a_ij a_kj / d_jj on the unscaled column, the scaled column goes to a separate array) and
//      inverts it (32 forward substitutions) -- redundantly, which saves a broadcast and a cluster barrier;
//   2. 32x32 tiles, round-robin over the CTAs, each a DMMA (mma.sync m8n8k4 f64) product from shared memory:
//         L[R,P]      = A[R,P] inv(L_PP)^T          (the panel solve as a product with the explicit 32x32 inverse)
//         W[P,0:p0]   = inv(L_PP) W[P,0:p0]
//      cluster barrier
//   3. trailing tiles:   A[R,R] -= L[R,P] L[R,P]^T (lower tiles)   and   W[R,0:p0+32] -= L[R,P] W[P,0:p0+32]
//      cluster barrier
// The matrices stay in global memory (<= 2 MB each: L2-resident); cross-CTA data is read with ld.global.cg.
#pragma once
#include "common.cuh"

namespace {

constexpr int kFcThreads = 256;
constexpr int kFcB = 32;                 // panel width = tile size
constexpr int kFcLd = kFcB + 1;
constexpr int kFcMaxN = 512;
constexpr size_t kFcDynSmem = (size_t)(kFcThreads / 32) * 2 * kFcB * kFcLd * sizeof(double);     // 8 warps x two 32x33 tiles = 135 KB

struct FitClusterParams {
    const double* X;         // N x d raw inputs (device)
    const double* Y;         // N
    double* Xs;              // Npad x d scaled inputs out
    double* K;               // work / L, leading dimension ld
    double* W;               // L^-1, leading dimension ld
    double* alpha;
    double* zvec;
    double2* Afrag;
    int N, Npad, NP, ld, d, kind, NB;
    double variance, diag_add;
    double inv_ls[SO_MAX_DIM];
    int* status;
#ifdef SO_FIT_STAMPS
    long long* stamps;       // timing experiments only (tools/build_variant.sh): clock64 of CTA 0 at the phase boundaries
#endif
};

#ifdef SO_FIT_STAMPS
#define FC_STAMP(i) do { if (rank == 0 && tid == 0) fp.stamps[i] = clock64(); } while (0)
#else
#define FC_STAMP(i) do {} while (0)
#endif

__device__ __forceinline__ unsigned fc_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned fc_cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(r)); return r; }
// The release / acquire pair of the hardware cluster barrier orders the CTAs' global-memory traffic at cluster scope (every
// consumer reads with ld.global.cg); a __threadfence() in front of it cost 6 % of the warp samples and is not needed.
__device__ __forceinline__ void fc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__device__ __forceinline__ double fc_kernel(int kind, double r2, double variance) {
    switch (kind) {
        case SO_KERNEL_RBF: return kernel_of_r2<SO_KERNEL_RBF>(r2, variance);
        case SO_KERNEL_MATERN32: return kernel_of_r2<SO_KERNEL_MATERN32>(r2, variance);
        default: return kernel_of_r2<SO_KERNEL_MATERN52>(r2, variance);
    }
}

// Work units are 8x32 STRIPS of a 32x32 output tile, one per warp (each warp has its own shared-memory staging), so a CTA has
// eight units in flight and a 16-CTA cluster 128: even at N = 128 (a handful of tiles per phase) most warps have work, and no
// block-level barrier sits between a unit's loads and its use.  A unit is bounded by one L2 round trip for its operands.
// rows x 32 block <- global (rows r0.., cols c0.., leading dimension ld); optional transpose.  Lane l loads column l of each row.
template <int ROWS>
__device__ __forceinline__ void fc_load(double (*s)[kFcLd], const double* __restrict__ G, int ld, int r0, int c0, bool transpose) {
    const int lane = threadIdx.x & 31;
    double v[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) v[i] = __ldcg(G + (size_t)(r0 + i) * ld + c0 + lane);
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
        if (transpose) s[lane][i] = v[i]; else s[i][lane] = v[i];
    }
}

// C(8x32) = A(8x32, [i][k]) . B(32x32, [k][j]) on the fp64 tensor pipe, by one warp: four 8x8 blocks, lane l ends with
// C[l/4][8 bj + 2 (l%4) + {0,1}] in c[bj][0..1].  `arow` = first row of the strip inside sA.
__device__ __forceinline__ void fc_strip_mma(const double (*sA)[kFcLd], int arow, const double (*sB)[kFcLd], double (&c)[4][2]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int bj = 0; bj < 4; ++bj) c[bj][0] = c[bj][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < kFcB / 4; ++ks) {
        const double a = sA[arow + (lane >> 2)][4 * ks + (lane & 3)];
#pragma unroll
        for (int bj = 0; bj < 4; ++bj) dmma884(c[bj][0], c[bj][1], a, sB[4 * ks + (lane & 3)][8 * bj + (lane >> 2)]);
    }
}

// dst strip (global rows r0 .. r0+8): mode 0 = C, 1 = -C (first touch: starts from zero), 2 = dst - C
__device__ __forceinline__ void fc_store_strip(double* __restrict__ G, int ld, int r0, int c0, const double (&c)[4][2], int mode) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int bj = 0; bj < 4; ++bj) {
        double2* p = reinterpret_cast<double2*>(G + (size_t)(r0 + (lane >> 2)) * ld + c0 + 8 * bj + 2 * (lane & 3));
        double2 v = make_double2(c[bj][0], c[bj][1]);
        if (mode == 1) { v.x = -v.x; v.y = -v.y; }
        else if (mode == 2) { const double2 o = __ldcg(p); v.x = o.x - v.x; v.y = o.y - v.y; }
        *p = v;
    }
}

__global__ void __launch_bounds__(kFcThreads, 1) k_fit_cluster(const __grid_constant__ FitClusterParams fp) {
    extern __shared__ __align__(16) unsigned char fc_dyn[];      // per warp: two 32x33 tiles
    __shared__ double sL[kFcB][kFcLd];      // L_PP
    __shared__ double sI[kFcB][kFcLd];      // inv(L_PP), lower
    __shared__ double sIT[kFcB][kFcLd];     // its transpose
    __shared__ double sU[kFcB][kFcLd];      // sU[j][t] = unscaled column j of the block being factored (d_jj on the diagonal)
    __shared__ double sRcp[kFcB];           // 1 / d_jj
    __shared__ double sD[kFcB];             // d_jj (1 where the factorisation failed)
    __shared__ int sBad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = fc_cluster_rank(), csize = fc_cluster_size();
    const int N = fp.N, NP = fp.NP, ld = fp.ld, d = fp.d;
    const int gtid = rank * kFcThreads + tid, gthreads = csize * kFcThreads;
    const int gwarp = rank * (kFcThreads / 32) + warp, gwarps = csize * (kFcThreads / 32);
    const int nblk = NP / kFcB;
    double (*wA)[kFcLd] = reinterpret_cast<double (*)[kFcLd]>(fc_dyn + (size_t)warp * 2 * kFcB * kFcLd * sizeof(double));
    double (*wB)[kFcLd] = wA + kFcB;

    // Diagonal block, two roles that meet at one barrier per column (see the panel loop): warps 0..5 factor it (a thread keeps up
    // to three elements of the 32x32 lower triangle in registers), warps 6..7 invert it one column step behind.
    constexpr int kFcFactorThreads = kFcThreads - 2 * kFcB, kFcElems = 3;
    static_assert(kFcFactorThreads * kFcElems >= kFcB * (kFcB + 1) / 2, "the factoring warps cover the lower triangle");
    const bool factoring = tid < kFcFactorThreads;
    int ei[kFcElems], ek[kFcElems];
#pragma unroll
    for (int q = 0; q < kFcElems; ++q) {
        const int e = tid + q * kFcFactorThreads;
        int ii = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while ((ii + 1) * (ii + 2) / 2 <= e) ++ii;
        while (ii * (ii + 1) / 2 > e) --ii;
        const bool mine = factoring && e < kFcB * (kFcB + 1) / 2;
        ei[q] = mine ? ii : -1;
        ek[q] = mine ? e - ii * (ii + 1) / 2 : 0;
    }

    FC_STAMP(0);
    // ---- phase 0: scaled inputs, Ky (identity in the padding), W = 0
    for (int e = gtid; e < fp.Npad * d; e += gthreads) {
        const int n = e / d, j = e - n * d;
        fp.Xs[e] = n < N ? fp.X[e] * fp.inv_ls[j] : 0.0;
    }
    double* sXs = reinterpret_cast<double*>(fc_dyn);           // N x d scaled inputs (<= 64 KB of the 135 KB the tiles use later)
    for (int e = tid; e < N * d; e += kFcThreads) sXs[e] = fp.X[e] * fp.inv_ls[e % d];
    __syncthreads();
    // Ky: only the lower triangle is ever read (diagonal blocks, panels and trailing tiles below the diagonal), so only it is
    // evaluated -- two elements per thread and trip, so that two distance -> exp chains are in flight
    {
        const int ntri = NP * (NP + 1) / 2;
        auto tri_row = [](int t) {
            int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while ((i + 1) * (i + 2) / 2 <= t) ++i;
            while (i * (i + 1) / 2 > t) --i;
            return i;
        };
        for (int t0 = gtid; t0 < ntri; t0 += 2 * gthreads) {
            const int t1 = t0 + gthreads < ntri ? t0 + gthreads : t0;       // the last odd trip repeats t0 (same store)
            const int i0 = tri_row(t0), j0 = t0 - i0 * (i0 + 1) / 2, i1 = tri_row(t1), j1 = t1 - i1 * (i1 + 1) / 2;
            const bool in0 = i0 < N, in1 = i1 < N;                          // j <= i: inside as soon as the row is
            const double* xa0 = sXs + (in0 ? i0 : 0) * d;
            const double* xb0 = sXs + (in0 ? j0 : 0) * d;
            const double* xa1 = sXs + (in1 ? i1 : 0) * d;
            const double* xb1 = sXs + (in1 ? j1 : 0) * d;
            double r0 = 0.0, r1 = 0.0;
            for (int c = 0; c < d; ++c) {
                const double u0 = xa0[c] - xb0[c], u1 = xa1[c] - xb1[c];
                r0 = fma(u0, u0, r0);
                r1 = fma(u1, u1, r1);
            }
            double v0 = fc_kernel(fp.kind, r0, fp.variance), v1 = fc_kernel(fp.kind, r1, fp.variance);
            if (i0 == j0) v0 += fp.diag_add;
            if (i1 == j1) v1 += fp.diag_add;
            if (!in0) v0 = i0 == j0 ? 1.0 : 0.0;                            // identity in the padding
            if (!in1) v1 = i1 == j1 ? 1.0 : 0.0;
            fp.K[(size_t)i0 * ld + j0] = v0;
            fp.K[(size_t)i1 * ld + j1] = v1;
        }
        for (int e = gtid; e < NP * NP; e += gthreads) {
            const int i = e / NP, j = e - i * NP;
            fp.W[(size_t)i * ld + j] = 0.0;
            if (j > i) fp.K[(size_t)i * ld + j] = 0.0;
        }
    }
    if (tid == 0) sBad = 0;
    FC_STAMP(1);
    fc_cluster_sync();

    for (int pi = 0; pi < nblk; ++pi) {
        const int p0 = pi * kFcB;
        FC_STAMP(2 + 6 * pi);
        // ---- 1. diagonal block: factor and invert (every CTA, redundantly).  The trailing elements live in registers; per
        //         column j the owners of a_ij (k == j) put the unscaled column into shared memory, one barrier, and every
        //         remaining element takes a_ik -= a_ij a_kj / d_jj.  1 / L_jj = rsqrt(d_jj), no division, no sqrt.
        double a[kFcElems];
#pragma unroll
        for (int q = 0; q < kFcElems; ++q) a[q] = ei[q] >= 0 ? __ldcg(fp.K + (size_t)(p0 + ei[q]) * ld + p0 + ek[q]) : 0.0;
        // A dependent fp64 operation takes ~40 cycles here and every column step is a chain of them (measured: ~340 cycles per column
        // of the factorisation and ~310 per row of the inverse when all eight warps did one after the other), so
        //  * the factorisation works on the UNSCALED columns u_j (L = U diag(d)^-1/2): per column only d_jj -> 1 / d_jj ->
        //    a_ik -= u_ij u_kj / d_jj is sequential.  1 / d = r0 (1 + e + e^2 + ...) with the MUFU seed r0 and e = 1 - d r0
        //    (|e| < 2^-20); with t = u_ij u_kj r0 formed while e is in flight, a - t - t (e + e^2) leaves three dependent
        //    operations behind the seed.  No square root in the loop at all;
        //  * the inverse rides along one column step behind in warps 6..7, also unscaled and RIGHT-looking: Y = inv(U), column c by
        //    two threads that keep the running right-hand sides P_i = delta_ic - sum_{k' < k} u_ik' y_k' of the rows i = sub mod 2
        //    in registers.  When column k of U and 1 / d_kk are published: y_k = P_k / d_kk (one multiply), one shuffle to the
        //    partner, P_i -= u_ik y_k for i > k (independent FMAs) -- two dependent operations per step instead of a 31-term sum;
        //  * after the last column one pass scales both: L_ik = u_ik rsqrt(d_kk), inv(L)_ic = sqrt(d_ii) Y_ic.
        const int c = (tid - kFcFactorThreads) >> 1, sub = tid & 1;
        double P[16];                                       // P[m] = running right-hand side of row 2 m + sub
#pragma unroll
        for (int m = 0; m < 16; ++m) P[m] = (2 * m + sub == c) ? 1.0 : 0.0;
        auto inverse_col = [&](const int k) {
            double yk = P[k >> 1] * sRcp[k];                // right in the thread that owns row k
            yk = __shfl_sync(0xffffffffu, yk, (lane & ~1) | (k & 1));
            if (sub == 0) sI[k][c] = yk;
#pragma unroll
            for (int m = 0; m < 16; ++m)
                if (2 * m + 1 > k) {                        // rows i = 2 m + sub > k (checked per thread below)
                    const int i = 2 * m + sub;
                    if (i > k) P[m] = fma(-sU[k][i], yk, P[m]);
                }
        };
#pragma unroll
        for (int j = 0; j < kFcB; ++j) {
            double* col = sU[j];
#pragma unroll
            for (int q = 0; q < kFcElems; ++q)
                if (ei[q] >= 0 && ek[q] == j) col[ei[q]] = a[q];
            __syncthreads();
            if (factoring) {
#ifdef SO_FC_SKIP_FACTOR      // SO_FC_SKIP_*: timing experiments only (tools/build_variant.sh), results are garbage
                if (tid > 1000)
#endif
                {
                    double prod[kFcElems];
#pragma unroll
                    for (int q = 0; q < kFcElems; ++q) prod[q] = col[ei[q] >= 0 ? ei[q] : 0] * col[ek[q]];
                    const double djj = col[j];
                    const bool ok = djj > 0.0 && !isinf(djj);
                    const double dd = ok ? djj : 1.0;       // a failed factorisation is flagged; the arithmetic stays finite
                    double r0;
                    asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(r0) : "d"(dd));
                    const double e = fma(-dd, r0, 1.0);
                    const double pe = fma(e, e, e);
#pragma unroll
                    for (int q = 0; q < kFcElems; ++q) {
                        const double t = prod[q] * r0;
                        const double u = a[q] - t;
                        if (ei[q] >= 0 && ek[q] > j) a[q] = fma(-t, pe, u);
                    }
                    if (tid == 0) {
                        sRcp[j] = fma(r0, pe, r0);
                        sD[j] = dd;
                        if (!ok) sBad = 1;
                    }
                }
            } else if (j > 0) {
#ifdef SO_FC_SKIP_INV
                if (tid > 1000)
#endif
                inverse_col(j - 1);                         // column j-1 of U and 1 / d_{j-1,j-1} were published before this step's barrier
            }
        }
        __syncthreads();
        if (!factoring) inverse_col(kFcB - 1);
        __syncthreads();
        // scaling pass: rsqrt(d_kk) once per column, then L, inv(L) and its transpose
        if (tid < kFcB) sRcp[tid] = rsqrt(sD[tid]);         // the reciprocals are not needed any more
        __syncthreads();
        {
            const int i = tid >> 3, k0 = (tid & 7) * 4;
            const double sq = sD[i] * sRcp[i];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + q;
                const double y = k <= i ? sI[i][k] * sq : 0.0;
                sI[i][k] = y;
                sIT[k][i] = y;
                sL[i][k] = k <= i ? sU[k][i] * sRcp[k] : 0.0;
            }
        }
        __syncthreads();
        FC_STAMP(4 + 6 * pi);
        if (rank == 0) {                                // L_PP into K, inv(L_PP) into W[P,P]
            for (int e = tid; e < kFcB * kFcB; e += kFcThreads) {
                const int i = e >> 5, j = e & 31;
                if (j <= i) {
                    fp.K[(size_t)(p0 + i) * ld + p0 + j] = sL[i][j];
                    fp.W[(size_t)(p0 + i) * ld + p0 + j] = sI[i][j];
                }
            }
        }
        // ---- 2. panel solve and W[P, 0:p0]: 8x32 strips, one per warp, round-robin over the cluster's warps
        {
            const int n_rows = nblk - 1 - pi, n_cols = pi;
            for (int u = gwarp; u < 4 * n_rows + n_cols; u += gwarps) {
                if (u < 4 * n_rows) {
                    // L[rb,P] = A[rb,P] inv(L_PP)^T : A strip [i][k], B[k][j] = inv[j][k]
                    double c[4][2];
                    const int r0 = (pi + 1 + (u >> 2)) * kFcB + 8 * (u & 3);
                    fc_load<8>(wA, fp.K, ld, r0, p0, false);
                    __syncwarp();
                    fc_strip_mma(wA, 0, sIT, c);
                    fc_store_strip(fp.K, ld, r0, p0, c, 0);
                } else {
                    // W[P,cb] = inv(L_PP) W[P,cb], in place: the whole tile by ONE warp (every strip needs all rows of the old tile)
                    const int c0 = (u - 4 * n_rows) * kFcB;
                    fc_load<kFcB>(wB, fp.W, ld, p0, c0, false);
                    __syncwarp();
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        double c[4][2];
                        fc_strip_mma(sI, 8 * st, wB, c);
                        fc_store_strip(fp.W, ld, p0 + 8 * st, c0, c, 0);
                    }
                }
                __syncwarp();
            }
        }
        FC_STAMP(5 + 6 * pi);
        fc_cluster_sync();
        FC_STAMP(6 + 6 * pi);
        // ---- 3. trailing update: A[rb,cb] -= L[rb,P] L[cb,P]^T (pi < cb <= rb) and W[rb,cb] -= L[rb,P] W[P,cb] (cb <= pi)
        {
            const int m = nblk - 1 - pi;                // row blocks below the panel
            const int per_row = pi + 1;                 // W tiles per row block
            const int n_a = m * (m + 1) / 2, n_w = m * per_row;
            for (int u = gwarp; u < 4 * (n_a + n_w); u += gwarps) {
                const int t = u >> 2, st = u & 3;
                double c[4][2];
                if (t < n_a) {
                    int ii = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                    while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
                    while (ii * (ii + 1) / 2 > t) --ii;
                    const int kk = t - ii * (ii + 1) / 2;
                    const int r0 = (pi + 1 + ii) * kFcB + 8 * st, c0 = (pi + 1 + kk) * kFcB;
                    fc_load<8>(wA, fp.K, ld, r0, p0, false);
                    fc_load<kFcB>(wB, fp.K, ld, c0, p0, true);            // B[k][j] = L[c0 + j][p0 + k]
                    __syncwarp();
                    fc_strip_mma(wA, 0, wB, c);
                    fc_store_strip(fp.K, ld, r0, c0, c, 2);
                } else {
                    const int q = t - n_a;
                    const int rb = q / per_row, cb = q - rb * per_row;
                    const int r0 = (pi + 1 + rb) * kFcB + 8 * st, c0 = cb * kFcB;
                    fc_load<8>(wA, fp.K, ld, r0, p0, false);
                    fc_load<kFcB>(wB, fp.W, ld, p0, c0, false);
                    __syncwarp();
                    fc_strip_mma(wA, 0, wB, c);
                    fc_store_strip(fp.W, ld, r0, c0, c, cb == pi ? 1 : 2);  // W[R,P] starts from zero
                }
                __syncwarp();
            }
        }
        FC_STAMP(7 + 6 * pi);
        fc_cluster_sync();
    }
    FC_STAMP(2 + 6 * nblk);

    // ---- z = W y (one warp per row), alpha = W^T z (one thread per column), fragment packing
    for (int i = rank * (kFcThreads / 32) + warp; i < fp.Npad; i += csize * (kFcThreads / 32)) {
        double part = 0.0;
        if (i < N)
            for (int k = lane; k <= i; k += 32) part = fma(__ldcg(fp.W + (size_t)i * ld + k), fp.Y[k], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) fp.zvec[i] = i < N ? part : 0.0;
    }
    fc_cluster_sync();
    // alpha_c = sum_{i >= c} W[i][c] z_i: eight threads per column (i = c + sub, + 8, ...), sixteen loads in flight per thread --
    // a single thread per column walks up to N dependent-latency L2 loads and took ~30 us at N = 256
    for (int base = 0; base < fp.Npad; base += gthreads >> 3) {      // uniform trip count: the shuffles below need whole warps
        const int c = base + (gtid >> 3), sub = tid & 7;
        double s0 = 0.0, s1 = 0.0;
        if (c < N) {
            int i = c + sub;
#pragma unroll 4
            for (; i + 8 < N; i += 16) {
                s0 = fma(__ldcg(fp.W + (size_t)i * ld + c), __ldcg(fp.zvec + i), s0);
                s1 = fma(__ldcg(fp.W + (size_t)(i + 8) * ld + c), __ldcg(fp.zvec + i + 8), s1);
            }
            if (i < N) s0 = fma(__ldcg(fp.W + (size_t)i * ld + c), __ldcg(fp.zvec + i), s0);
        }
        double sum = s0 + s1;
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        if (sub == 0 && c < fp.Npad) fp.alpha[c] = sum;
    }
    {
        const size_t nfrag = tri_blocks(fp.NB) * 32;
        for (size_t e = gtid; e < nfrag + 4 * 32; e += gthreads) {
            double2 v = make_double2(0.0, 0.0);
            if (e < nfrag) {
                const size_t blk = e >> 5;
                const int l = (int)(e & 31);
                int i = (int)((sqrt(8.0 * (double)blk + 1.0) - 1.0) * 0.5);
                while ((size_t)(i + 1) * (i + 2) / 2 <= blk) ++i;
                while ((size_t)i * (i + 1) / 2 > blk) --i;
                const int kb = (int)(blk - (size_t)i * (i + 1) / 2);
                const int r = 8 * i + (l >> 2), c0 = 8 * kb + 2 * (l & 3);
                if (r < N) {
                    if (c0 < N && c0 <= r) v.x = __ldcg(fp.W + (size_t)r * ld + c0);
                    if (c0 + 1 < N && c0 + 1 <= r) v.y = __ldcg(fp.W + (size_t)r * ld + c0 + 1);
                }
            }
            fp.Afrag[e] = v;
        }
    }
    FC_STAMP(3 + 6 * nblk);
    if (rank == 0 && tid == 0 && sBad) { *fp.status = SO_ERR_NOT_PD; __threadfence_system(); }
}

}  // namespace
