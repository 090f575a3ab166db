// Device building blocks shared by the posterior kernel (posterior.cu) and the batched expander
// test (expander.cu): tile parameters, shared-memory layout, the two kernel-row generators that
// write the k(x*, X) tile straight into DMMA B-fragment order, and the DMMA K-segment loop.
#pragma once
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kGridMaxDim = 6;   // grid fast path (larger d uses explicit rows)

struct PostParams {
    int N, NB, d, RG, CG, T, TB, npass, kind;
    const double2* Afrag;
    const double* alpha;
    const double* Xs;
    double inv_ls[SO_MAX_DIM];
    double variance;
    const double* Xstar;
    int64_t M, row0, ntiles;
    int gd;
    int gn[kGridMaxDim];
    int goff[kGridMaxDim];
    int64_t gstride[kGridMaxDim];
    const double* E;
    double beta, fmin;
    double* mean;
    double* var;
    double* Q;
    int q_stride, q_col;
    uint8_t* S;
    int safe_mode;
};

struct SmemLayout {
    size_t k_bytes, alpha_off, xs_off, xt_off, mean_off, ss_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(int NB, int T, int d, int RG, bool grid) {
    SmemLayout L;
    const size_t Npad = 8 * (size_t)NB;
    L.k_bytes = Npad * T * sizeof(double);
    L.alpha_off = L.k_bytes;
    L.xs_off = L.alpha_off + Npad * sizeof(double);
    L.xt_off = L.xs_off + (grid ? 0 : Npad * d * sizeof(double));
    L.mean_off = L.xt_off + (grid ? 0 : 2 * (size_t)T * d * sizeof(double));
    L.ss_off = L.mean_off + 2 * (size_t)T * sizeof(double);
    L.total = L.ss_off + 2 * (size_t)RG * T * sizeof(double);
    return L;
}

// ---------------------------------------------------------------- gen: explicit rows
template <int KIND>
__device__ __forceinline__ void gen_rows(const PostParams& p, double2* __restrict__ sK, const double* __restrict__ sAlpha,
                                         const double* __restrict__ sXs, const double* __restrict__ sXt,
                                         double* __restrict__ sMean, int warp, int lane) {
    const int d = p.d, N = p.N, NB = p.NB, TB = p.TB;
    const int q = lane & 3, tl = lane >> 2;
    for (int ct = warp; ct < TB; ct += kWarps) {
        const int t = ct * 8 + tl;
        const double* xt = sXt + t * d;
        double m = 0.0;
        for (int kb = 0; kb < NB; ++kb) {
            const int n0 = 8 * kb + 2 * q;
            const double* x0 = sXs + n0 * d;
            double r0 = 0.0, r1 = 0.0;
            for (int j = 0; j < d; ++j) {
                const double xv = xt[j];
                const double t0 = xv - x0[j], t1 = xv - x0[d + j];
                r0 = fma(t0, t0, r0);
                r1 = fma(t1, t1, r1);
            }
            const double k0 = n0 < N ? kernel_of_r2<KIND>(r0, p.variance) : 0.0;
            const double k1 = n0 + 1 < N ? kernel_of_r2<KIND>(r1, p.variance) : 0.0;
            m = fma(k0, sAlpha[n0], m);
            m = fma(k1, sAlpha[n0 + 1], m);
            sK[(kb * TB + ct) * 32 + lane] = make_double2(k0, k1);
        }
        m += __shfl_xor_sync(0xffffffffu, m, 1);
        m += __shfl_xor_sync(0xffffffffu, m, 2);
        if (q == 0) sMean[t] = m;
    }
}

// ---------------------------------------------------------------- gen: separable RBF on a grid
// k(x*, x_n) = prod_j E_j[idx_j(row)][n]; the tables (sum_j n_j rows of Npad doubles, built once per
// fit by k_grid_tables) replace N fp64 exp() per row by (d-1) multiplies -- exp costs ~21 FMA
// slots of the one FP64 pipe the contraction also needs (profiles/r01_fp64_rates_b200.jsonl).
__device__ __forceinline__ void gen_grid(const PostParams& p, double2* __restrict__ sK, const double* __restrict__ sAlpha,
                                         double* __restrict__ sMean, int64_t tile_row0, int warp, int lane) {
    const int NB = p.NB, TB = p.TB, Npad = 8 * p.NB;
    const int q = lane & 3, tl = lane >> 2;
    const double2* sA2 = reinterpret_cast<const double2*>(sAlpha);
    const int64_t last = p.row0 + p.M - 1;
    for (int ct = warp; ct < TB; ct += kWarps) {
        const int t = ct * 8 + tl;
        int64_t row = tile_row0 + t;
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
        double m = 0.0;
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
            const double2 a = sA2[4 * kb + q];
            m = fma(v.x, a.x, m);
            m = fma(v.y, a.y, m);
            sK[(kb * TB + ct) * 32 + lane] = v;
        }
        m += __shfl_xor_sync(0xffffffffu, m, 1);
        m += __shfl_xor_sync(0xffffffffu, m, 2);
        if (q == 0) sMean[t] = m;
    }
}

// ---------------------------------------------------------------- mma
// One K segment: accumulator slots FIRST..3 are active.  `a` holds the fragments of the current
// k-block on entry and those of k-block kb_hi+1 on exit (software prefetch, distance one block).
template <int BT, int FIRST>
__device__ __forceinline__ void mma_segment(double (&acc)[4][BT][2], double2 (&a)[4], const double2* __restrict__ Afrag,
                                            const size_t (&abase)[4], const double2* __restrict__ sB, int TB,
                                            int kb_lo, int kb_hi) {
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

__device__ __forceinline__ int pick4(int r0, int r1, int r2, int r3, int i) {
    return i == 0 ? r0 : (i == 1 ? r1 : (i == 2 ? r2 : r3));
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
