// Lipschitz expander test (the original SafeOpt rule), safeopt/gp_opt.py:558-576:
//     d        = cdist(x_c, inputs[~S])                       (Euclidean, raw inputs incl. contexts)
//     G_safe_c = any(u_c,i - L_i * d >= fmin_i)               for every constrained GP i (AND over GPs on the host)
// One streaming sweep over the rows for a batch of up to 32 candidates: 8*d bytes per unsafe row on the
// explicit-rows path (0 on the grid path), B*(3d+4) flops per unsafe row -- HBM/latency-bound, no GP arithmetic.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBatch = 32;

struct LipGrid {
    int d;
    int n[6];
    int off[6];
    int64_t stride[6];
};

__global__ void __launch_bounds__(kThreads) k_lipschitz(const double* __restrict__ X, LipGrid lg, const double* __restrict__ axis,
                                                       int d, int64_t row0, int64_t M, const uint8_t* __restrict__ S,
                                                       const double* __restrict__ xc, const double* __restrict__ uc, int B,
                                                       double lipschitz, double fmin, uint8_t* __restrict__ flags) {
    __shared__ double s_xc[kMaxBatch * SO_MAX_DIM];
    __shared__ double s_uc[kMaxBatch];
    __shared__ unsigned s_hit;
    for (int i = threadIdx.x; i < B * d; i += kThreads) s_xc[i] = xc[i];
    for (int i = threadIdx.x; i < B; i += kThreads) s_uc[i] = uc[i];
    if (threadIdx.x == 0) s_hit = 0u;
    __syncthreads();
    unsigned hit = 0u;
    for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < M; r += (int64_t)gridDim.x * kThreads) {
        if (S[r]) continue;
        double x[SO_MAX_DIM];
#pragma unroll
        for (int j = 0; j < SO_MAX_DIM; ++j) {
            if (j < d) {
                if (X) {
                    x[j] = X[(size_t)r * d + j];
                } else {
                    const int jj = j < 6 ? j : 0;
                    const int idx = (int)(((row0 + r) / lg.stride[jj]) % lg.n[jj]);
                    x[j] = axis[lg.off[jj] + idx];
                }
            } else {
                x[j] = 0.0;
            }
        }
        for (int b = 0; b < B; ++b) {
            double s2 = 0.0;
#pragma unroll
            for (int j = 0; j < SO_MAX_DIM; ++j)
                if (j < d) {
                    const double t = s_xc[b * d + j] - x[j];
                    s2 = fma(t, t, s2);
                }
            const double dist = sqrt(s2);
            if (__dsub_rn(s_uc[b], __dmul_rn(lipschitz, dist)) >= fmin) hit |= 1u << b;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hit |= __shfl_xor_sync(0xffffffffu, hit, o);
    if ((threadIdx.x & 31) == 0 && hit) atomicOr(&s_hit, hit);
    __syncthreads();
    if (threadIdx.x < B && (s_hit >> threadIdx.x & 1u)) flags[threadIdx.x] = 1;   // benign race: all writers store 1
}

}  // namespace

extern "C" int so_expander_lipschitz(so_handle* h, const double* Xstar_d, int d, int64_t row0, int64_t M, const uint8_t* S_d,
                                     const double* xc_d, const double* u_c_d, int B, double lipschitz, double fmin,
                                     uint8_t* flags_d, void* stream) {
    if (!h) return SO_ERR_BAD_ARG;
    if (!S_d || !xc_d || !u_c_d || !flags_d || M < 0) return so_fail(h, SO_ERR_BAD_ARG, "expander_lipschitz: null argument");
    if (B < 1 || B > kMaxBatch) return so_fail(h, SO_ERR_BAD_ARG, "expander_lipschitz: batch must be in [1, 32]");
    LipGrid lg;
    lg.d = 0;
    for (int j = 0; j < 6; ++j) { lg.n[j] = 1; lg.off[j] = 0; lg.stride[j] = 1; }
    if (!Xstar_d) {
        const GridSpec& gs = h->grid;
        if (!gs.defined) return so_fail(h, SO_ERR_BAD_ARG, "expander_lipschitz: no rows and no grid defined");
        if (row0 < 0 || row0 + M > gs.rows) return so_fail(h, SO_ERR_BAD_ARG, "expander_lipschitz: rows outside the grid");
        d = gs.d;
        lg.d = gs.d;
        for (int j = 0; j < gs.d; ++j) { lg.n[j] = gs.n[j]; lg.off[j] = gs.off[j]; lg.stride[j] = gs.stride[j]; }
    }
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "expander_lipschitz: 1 <= d <= 16");
    if (M == 0) return SO_OK;
    DeviceGuard guard(h->device);
    int64_t blocks = (M + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)h->num_sms * 8;
    if (blocks > cap) blocks = cap;
    k_lipschitz<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(Xstar_d, lg, h->grid.axis, d, row0, M, S_d, xc_d, u_c_d, B,
                                                                        lipschitz, fmin, flags_d);
    SO_CHECK_LAUNCH(h, "k_lipschitz");
    return SO_OK;
}
