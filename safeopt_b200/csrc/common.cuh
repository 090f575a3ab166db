// Internal definitions shared by the sm_100a translation units behind include/safeopt_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/safeopt_b200.h"
#include "fastexp.cuh"

#define SO_MAX_DIM 16            // max input dimension (parameters + contexts)
#define SO_JITTER 1e-8           // GPy adds this to the noise variance
#define SO_VAR_FLOOR 1e-15       // GPy clips predictive variances here
#define SO_WS_MAX_BLOCKS 1184     // 8 x 148: upper bound on the grid of the reduction passes

// Fitted state of one GP, all in device memory (fp64).
struct GPState {
    bool fitted = false;
    int N = 0, d = 0, kind = 0;
    int NB = 0;                  // number of 8-row blocks, Npad = 8*NB
    int capN = 0;                // allocated for this many padded rows
    int ld = 0;                  // leading dimension of K / Linv (= capN: stays put while points are appended)
    double variance = 0, noise = 0;
    double inv_ls[SO_MAX_DIM];   // 1 / lengthscale_j
    double* X = nullptr;         // N x d raw training inputs
    double* Xs = nullptr;        // Npad x d, scaled by 1/lengthscale (padding rows = 0)
    double* Y = nullptr;         // N
    double* K = nullptr;         // Npad x Npad work / L (lower, row-major, leading dimension ld)
    double* Linv = nullptr;      // Npad x Npad lower, row-major, leading dimension ld
    double* alpha = nullptr;     // Npad (padding = 0)
    double* zvec = nullptr;      // Npad: z = L^-1 y (padding = 0); mean(x*) = (L^-1 k).z, taken from the same accumulators as |L^-1 k|^2
    double2* Afrag = nullptr;    // L^-1 packed in DMMA A-fragment order (+4 blocks of slack for the prefetch)
    // grid tables (grid fast path): per axis j, E_j[i][n], i < n_j, n < Npad
    double* E = nullptr;
    size_t capE = 0;
    // two-level product tables of the grid path: fast_rows + slow_rows rows of Npad doubles (Pfast, Pslow)
    double* P2 = nullptr;
    size_t capP2 = 0;
    // tables of the TMA kernel: fragment-ordered fast table, scaled operands A'(s) (see posterior_tma.cuh)
    double2* PfFrag = nullptr;
    size_t capPfFrag = 0;
    double2* Aprime = nullptr;
    size_t capAprime = 0;
    size_t a_stride = 0;
    int64_t ap_s0 = 0, ap_s1 = 0;   // slow indices [ap_s0, ap_s1) resident in Aprime (the rows this rank evaluates)
    int tma_T = 0, tma_tpb = 0, tma_BT = 0, tma_RG = 0, tma_CG = 0, tma_kb_pad = 0, tma_warps = 8;
    bool tma_ready = false;
    bool tma_ns2x = false;       // N <= 128: two block rows per warp, two CTAs per SM (k_posterior_tma<.., .., true>)
    bool tma_ns6 = false;        // 33..48 block rows on a 32-row tile: six block rows per warp, one pass
    bool tma_ring = false;       // B streams through the k-chunk ring (posterior_ring.cuh) instead of a resident double buffer
    bool grid_ready = false;
    // fp32 arithmetic mode (posterior_f32.cuh): TF32 hi/lo operand planes in UMMA canonical layout
    float* f32_A = nullptr;      // fast-table tiles
    size_t capF32A = 0;
    float* f32_B = nullptr;      // scaled operands A'(s) for the slow indices [f32_s0, f32_s1)
    size_t capF32B = 0;
    double* f32_PfT = nullptr;   // transposed fp64 fast table of the mean kernel
    size_t capF32PfT = 0;
    int64_t f32_Fpad = 0;
    int f32_Np = 0, f32_tpb = 0;
    int64_t f32_s0 = 0, f32_s1 = 0;
    bool f32_ready = false;
};

struct GridSpec {
    bool defined = false;
    int d = 0;
    int n[SO_MAX_DIM];
    int64_t stride[SO_MAX_DIM];  // row = sum_j idx_j * stride_j  (reference row order)
    int off[SO_MAX_DIM];         // offset of axis j in `axis` / table rows
    int total = 0;               // sum n_j
    double* axis = nullptr;      // device copy of the axis values
    int cap = 0;
    int64_t rows = 0;
    int in_fast[SO_MAX_DIM];     // axis belongs to the fast (low-order) group of the product tables
    int64_t fast_rows = 1, slow_rows = 1;
};

struct XchgBuf;                  // xchg.cuh: per-rank exchange buffer of the cross-rank record protocols

struct so_handle {
    int device = 0;
    int max_gps = 0;
    int num_sms = 0;
    int smem_optin = 0;
    std::vector<GPState> gps;
    GridSpec grid;
    int* d_status = nullptr;     // device int written by fit kernels
    int* h_status = nullptr;     // pinned (mapped) host mirror
    int* status_mapped_d = nullptr;  // device-side address of h_status (the one-launch fit writes its status straight to the host)
    void* fit_stage_h = nullptr; // pinned staging buffer for X, Y of a fit, one slot of fit_stage_bytes per GP
    size_t fit_stage_bytes = 0;
    int* fit_status_h = nullptr; // mapped pinned status word per GP, written by the one-launch fit kernel
    int* fit_status_d = nullptr; // its device-side address
    void* ws_partials = nullptr; // per-block partial records of the reduction passes (sets.cu)
    unsigned int* ws_counter = nullptr;  // last-block-done ticket, self-resetting
    double* ws_z = nullptr;      // expander batch workspace (expander.cu)
    size_t ws_z_cap = 0;
    // cross-rank record exchange (xchg.cu): this rank's buffer and the peers' buffers as mapped here (peer[rank] = local)
    XchgBuf* xchg_local = nullptr;
    XchgBuf* xchg_peer[16] = {};
    bool xchg_opened[16] = {};
    int xchg_world = 1, xchg_rank = 0;
    unsigned long long* xchg_epochs = nullptr;   // device: [0] set-pass epoch, [1] swarm epoch, [2] swarm iteration counter
    // fused set pass (sets.cu: k_sets_fused): grid barrier {count, generation}, two partial arrays, candidate counter,
    // pinned host mirror of the combined records
    unsigned int* fused_bar = nullptr;
    void* fused_part = nullptr;
    unsigned long long* fused_ncand = nullptr;
    int* fused_status = nullptr;
    long long* fused_dbg = nullptr;
    double* f32_mean_scratch = nullptr;   // fp64 means of the fp32 mode when the caller wants none (G x M)
    size_t f32_mean_cap = 0;
    void* fused_result_h = nullptr;  // world x 136 bytes + status + epoch stamp: mapped pinned host memory
    void* fused_result_d = nullptr;  // its device-side address (zero copy)
    int fused_grid = 0;
    int fused_inflight = 0;          // so_sets_fused launches without a fetch since the last synchronising fetch
    std::string err;
};

inline int so_fail(so_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define SO_CUDA(h, expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return so_fail((h), SO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define SO_CHECK_LAUNCH(h, what)                                                                \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return so_fail((h), SO_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- device helpers -------------------------------------------------------------------
// fp64 tensor-core tile: D(8x8) += A(8x4) * B(4x8).  Lane l holds A[l/4][l%4], B[l%4][l/4],
// C[l/4][2*(l%4) + {0,1}].  SASS: DMMA.8x8x4 (measured 37.1 TFLOP/s on B200, profiles/r01_fp64_rates_b200.jsonl).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Stationary kernel value from the squared scaled distance (GPy K_of_r).
template <int KIND>
__device__ __forceinline__ double kernel_of_r2(double r2, double variance) {
    if (KIND == SO_KERNEL_RBF) {
        return variance * exp(-0.5 * r2);
    } else if (KIND == SO_KERNEL_MATERN32) {
        const double s3 = 1.7320508075688772;
        double r = sqrt(r2);
        return variance * (1.0 + s3 * r) * exp(-s3 * r);
    } else {
        const double s5 = 2.23606797749979;
        double r = sqrt(r2);
        return variance * (1.0 + s5 * r + (5.0 / 3.0) * r2) * exp(-s5 * r);
    }
}

// Same value through so_exp_neg (fastexp.cuh): ~11 fp64 operations per exp instead of libdevice's ~21.  `T` is the 64-entry
// 2^(j/64) table in shared memory (load_exp_table).  Used by the per-row generators; one-off kernels keep exp().
template <int KIND>
__device__ __forceinline__ double kernel_of_r2_fast(double r2, double variance, const double* __restrict__ T) {
    if (KIND == SO_KERNEL_RBF) {
        return variance * so_exp_neg(-0.5 * r2, T);
    } else if (KIND == SO_KERNEL_MATERN32) {
        const double s3 = 1.7320508075688772;
        double r = sqrt(r2);
        return variance * (1.0 + s3 * r) * so_exp_neg(-s3 * r, T);
    } else {
        const double s5 = 2.23606797749979;
        double r = sqrt(r2);
        return variance * (1.0 + s5 * r + (5.0 / 3.0) * r2) * so_exp_neg(-s5 * r, T);
    }
}

static __constant__ double c_exp_table[64] = {SO_EXP_TABLE_VALUES};

// Copies the exp table into shared memory (divergent table indices would serialise on the constant cache).
// The caller synchronises the block before the first use.
__device__ __forceinline__ void load_exp_table(double* __restrict__ sT) {
    if (threadIdx.x < 64) sT[threadIdx.x] = c_exp_table[threadIdx.x];
}

// Number of DMMA A-fragment blocks in the packed lower triangle of an NB x NB block matrix.
__host__ __device__ inline size_t tri_blocks(int NB) { return (size_t)NB * (NB + 1) / 2; }

// Ordered-integer encoding so that atomicMax on int64 orders doubles (incl. -inf).
__device__ __forceinline__ long long order_encode(double x) {
    long long b = __double_as_longlong(x);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double order_decode(long long k) {
    long long b = k >= 0 ? k : (k ^ 0x7fffffffffffffffLL);
    return __longlong_as_double(b);
}
