// f1 -- SafeOptSwarm safe-set maintenance on the device (safeopt/gp_opt.py:1088-1110).
//
// The reference builds the dense P x (|S| + P) prior-correlation matrix of the swarm's best positions
// against the safe set and themselves (80 GB at BASELINE config 5) and then walks the particles in index
// order, accepting particle j iff corr(j, i) <= 0.95 for every i already in the set (old safe points and
// particles accepted before j).  Here nothing P x P is stored:
//   so_safeset_filter : keep[j] = all_i corr(cand_j, ref_i) <= thresh          (fully parallel, tiled)
//   so_safeset_insert : the sequential-accept walk in blocks of kTB candidates: each block is first
//                       checked against everything accepted in earlier blocks (parallel, k_vs_accepted),
//                       then resolved internally by ONE CTA whose thread t owns candidate t and drops out
//                       as soon as an accepted lower-index candidate is too close (k_resolve).
// corr = k(x, x') / scale2 with the stationary kernel of GP `gp` (kind, ARD lengthscales, variance) as
// fitted on the handle; `scale2` is the host's scaling[0]**2 (gp_opt.py:1095).  Both kernels are bound by
// the fp64 pipe (distance + profile per pair); pairs that are obviously uncorrelated (r^2 beyond the
// host-computed r2_skip, where corr < thresh / 2) skip the exponential.
#include "common.cuh"

namespace {

constexpr int kTB = 1024;        // candidates per sequential block (one thread each in k_resolve)
constexpr int kRefChunk = 256;   // reference rows staged in shared memory per CTA

struct CorrParams {
    int d;
    double variance, scale2, thresh, r2_skip;
    double inv_ls[SO_MAX_DIM];
};

template <int KIND>
__device__ __forceinline__ bool too_close(double r2, const CorrParams& p) {
    if (r2 > p.r2_skip) return false;
    return __ddiv_rn(kernel_of_r2<KIND>(r2, p.variance), p.scale2) > p.thresh;
}

// keep[j] (j in [j0, j0+nb)) is cleared when candidate j is too close to any of the first m reference rows;
// m = *m_dev when m_dev != NULL (the accepted list grows on the device), capped by m_cap.
// grid = (ceil(nb / 256), chunks of kRefChunk references).
template <int KIND>
__global__ void __launch_bounds__(256) k_vs_refs(CorrParams p, const double* __restrict__ cand, int64_t j0, int64_t nb,
                                                 const double* __restrict__ ref, int64_t m_cap, const int64_t* __restrict__ m_dev,
                                                 uint8_t* __restrict__ keep) {
    __shared__ double sref[kRefChunk * SO_MAX_DIM];
    const int64_t m = m_dev ? (*m_dev < m_cap ? *m_dev : m_cap) : m_cap;
    const int64_t r0 = (int64_t)blockIdx.y * kRefChunk;
    if (r0 >= m) return;
    const int cnt = (int)((m - r0) < kRefChunk ? (m - r0) : kRefChunk);
    const int d = p.d;
    for (int e = threadIdx.x; e < cnt * d; e += 256) sref[e] = ref[(size_t)r0 * d + e] * p.inv_ls[e % d];
    __syncthreads();
    const int64_t j = j0 + (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= j0 + nb || !keep[j]) return;
    double x[SO_MAX_DIM];
#pragma unroll
    for (int k = 0; k < SO_MAX_DIM; ++k) x[k] = k < d ? cand[(size_t)j * d + k] * p.inv_ls[k] : 0.0;
    bool hit = false;
    for (int i = 0; i < cnt && !hit; ++i) {
        double r2 = 0.0;
#pragma unroll
        for (int k = 0; k < SO_MAX_DIM; ++k) {
            if (k < d) { const double t = x[k] - sref[i * d + k]; r2 = fma(t, t, r2); }
        }
        hit = too_close<KIND>(r2, p);
    }
    if (hit) keep[j] = 0;
}

// One CTA, kTB threads: candidates j0 .. j0+nb-1 in index order.  alive[t] starts as keep[j0+t]; candidate i,
// once final and alive, removes every later alive candidate that is too close.  Survivors are appended (in
// index order) to acc_pos and counted in *n_acc; accept[j] is written for the whole block.
template <int KIND>
__global__ void __launch_bounds__(kTB) k_resolve(CorrParams p, const double* __restrict__ cand, int64_t j0, int nb,
                                                 const uint8_t* __restrict__ keep, uint8_t* __restrict__ accept,
                                                 double* __restrict__ acc_pos, int64_t* __restrict__ n_acc) {
    extern __shared__ double spos[];                  // nb x d scaled coordinates
    __shared__ uint8_t alive[kTB];
    __shared__ int warp_count[kTB / 32];
    const int t = threadIdx.x, d = p.d;
    const int64_t j = j0 + t;
    double x[SO_MAX_DIM];
#pragma unroll
    for (int k = 0; k < SO_MAX_DIM; ++k) x[k] = 0.0;
    bool mine = false;
    if (t < nb) {
        mine = keep[j] != 0;
#pragma unroll
        for (int k = 0; k < SO_MAX_DIM; ++k) {
            if (k < d) { x[k] = cand[(size_t)j * d + k] * p.inv_ls[k]; spos[t * d + k] = x[k]; }
        }
    }
    alive[t] = mine ? 1 : 0;
    __syncthreads();
    for (int i = 0; i < nb; ++i) {
        if (alive[i]) {                               // final: every write to alive[i] was followed by a barrier
            if (t > i && mine) {
                double r2 = 0.0;
#pragma unroll
                for (int k = 0; k < SO_MAX_DIM; ++k) {
                    if (k < d) { const double u = x[k] - spos[i * d + k]; r2 = fma(u, u, r2); }
                }
                if (too_close<KIND>(r2, p)) { mine = false; alive[t] = 0; }
            }
            __syncthreads();
        }
    }
    // ordered compaction
    const unsigned ballot = __ballot_sync(0xffffffffu, mine);
    const int lane = t & 31, warp = t >> 5;
    if (lane == 0) warp_count[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kTB / 32; ++w) {
        const int c = warp_count[w];
        if (w < warp) before += c;
        total += c;
    }
    const int64_t base = *n_acc;
    if (t < nb) accept[j] = mine ? 1 : 0;
    if (mine) {
        const int64_t slot = base + before + __popc(ballot & ((1u << lane) - 1u));
        for (int k = 0; k < d; ++k) acc_pos[(size_t)slot * d + k] = cand[(size_t)j * d + k];
    }
    __syncthreads();
    if (t == 0) *n_acc = base + total;
}

__global__ void k_fill_u8(uint8_t* p, int64_t n, uint8_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

double profile_host(int kind, double r2) {
    if (kind == SO_KERNEL_RBF) return std::exp(-0.5 * r2);
    const double r = std::sqrt(r2);
    if (kind == SO_KERNEL_MATERN32) return (1.0 + 1.7320508075688772 * r) * std::exp(-1.7320508075688772 * r);
    return (1.0 + 2.23606797749979 * r + (5.0 / 3.0) * r2) * std::exp(-2.23606797749979 * r);
}

// Largest r^2 beyond which corr < thresh/2 for sure (monotone profiles): pairs farther apart skip the exp.
double skip_radius2(int kind, double variance, double scale2, double thresh) {
    const double target = 0.5 * thresh * scale2 / variance;
    if (!(target > 0.0)) return INFINITY;             // thresh <= 0: evaluate every pair
    if (profile_host(kind, 0.0) < target) return -1.0;   // nothing can reach thresh/2: skip all
    double lo = 0.0, hi = 1.0;
    while (profile_host(kind, hi) >= target && hi < 1e12) hi *= 2.0;
    for (int it = 0; it < 80; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (profile_host(kind, mid) >= target) lo = mid; else hi = mid;
    }
    return hi;
}

int make_params(so_handle* h, int gp, double scale2, double thresh, CorrParams& p, const char* where) {
    if (gp < 0 || gp >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, std::string(where) + ": bad gp index");
    const GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, std::string(where) + ": GP not fitted");
    if (!(scale2 > 0.0)) return so_fail(h, SO_ERR_BAD_ARG, std::string(where) + ": scale2 must be positive");
    p.d = g.d;
    p.variance = g.variance;
    p.scale2 = scale2;
    p.thresh = thresh;
    p.r2_skip = skip_radius2(g.kind, g.variance, scale2, thresh);
    for (int k = 0; k < SO_MAX_DIM; ++k) p.inv_ls[k] = k < g.d ? g.inv_ls[k] : 0.0;
    return SO_OK;
}

template <int KIND>
int launch_vs_refs(so_handle* h, const CorrParams& p, const double* cand, int64_t j0, int64_t nb, const double* ref,
                   int64_t m_cap, const int64_t* m_dev, uint8_t* keep, cudaStream_t stream) {
    if (nb <= 0 || m_cap <= 0) return SO_OK;
    const int64_t chunks = (m_cap + kRefChunk - 1) / kRefChunk;
    if (chunks > 65535) return so_fail(h, SO_ERR_CAPACITY, "safe-set kernels: at most 16.7M reference rows");
    dim3 grid((unsigned)((nb + 255) / 256), (unsigned)chunks);
    k_vs_refs<KIND><<<grid, 256, 0, stream>>>(p, cand, j0, nb, ref, m_cap, m_dev, keep);
    SO_CHECK_LAUNCH(h, "k_vs_refs");
    return SO_OK;
}

}  // namespace

extern "C" int so_safeset_filter(so_handle* h, int gp, const double* cand_d, int64_t n, const double* ref_d, int64_t m,
                                 double scale2, double thresh, uint8_t* keep_d, void* stream_) {
    if (!h || n < 0 || m < 0 || (n > 0 && (!cand_d || !keep_d)) || (m > 0 && !ref_d)) return SO_ERR_BAD_ARG;
    CorrParams p;
    int rc = make_params(h, gp, scale2, thresh, p, "so_safeset_filter");
    if (rc != SO_OK) return rc;
    if (n == 0) return SO_OK;
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    k_fill_u8<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keep_d, n, 1);
    SO_CHECK_LAUNCH(h, "k_fill_u8");
    switch (h->gps[gp].kind) {
        case SO_KERNEL_RBF: return launch_vs_refs<SO_KERNEL_RBF>(h, p, cand_d, 0, n, ref_d, m, nullptr, keep_d, stream);
        case SO_KERNEL_MATERN32: return launch_vs_refs<SO_KERNEL_MATERN32>(h, p, cand_d, 0, n, ref_d, m, nullptr, keep_d, stream);
        default: return launch_vs_refs<SO_KERNEL_MATERN52>(h, p, cand_d, 0, n, ref_d, m, nullptr, keep_d, stream);
    }
}

template <int KIND>
static int run_insert(so_handle* h, const CorrParams& p, const double* cand, int64_t n, uint8_t* keep, uint8_t* accept,
                      double* acc_pos, int64_t* n_acc, cudaStream_t stream) {
    const size_t smem = sizeof(double) * (size_t)kTB * p.d;
    if (smem > 32 * 1024) {     // dynamic + ~1.2 KB static must stay under the 48 KB default, else opt in
        cudaError_t e = cudaFuncSetAttribute(k_resolve<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return so_fail(h, SO_ERR_CUDA, std::string("k_resolve smem: ") + cudaGetErrorString(e));
    }
    for (int64_t j0 = 0; j0 < n; j0 += kTB) {
        const int nb = (int)(n - j0 < kTB ? n - j0 : kTB);
        // against everything accepted in earlier blocks (at most j0 rows; the true count is read on the device)
        const int rc = launch_vs_refs<KIND>(h, p, cand, j0, nb, acc_pos, j0, n_acc, keep, stream);
        if (rc != SO_OK) return rc;
        k_resolve<KIND><<<1, kTB, smem, stream>>>(p, cand, j0, nb, keep, accept, acc_pos, n_acc);
    }
    SO_CHECK_LAUNCH(h, "k_resolve");
    return SO_OK;
}

extern "C" int so_safeset_insert(so_handle* h, int gp, const double* cand_d, int64_t n, uint8_t* keep_d, double scale2,
                                 double thresh, uint8_t* accept_d, double* accepted_pos_d, int64_t* n_accept_d,
                                 void* stream_) {
    if (!h || n < 0 || !n_accept_d || (n > 0 && (!cand_d || !keep_d || !accept_d || !accepted_pos_d))) return SO_ERR_BAD_ARG;
    if (n > (int64_t)65535 * kRefChunk) return so_fail(h, SO_ERR_CAPACITY, "so_safeset_insert: at most 16.7M candidates");
    CorrParams p;
    int rc = make_params(h, gp, scale2, thresh, p, "so_safeset_insert");
    if (rc != SO_OK) return rc;
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    SO_CUDA(h, cudaMemsetAsync(n_accept_d, 0, sizeof(int64_t), stream));
    if (n == 0) return SO_OK;
    switch (h->gps[gp].kind) {
        case SO_KERNEL_RBF: return run_insert<SO_KERNEL_RBF>(h, p, cand_d, n, keep_d, accept_d, accepted_pos_d, n_accept_d, stream);
        case SO_KERNEL_MATERN32: return run_insert<SO_KERNEL_MATERN32>(h, p, cand_d, n, keep_d, accept_d, accepted_pos_d, n_accept_d, stream);
        default: return run_insert<SO_KERNEL_MATERN52>(h, p, cand_d, n, keep_d, accept_d, accepted_pos_d, n_accept_d, stream);
    }
}
