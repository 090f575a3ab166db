// K3 -- the set logic of SafeOpt.compute_sets as HBM-streaming passes over Q (fp64, row-major,
// 2G columns) and the byte masks.  Each pass is one coalesced sweep (16*G + 1..3 bytes per row)
// with a warp-shuffle + shared-memory block reduction and a last-block-done final combine, so
// results are deterministic (first-row tie-breaks follow NumPy's argmax, gp_opt.py:635,:644,:710).
//   so_sets_reduce_safe : gp_opt.py:504 (any S), :512 (max l0[S]), :634-636, :708-712
//   so_sets_maximizers  : gp_opt.py:511-513 and the M part of :642-644
//   so_sets_candidates  : gp_opt.py:531-536 and the sort key of :551
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = SO_WS_MAX_BLOCKS;

struct SafePartial { long long n; double max_l; long long arg_l; double max_u; long long arg_u; };
struct MaxPartial { long long n; double max_w0; double best; long long best_row; };

__device__ __forceinline__ void take_max_first(double& v, long long& r, double ov, long long orow) {
    // keep the larger value; on ties keep the smaller (global) row; row < 0 means "empty"
    if (orow >= 0 && (r < 0 || ov > v || (ov == v && orow < r))) { v = ov; r = orow; }
}

__global__ void __launch_bounds__(kThreads) k_reduce_safe(const double* __restrict__ Q, int q_stride, int64_t M, int64_t row0,
                                                         const uint8_t* __restrict__ S, SafePartial* __restrict__ part,
                                                         unsigned int* __restrict__ counter, so_safe_record* __restrict__ out) {
    SafePartial acc = {0, -INFINITY, -1, -INFINITY, -1};
    for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < M; r += (int64_t)gridDim.x * kThreads) {
        if (S[r]) {
            const double l = Q[(size_t)r * q_stride], u = Q[(size_t)r * q_stride + 1];
            acc.n += 1;
            take_max_first(acc.max_l, acc.arg_l, l, row0 + r);
            take_max_first(acc.max_u, acc.arg_u, u, row0 + r);
        }
    }
    __shared__ SafePartial sm[kThreads / 32];
    __shared__ bool last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        SafePartial other;
        other.n = __shfl_xor_sync(0xffffffffu, acc.n, o);
        other.max_l = __shfl_xor_sync(0xffffffffu, acc.max_l, o);
        other.arg_l = __shfl_xor_sync(0xffffffffu, acc.arg_l, o);
        other.max_u = __shfl_xor_sync(0xffffffffu, acc.max_u, o);
        other.arg_u = __shfl_xor_sync(0xffffffffu, acc.arg_u, o);
        acc.n += other.n;
        take_max_first(acc.max_l, acc.arg_l, other.max_l, other.arg_l);
        take_max_first(acc.max_u, acc.arg_u, other.max_u, other.arg_u);
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        SafePartial t = sm[0];
        for (int w = 1; w < kThreads / 32; ++w) {
            t.n += sm[w].n;
            take_max_first(t.max_l, t.arg_l, sm[w].max_l, sm[w].arg_l);
            take_max_first(t.max_u, t.arg_u, sm[w].max_u, sm[w].arg_u);
        }
        part[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        SafePartial t = {0, -INFINITY, -1, -INFINITY, -1};
        for (unsigned b = 0; b < gridDim.x; ++b) {
            const SafePartial o = part[b];
            t.n += o.n;
            take_max_first(t.max_l, t.arg_l, o.max_l, o.arg_l);
            take_max_first(t.max_u, t.arg_u, o.max_u, o.arg_u);
        }
        out->n_safe = t.n; out->max_l0 = t.max_l; out->argmax_l0 = t.arg_l; out->max_u0 = t.max_u; out->argmax_u0 = t.arg_u;
        out->reserved[0] = out->reserved[1] = out->reserved[2] = 0;
        *counter = 0;
    }
}

struct Scal { double v[64]; };

__global__ void __launch_bounds__(kThreads) k_maximizers(const double* __restrict__ Q, int G, int64_t M, int64_t row0,
                                                        const uint8_t* __restrict__ S, double max_l0, Scal scaling,
                                                        uint8_t* __restrict__ Mmask, MaxPartial* __restrict__ part,
                                                        unsigned int* __restrict__ counter, so_max_record* __restrict__ out) {
    MaxPartial acc = {0, -INFINITY, -INFINITY, -1};
    const int qs = 2 * G;
    for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < M; r += (int64_t)gridDim.x * kThreads) {
        uint8_t m = 0;
        if (S[r]) {
            const double* q = Q + (size_t)r * qs;
            if (q[1] >= max_l0) {
                m = 1;
                acc.n += 1;
                const double w0 = q[1] - q[0];
                acc.max_w0 = w0 > acc.max_w0 ? w0 : acc.max_w0;
                double val = w0 / scaling.v[0];
                for (int i = 1; i < G; ++i) {
                    const double wi = (q[2 * i + 1] - q[2 * i]) / scaling.v[i];
                    val = wi > val ? wi : val;
                }
                take_max_first(acc.best, acc.best_row, val, row0 + r);
            }
        }
        Mmask[r] = m;
    }
    __shared__ MaxPartial sm[kThreads / 32];
    __shared__ bool last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxPartial other;
        other.n = __shfl_xor_sync(0xffffffffu, acc.n, o);
        other.max_w0 = __shfl_xor_sync(0xffffffffu, acc.max_w0, o);
        other.best = __shfl_xor_sync(0xffffffffu, acc.best, o);
        other.best_row = __shfl_xor_sync(0xffffffffu, acc.best_row, o);
        acc.n += other.n;
        acc.max_w0 = other.max_w0 > acc.max_w0 ? other.max_w0 : acc.max_w0;
        take_max_first(acc.best, acc.best_row, other.best, other.best_row);
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        MaxPartial t = sm[0];
        for (int w = 1; w < kThreads / 32; ++w) {
            t.n += sm[w].n;
            t.max_w0 = sm[w].max_w0 > t.max_w0 ? sm[w].max_w0 : t.max_w0;
            take_max_first(t.best, t.best_row, sm[w].best, sm[w].best_row);
        }
        part[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        MaxPartial t = {0, -INFINITY, -INFINITY, -1};
        for (unsigned b = 0; b < gridDim.x; ++b) {
            const MaxPartial o = part[b];
            t.n += o.n;
            t.max_w0 = o.max_w0 > t.max_w0 ? o.max_w0 : t.max_w0;
            take_max_first(t.best, t.best_row, o.best, o.best_row);
        }
        out->n_max = t.n; out->max_width0 = t.max_w0; out->best_value = t.best; out->best_row = t.best_row;
        out->reserved[0] = out->reserved[1] = out->reserved[2] = out->reserved[3] = 0;
        *counter = 0;
    }
}

__global__ void __launch_bounds__(kThreads) k_candidates(const double* __restrict__ Q, int G, int64_t M, int64_t row0,
                                                        const uint8_t* __restrict__ S, const uint8_t* __restrict__ Mmask,
                                                        double max_var, Scal scaling, Scal thr, uint8_t* __restrict__ cmask,
                                                        double* __restrict__ ckey, int64_t* __restrict__ crow, int64_t cap,
                                                        unsigned long long* __restrict__ n_cand) {
    const int qs = 2 * G;
    // whole warps iterate together (the append below uses full-mask warp collectives)
    for (int64_t base = (int64_t)blockIdx.x * kThreads; base < M; base += (int64_t)gridDim.x * kThreads) {
        const int64_t r = base + threadIdx.x;
        uint8_t c = 0;
        double key = 0.0;
        if (r < M && S[r] && !Mmask[r]) {
            const double* q = Q + (size_t)r * qs;
            double smax = -INFINITY, wmax = -INFINITY;
            bool over = false;
            for (int i = 0; i < G; ++i) {
                const double w = q[2 * i + 1] - q[2 * i];
                const double ws = w / scaling.v[i];
                smax = ws > smax ? ws : smax;
                wmax = w > wmax ? w : wmax;
                over = over || (w > thr.v[i]);
            }
            if (smax > max_var && over) { c = 1; key = wmax; }
        }
        if (cmask && r < M) cmask[r] = c;
        const unsigned ballot = __ballot_sync(0xffffffffu, c);
        if (ballot) {
            const unsigned lane = threadIdx.x & 31;
            const int leader = __ffs(ballot) - 1;
            unsigned long long slot0 = 0;
            if ((int)lane == leader) slot0 = atomicAdd(n_cand, (unsigned long long)__popc(ballot));
            slot0 = __shfl_sync(0xffffffffu, slot0, leader);
            const unsigned long long slot = slot0 + __popc(ballot & ((1u << lane) - 1));
            if (c && (int64_t)slot < cap) { ckey[slot] = key; crow[slot] = row0 + r; }
        }
    }
}

int grid_for(const so_handle* h, int64_t M) {
    int64_t blocks = (M + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)h->num_sms * 8;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" int so_sets_reduce_safe(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                                   so_safe_record* rec_d, void* stream) {
    if (!h || !Q_d || !S_d || !rec_d || n_gps < 1 || n_gps > 64 || M < 0) return SO_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    static_assert(sizeof(SafePartial) <= 64 && sizeof(MaxPartial) <= 64, "partials must fit the workspace stride");
    k_reduce_safe<<<grid_for(h, M), kThreads, 0, (cudaStream_t)stream>>>(Q_d, 2 * n_gps, M, row0, S_d, (SafePartial*)h->ws_partials,
                                                                         h->ws_counter, rec_d);
    SO_CHECK_LAUNCH(h, "k_reduce_safe");
    return SO_OK;
}

extern "C" int so_sets_maximizers(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                                  double max_l0, const double* scaling_h, uint8_t* Mmask_d, so_max_record* rec_d, void* stream) {
    if (!h || !Q_d || !S_d || !scaling_h || !Mmask_d || !rec_d || n_gps < 1 || n_gps > 64 || M < 0) return SO_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    Scal sc;
    for (int i = 0; i < 64; ++i) sc.v[i] = i < n_gps ? scaling_h[i] : 1.0;
    k_maximizers<<<grid_for(h, M), kThreads, 0, (cudaStream_t)stream>>>(Q_d, n_gps, M, row0, S_d, max_l0, sc, Mmask_d,
                                                                        (MaxPartial*)h->ws_partials, h->ws_counter, rec_d);
    SO_CHECK_LAUNCH(h, "k_maximizers");
    return SO_OK;
}

extern "C" int so_sets_candidates(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                                  const uint8_t* Mmask_d, double max_var, const double* scaling_h, const double* thr_h,
                                  uint8_t* cand_mask_d, double* cand_key_d, int64_t* cand_row_d, int64_t cap,
                                  int64_t* n_cand_d, void* stream) {
    if (!h || !Q_d || !S_d || !Mmask_d || !scaling_h || !thr_h || !n_cand_d || n_gps < 1 || n_gps > 64 || M < 0 || cap < 0)
        return SO_ERR_BAD_ARG;
    if (cap > 0 && (!cand_key_d || !cand_row_d)) return SO_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    Scal sc, th;
    for (int i = 0; i < 64; ++i) { sc.v[i] = i < n_gps ? scaling_h[i] : 1.0; th.v[i] = i < n_gps ? thr_h[i] : 0.0; }
    SO_CUDA(h, cudaMemsetAsync(n_cand_d, 0, sizeof(int64_t), (cudaStream_t)stream));
    k_candidates<<<grid_for(h, M), kThreads, 0, (cudaStream_t)stream>>>(Q_d, n_gps, M, row0, S_d, Mmask_d, max_var, sc, th, cand_mask_d,
                                                                        cand_key_d, cand_row_d, cap,
                                                                        reinterpret_cast<unsigned long long*>(n_cand_d));
    SO_CHECK_LAUNCH(h, "k_candidates");
    return SO_OK;
}
