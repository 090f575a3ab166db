// K4 -- batched expander test.  The reference decides whether a candidate x_c is an expander by
// refitting the GP with the fake observation (x_c, u_c), predicting every unsafe row and refitting
// again (safeopt/gp_opt.py:579-606; two O(N^3) refits + one O(M N^2) predict per candidate).  Adding
// one observation is a rank-1 update of the posterior:
//     c(x)  = k(x, x_c) - k_x^T Ky^-1 k_c              (posterior covariance with the candidate)
//     s     = var(x_c) + noise + 1e-8
//     mean2 = mean(x) + c(x) (u_c - mean(x_c)) / s ,  var2 = max(var(x) - c(x)^2 / s, 1e-15)
// so B candidates cost ONE sweep over the rows: a (T x N).(N x B) contraction on the fp64 tensor
// pipe with z_b = Ky^-1 k_cb as the stationary operand, reusing the posterior kernel's row
// generators.  Tiles whose rows are all safe are skipped.
#include "posterior_core.cuh"

namespace {

constexpr int kMaxBatch = 32;          // candidates per launch (4 accumulator blocks of 8)
constexpr int kBB = kMaxBatch / 8;

struct CandInfo {                       // per candidate, device memory
    double coef;                        // (u_c - mean_c) / s
    double inv_s;                       // 1 / s
};

struct ExpParams {
    PostParams p;
    const double2* Zfrag;               // kBB x NB blocks of 32 double2, DMMA fragment order
    const CandInfo* cinfo;              // kMaxBatch
    const double* xcs;                  // kMaxBatch x d, candidate coordinates scaled by 1/lengthscale
    const double* axis;                 // grid axis values (grid path)
    int B;
    uint8_t* flags;
};

// z_b = Ky^-1 k_cb = Linv^T (Linv k_cb), written in fragment order; one CTA per candidate.
__global__ void __launch_bounds__(256) k_expander_prep(const double* __restrict__ Linv, const double* __restrict__ Xs,
                                                       int N, int NB, int ld, int d, int kind, double variance, double noise_j,
                                                       const double* __restrict__ inv_ls_d, const double* __restrict__ xc,
                                                       const double* __restrict__ mean_c, const double* __restrict__ var_c,
                                                       const double* __restrict__ u_c, double* __restrict__ Zfrag,
                                                       CandInfo* __restrict__ cinfo, double* __restrict__ xcs) {
    extern __shared__ double sm[];
    const int Npad = 8 * NB;
    double* kc = sm;            // Npad
    double* v = sm + Npad;      // Npad
    const int b = blockIdx.x;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double r2 = 0.0;
        for (int j = 0; j < d; ++j) {
            const double t = xc[(size_t)b * d + j] * inv_ls_d[j] - Xs[(size_t)n * d + j];
            r2 = fma(t, t, r2);
        }
        double k;
        switch (kind) {
            case SO_KERNEL_RBF: k = kernel_of_r2<SO_KERNEL_RBF>(r2, variance); break;
            case SO_KERNEL_MATERN32: k = kernel_of_r2<SO_KERNEL_MATERN32>(r2, variance); break;
            default: k = kernel_of_r2<SO_KERNEL_MATERN52>(r2, variance); break;
        }
        kc[n] = k;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double acc = 0.0;
        const double* row = Linv + (size_t)n * ld;
        for (int m = 0; m <= n; ++m) acc = fma(row[m], kc[m], acc);
        v[n] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        double acc = 0.0;
        for (int i = c; i < N; ++i) acc = fma(Linv[(size_t)i * ld + c], v[i], acc);
        const int kb = c >> 3, q = (c & 7) >> 1, half = c & 1;
        const int lane = (b & 7) * 4 + q, bb = b >> 3;
        Zfrag[(((size_t)bb * NB + kb) * 32 + lane) * 2 + half] = acc;
    }
    if (threadIdx.x == 0) {
        const double s = var_c[b] + noise_j;
        CandInfo ci;
        ci.inv_s = 1.0 / s;
        ci.coef = (u_c[b] - mean_c[b]) / s;
        cinfo[b] = ci;
    }
    for (int j = threadIdx.x; j < d; j += blockDim.x) xcs[(size_t)b * d + j] = xc[(size_t)b * d + j] * inv_ls_d[j];
}

template <int KIND, bool GRID>
__global__ void __launch_bounds__(kThreads, 1) k_expander(const __grid_constant__ ExpParams ep) {
    const PostParams& p = ep.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SmemLayout L = smem_layout(p.NB, p.T, p.d, p.RG, GRID);
    double2* sK = reinterpret_cast<double2*>(smem_raw);
    double* sXs = reinterpret_cast<double*>(smem_raw + L.xs_off);
    double* sXt = reinterpret_cast<double*>(smem_raw + L.xt_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Npad = 8 * p.NB, T = p.T, TB = p.TB, NB = p.NB, d = p.d;

    double* sExpT = reinterpret_cast<double*>(smem_raw + L.exp_off);
    load_exp_table(sExpT);
    if (!GRID)
        for (int i = threadIdx.x; i < Npad * d; i += kThreads) sXs[i] = p.Xs[i];
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int64_t tile_local0 = tile * T;
        // skip tiles without unsafe rows (block-wide OR doubles as the barrier after the row load)
        int any = 0;
        for (int t = threadIdx.x; t < T; t += kThreads) {
            const int64_t row = tile_local0 + t;
            if (row < p.M && !p.S[row]) any = 1;
        }
        if (!GRID) load_tile_rows(p, sXt, tile_local0);
        if (!__syncthreads_or(any)) continue;   // uniform across the CTA
        if (GRID) gen_grid(p, sK, p.row0 + tile_local0, warp, lane);
        else gen_rows<KIND>(p, sK, sXs, sXt, sExpT, warp, lane);
        __syncthreads();

        for (int ct = warp; ct < TB; ct += kWarps) {
            // out[t][b] = sum_n Kx[n][t] z_b[n]: the Kx fragment is the A operand (rows = t), Z the B operand
            double acc[kBB][2];
#pragma unroll
            for (int bb = 0; bb < kBB; ++bb) { acc[bb][0] = 0.0; acc[bb][1] = 0.0; }
            const double2* kx = sK + (size_t)ct * 32 + lane;
            const double2* zf = ep.Zfrag + lane;
            for (int kb = 0; kb < NB; ++kb) {
                const double2 a = kx[(size_t)kb * TB * 32];
#pragma unroll
                for (int bb = 0; bb < kBB; ++bb) {
                    const double2 z = __ldg(zf + ((size_t)bb * NB + kb) * 32);
                    dmma884(acc[bb][0], acc[bb][1], a.x, z.x);
                    dmma884(acc[bb][0], acc[bb][1], a.y, z.y);
                }
            }
            // lane holds rows t = ct*8 + lane/4 and candidates b = 8*bb + 2*(lane%4) + {0,1}
            const int t = ct * 8 + (lane >> 2);
            const int64_t row = tile_local0 + t;
            const bool live = row < p.M && !p.S[row];
            double mu = 0.0, var = 0.0;
            double xt[SO_MAX_DIM];
            if (live) {
                mu = p.mean[row];
                var = p.var[row];
            }
#pragma unroll
            for (int j = 0; j < SO_MAX_DIM; ++j) {
                if (j < d) {
                    if (GRID) {
                        const int64_t grow = p.row0 + (row < p.M ? row : p.M - 1);
                        const int idx = (int)((grow / p.gstride[j < kGridMaxDim ? j : 0]) % p.gn[j < kGridMaxDim ? j : 0]);
                        xt[j] = ep.axis[p.goff[j < kGridMaxDim ? j : 0] + idx] * p.inv_ls[j];
                    } else {
                        xt[j] = sXt[t * d + j];
                    }
                } else {
                    xt[j] = 0.0;
                }
            }
            unsigned hit = 0;   // bit (2*bb + h) set if candidate 8*bb + 2*(lane%4) + h lifts this row
#pragma unroll
            for (int bb = 0; bb < kBB; ++bb) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int b = 8 * bb + 2 * (lane & 3) + hh;
                    if (live && b < ep.B) {
                        double r2 = 0.0;
#pragma unroll
                        for (int j = 0; j < SO_MAX_DIM; ++j) {
                            if (j < d) {
                                const double df = xt[j] - ep.xcs[(size_t)b * d + j];
                                r2 = fma(df, df, r2);
                            }
                        }
                        const double kxc = kernel_of_r2_fast<KIND>(r2, p.variance, sExpT);
                        const double c = kxc - acc[bb][hh];
                        const CandInfo ci = ep.cinfo[b];
                        const double mean2 = fma(c, ci.coef, mu);
                        double var2 = var - c * c * ci.inv_s;
                        var2 = var2 > SO_VAR_FLOOR ? var2 : SO_VAR_FLOOR;
                        const double l2 = mean2 - p.beta * sqrt(var2);
                        if (l2 >= p.fmin) hit |= 1u << (2 * bb + hh);
                    }
                }
            }
            // OR over the 8 rows of the tile (lane bits 2..4)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) hit |= __shfl_xor_sync(0xffffffffu, hit, o);
            if (lane < 4 && hit) {
#pragma unroll
                for (int bb = 0; bb < kBB; ++bb)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
                        if (hit & (1u << (2 * bb + hh))) ep.flags[8 * bb + 2 * lane + hh] = 1;   // benign race: all writers store 1
            }
        }
        __syncthreads();
    }
}

template <int KIND, bool GRID>
int launch_expander(so_handle* h, const ExpParams& ep, size_t smem, cudaStream_t stream) {
    static int configured_for = -1;
    if (configured_for != h->device) {
        SO_CUDA(h, cudaFuncSetAttribute(k_expander<KIND, GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        configured_for = h->device;
    }
    const int grid = (int)(ep.p.ntiles < (int64_t)h->num_sms ? ep.p.ntiles : (int64_t)h->num_sms);
    k_expander<KIND, GRID><<<grid, kThreads, smem, stream>>>(ep);
    SO_CHECK_LAUNCH(h, "k_expander");
    return SO_OK;
}

}  // namespace

extern "C" int so_expander_check(so_handle* h, int gp, const double* Xstar_d, int64_t row0, int64_t M, const uint8_t* S_d,
                                 const double* mean_d, const double* var_d, const double* xc_d, const double* mean_c_d,
                                 const double* var_c_d, const double* u_c_d, int B, double beta, double fmin,
                                 uint8_t* flags_d, void* stream_) {
    if (!h) return SO_ERR_BAD_ARG;
    if (gp < 0 || gp >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "expander: gp index out of range");
    GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "expander: GP not fitted");
    if (!S_d || !mean_d || !var_d || !xc_d || !mean_c_d || !var_c_d || !u_c_d || !flags_d || M < 0)
        return so_fail(h, SO_ERR_BAD_ARG, "expander: null argument");
    if (B < 1 || B > kMaxBatch) return so_fail(h, SO_ERR_BAD_ARG, "expander: batch must be in [1, 32]");
    const bool grid = Xstar_d == nullptr;
    if (grid) {
        if (!h->grid.defined || !g.grid_ready) return so_fail(h, SO_ERR_BAD_ARG, "expander: grid path needs so_grid_define + so_grid_prepare");
        if (g.kind != SO_KERNEL_RBF) return so_fail(h, SO_ERR_UNSUPPORTED, "expander: grid path needs an RBF kernel");
        if (row0 < 0 || row0 + M > h->grid.rows) return so_fail(h, SO_ERR_BAD_ARG, "expander: rows outside the grid");
    }
    if (M == 0) return SO_OK;
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int NB = g.NB, Npad = 8 * NB, d = g.d;

    // workspace: Zfrag | cinfo | xcs | inv_ls
    const size_t z_doubles = (size_t)kBB * NB * 64;
    const size_t need = z_doubles + 2 * kMaxBatch + (size_t)kMaxBatch * SO_MAX_DIM + SO_MAX_DIM;
    if (need > h->ws_z_cap) {
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (h->ws_z) cudaFree(h->ws_z);
        h->ws_z = nullptr;
        const size_t cap = need * 2;
        SO_CUDA(h, cudaMalloc(&h->ws_z, sizeof(double) * cap));
        h->ws_z_cap = cap;
    }
    double* Zfrag = h->ws_z;
    CandInfo* cinfo = reinterpret_cast<CandInfo*>(h->ws_z + z_doubles);
    double* xcs = h->ws_z + z_doubles + 2 * kMaxBatch;
    double* inv_ls_d = xcs + (size_t)kMaxBatch * SO_MAX_DIM;
    SO_CUDA(h, cudaMemsetAsync(Zfrag, 0, sizeof(double) * z_doubles, stream));
    SO_CUDA(h, cudaMemcpyAsync(inv_ls_d, g.inv_ls, sizeof(double) * SO_MAX_DIM, cudaMemcpyHostToDevice, stream));
    k_expander_prep<<<B, 256, sizeof(double) * 2 * Npad, stream>>>(g.Linv, g.Xs, g.N, NB, g.ld, d, g.kind, g.variance,
                                                                   g.noise + SO_JITTER, inv_ls_d, xc_d, mean_c_d, var_c_d,
                                                                   u_c_d, Zfrag, cinfo, xcs);
    SO_CHECK_LAUNCH(h, "k_expander_prep");

    // tile: 8 column tiles of 8 rows per CTA when shared memory allows, fewer for large N
    ExpParams ep;
    PostParams& p = ep.p;
    int T = 64;
    while (T >= 16 && smem_layout(NB, T, d, 1, grid).total > (size_t)h->smem_optin) T >>= 1;
    if (T < 16) return so_fail(h, SO_ERR_CAPACITY, "expander: N too large for the shared-memory tile");
    p.N = g.N; p.NB = NB; p.d = d; p.RG = 1; p.CG = 8; p.T = T; p.TB = T / 8; p.npass = 1; p.kind = g.kind;
    p.Afrag = g.Afrag; p.zvec = g.zvec; p.Xs = g.Xs;
    for (int j = 0; j < SO_MAX_DIM; ++j) p.inv_ls[j] = g.inv_ls[j];
    p.variance = g.variance;
    p.Xstar = Xstar_d; p.M = M; p.row0 = row0; p.ntiles = (M + T - 1) / T;
    p.gd = 0; p.E = g.E;
    for (int j = 0; j < kGridMaxDim; ++j) { p.gn[j] = 1; p.goff[j] = 0; p.gstride[j] = 1; }
    if (grid) {
        p.gd = h->grid.d;
        for (int j = 0; j < p.gd; ++j) { p.gn[j] = h->grid.n[j]; p.goff[j] = h->grid.off[j]; p.gstride[j] = h->grid.stride[j]; }
    }
    p.beta = beta; p.fmin = fmin;
    p.mean = const_cast<double*>(mean_d); p.var = const_cast<double*>(var_d); p.Q = nullptr; p.q_stride = 0; p.q_col = 0;
    p.S = const_cast<uint8_t*>(S_d); p.safe_mode = SO_SAFE_NONE; p.n_out = 1; p.use_row_table = 0;
    ep.Zfrag = reinterpret_cast<const double2*>(Zfrag);
    ep.cinfo = cinfo; ep.xcs = xcs; ep.axis = h->grid.axis; ep.B = B; ep.flags = flags_d;
    const size_t smem = smem_layout(NB, T, d, 1, grid).total;
    if (grid) return launch_expander<SO_KERNEL_RBF, true>(h, ep, smem, stream);
    switch (g.kind) {
        case SO_KERNEL_RBF: return launch_expander<SO_KERNEL_RBF, false>(h, ep, smem, stream);
        case SO_KERNEL_MATERN32: return launch_expander<SO_KERNEL_MATERN32, false>(h, ep, smem, stream);
        default: return launch_expander<SO_KERNEL_MATERN52, false>(h, ep, smem, stream);
    }
}
