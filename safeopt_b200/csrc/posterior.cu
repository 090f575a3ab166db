// K2 -- fused GP posterior over candidate rows: kernel-row build, the dense triangular contraction V = L^-1 k on the
// fp64 tensor pipe (DMMA.8x8x4), predictive mean/variance, confidence bounds and the safe bit.  Stands in for
// `gp.predict_noiseless(self.inputs)` (safeopt/gp_opt.py:469) + :471-476 + this GP's factor of :481.  Nothing of size
// N x M ever reaches HBM (the reference materialises three such temporaries).
//
// Two persistent kernels, one 256-thread CTA per SM, looping over tiles of T candidate rows:
//   k_posterior      (explicit rows, any stationary kernel; also the fallback of the grid path)
//       gen : k(x*, X) for the tile, written to shared memory directly in DMMA B-fragment order (distance + exp or
//             Matern profile per value; on a grid: product of per-axis table entries)
//       mma : contract_tile (posterior_core.cuh)   epi : finalize_row
//   k_posterior_tma  (product grids with an RBF kernel -- the headline path, posterior_tma.cuh)
//       no generation phase at all: scaled operands A'(s) + TMA double buffer of the fragment-ordered fast table.
// Algorithmic work per row and GP: N^2/2 FMA (contraction) + N kernel evaluations; algorithmic HBM bytes: d*8 in (0 on
// the grid path) + 32 out (mean, var, l, u) + 1 (S).  See DESIGN.md section 3.
#include "posterior_tma.cuh"
#include "posterior_ring.cuh"
#include "posterior_f32.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace {

template <int BT, int KIND, bool GRID>
__global__ void __launch_bounds__(kThreads, 1) k_posterior(const __grid_constant__ PostParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n_extra = p.n_out - 1;
    const SmemLayout L = smem_layout(p.NB, p.T, p.d, p.RG, GRID, n_extra);
    double2* sK = reinterpret_cast<double2*>(smem_raw);
    double* sMeanX = reinterpret_cast<double*>(smem_raw + L.meanx_off);
    double* sXs = reinterpret_cast<double*>(smem_raw + L.xs_off);
    double* sXt = reinterpret_cast<double*>(smem_raw + L.xt_off);
    double* sSS = reinterpret_cast<double*>(smem_raw + L.ss_off);
    double* sMean = reinterpret_cast<double*>(smem_raw + L.mean_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Npad = 8 * p.NB, T = p.T, RG = p.RG;
    const int g = warp % RG, cg = warp / RG;
    double* sExpT = reinterpret_cast<double*>(smem_raw + L.exp_off);

    if (!GRID) {
        load_exp_table(sExpT);
        for (int i = threadIdx.x; i < Npad * p.d; i += kThreads) sXs[i] = p.Xs[i];
        if ((int64_t)blockIdx.x < p.ntiles) load_tile_rows(p, sXt, (int64_t)blockIdx.x * T);
    }
    __syncthreads();

    int par = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, par ^= 1) {
        const int64_t tile_local0 = tile * T;
        double* sSST = sSS + (size_t)par * RG * T;
        double* sMeanT = sMean + (size_t)par * RG * T;
        double* sMeanXT = sMeanX + (size_t)par * n_extra * RG * T;
#ifndef SO_K2_SKIP_GEN     // defined for timing experiments only (tools/build_variant.sh): contracts whatever shared memory holds
        if (GRID) gen_grid(p, sK, p.row0 + tile_local0, warp, lane);
        else gen_rows<KIND>(p, sK, sXs, sXt + (size_t)par * T * p.d, sExpT, warp, lane);
#endif
        __syncthreads();
        if (!GRID) {
            // the next tile's candidate rows are fetched under the contraction
            const int64_t next = tile + gridDim.x;
            if (next < p.ntiles) load_tile_rows(p, sXt + (size_t)(par ^ 1) * T * p.d, next * T);
        }
        PlainB bsrc{sK + (size_t)(cg * BT) * 32 + lane, p.TB};
        contract_tile<BT>(p, p.Afrag + lane, bsrc, sSST, sMeanT, sMeanXT, g, cg, lane);
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += kThreads) {
            const int64_t row = tile_local0 + t;
            if (row < p.M) finalize_row(p, sSST, sMeanT, sMeanXT, t, row);
        }
        // the partial-sum buffers alternate with `par`; sK is rewritten only after the barrier at the top of the next
        // iteration's contraction, which every thread reaches after its epilogue
    }
}

// ---------------------------------------------------------------- cross-check kernel (tests only)
__global__ void __launch_bounds__(256) k_posterior_simple(const double* __restrict__ Linv, const double* __restrict__ alpha,
                                                          const double* __restrict__ Xs, const double* __restrict__ Xstar,
                                                          int N, int Npad, int d, int kind, double variance,
                                                          const double* __restrict__ inv_ls_d, int64_t M,
                                                          double* __restrict__ mean, double* __restrict__ var) {
    extern __shared__ double sk[];
    __shared__ double red[2][8];
    const int64_t row = blockIdx.x;
    if (row >= M) return;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double r2 = 0.0;
        for (int j = 0; j < d; ++j) {
            double t = Xstar[(size_t)row * d + j] * inv_ls_d[j] - Xs[(size_t)n * d + j];
            r2 = fma(t, t, r2);
        }
        double k;
        switch (kind) {
            case SO_KERNEL_RBF: k = kernel_of_r2<SO_KERNEL_RBF>(r2, variance); break;
            case SO_KERNEL_MATERN32: k = kernel_of_r2<SO_KERNEL_MATERN32>(r2, variance); break;
            default: k = kernel_of_r2<SO_KERNEL_MATERN52>(r2, variance); break;
        }
        sk[n] = k;
    }
    __syncthreads();
    double ss = 0.0, mu = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double v = 0.0;
        for (int m = 0; m <= n; ++m) v = fma(Linv[(size_t)n * Npad + m], sk[m], v);
        ss = fma(v, v, ss);
        mu = fma(sk[n], alpha[n], mu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        mu += __shfl_xor_sync(0xffffffffu, mu, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = mu; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += red[0][w]; m += red[1][w]; }
        double v = variance - s;
        var[row] = v > SO_VAR_FLOOR ? v : SO_VAR_FLOOR;
        mean[row] = m;
    }
}

// ---------------------------------------------------------------- grid tables / rows
__global__ void k_grid_tables(const double* __restrict__ axis, const double* __restrict__ Xs, double* __restrict__ E,
                              int total, int N, int Npad, int d, const int* __restrict__ axis_of, double variance,
                              const double* __restrict__ inv_ls_d) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (n >= Npad || i >= total) return;
    const int j = axis_of[i];
    double v = 0.0;
    if (n < N) {
        const double t = axis[i] * inv_ls_d[j] - Xs[(size_t)n * d + j];
        v = exp(-0.5 * (t * t));
        if (j == 0) v *= variance;
    }
    E[(size_t)i * Npad + n] = v;
}

struct GridDecode {
    int d;
    int n[kGridMaxDim];
    int off[kGridMaxDim];
    int64_t stride[kGridMaxDim];
};

__global__ void k_grid_rows(GridDecode gd, const double* __restrict__ axis, int64_t row0, int64_t M, double* __restrict__ X) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const int64_t row = row0 + r;
    for (int j = 0; j < gd.d; ++j) {
        const int idx = (int)((row / gd.stride[j]) % gd.n[j]);
        X[(size_t)r * gd.d + j] = axis[gd.off[j] + idx];
    }
}

// ---------------------------------------------------------------- host
struct LaunchPlan { int BT, RG, CG, T, TB, npass; size_t smem; };

void row_groups(int NB, int& RG, int& CG) {
    if (NB >= 32) { RG = 8; CG = 1; }
    else if (NB >= 16) { RG = 4; CG = 2; }
    else if (NB >= 8) { RG = 2; CG = 4; }
    else { RG = 1; CG = 8; }
}

int plan_launch(so_handle* h, const GPState& g, int64_t M, bool grid, int n_extra, LaunchPlan& lp) {
    const int NB = g.NB;
    row_groups(NB, lp.RG, lp.CG);
    const size_t limit = (size_t)h->smem_optin;
    int BT = 8;
    while (BT >= 2 && smem_layout(NB, 8 * BT * lp.CG, g.d, lp.RG, grid, n_extra).total > limit) BT >>= 1;
    if (BT < 2) return so_fail(h, SO_ERR_CAPACITY, "posterior: N too large for the shared-memory tile (N <= ~1600)");
    // few tiles -> smaller tiles so that more SMs get work
    while (BT > 2 && (M + 8 * BT * lp.CG - 1) / (8 * BT * lp.CG) < 2 * (int64_t)h->num_sms) BT >>= 1;
    lp.BT = BT;
    lp.T = 8 * BT * lp.CG;
    lp.TB = BT * lp.CG;
    lp.npass = (NB + 4 * lp.RG - 1) / (4 * lp.RG);
    lp.smem = smem_layout(NB, lp.T, g.d, lp.RG, grid, n_extra).total;
    return SO_OK;
}

// Block rows of the eight warps for an arbitrary NB (see PostParams::row_table): longest row first to the least loaded warp,
// every warp at most 4 * npass rows; each warp's rows ascending, four per pass.
int plan_rows(int NB, short (&table)[kMaxPass][8][kMaxSlots], int slots = 4) {
    const int per_warp = (NB + 7) / 8;
    const int npass = (per_warp + slots - 1) / slots;
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0}, count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int mine[8][kMaxSlots * kMaxPass];
    for (int i = NB - 1; i >= 0; --i) {
        int best = -1;
        for (int w = 0; w < 8; ++w)
            if (count[w] < slots * npass && (best < 0 || load[w] < load[best])) best = w;
        mine[best][count[best]++] = i;
        load[best] += i + 1;
    }
    for (int p = 0; p < kMaxPass; ++p)
        for (int w = 0; w < 8; ++w)
            for (int q = 0; q < kMaxSlots; ++q) table[p][w][q] = -1;
    for (int w = 0; w < 8; ++w) {
        std::sort(mine[w], mine[w] + count[w]);
        for (int q = 0; q < count[w]; ++q) table[q / slots][w][q % slots] = (short)mine[w][q];
    }
    return npass;
}

// Outputs 1..n-1 of a launch that evaluates several GPs sharing one factorisation (so_posterior_*_multi).
struct ExtraOut {
    int n = 0;
    int gp[kMaxOut - 1];
    double fmin[kMaxOut - 1];
    double* mean[kMaxOut - 1];
    double* var[kMaxOut - 1];
    int q_col[kMaxOut - 1];
};

template <int BT, int KIND, bool GRID>
int launch_one(so_handle* h, const PostParams& p, const LaunchPlan& lp, cudaStream_t stream) {
    static int configured_for = -1;
    if (configured_for != h->device) {
        SO_CUDA(h, cudaFuncSetAttribute(k_posterior<BT, KIND, GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        configured_for = h->device;
    }
    const int grid = (int)(p.ntiles < (int64_t)h->num_sms ? p.ntiles : (int64_t)h->num_sms);
    k_posterior<BT, KIND, GRID><<<grid, kThreads, lp.smem, stream>>>(p);
    SO_CHECK_LAUNCH(h, "k_posterior");
    return SO_OK;
}

template <int KIND, bool GRID>
int launch_bt(so_handle* h, const PostParams& p, const LaunchPlan& lp, cudaStream_t stream) {
    switch (lp.BT) {
        case 8: return launch_one<8, KIND, GRID>(h, p, lp, stream);
        case 4: return launch_one<4, KIND, GRID>(h, p, lp, stream);
        default: return launch_one<2, KIND, GRID>(h, p, lp, stream);
    }
}

// ---- TMA double-buffer kernel (grid path default) ----------------------------------------------------------------
struct TmaPlan { int BT, RG, CG, T, TB, npass, kb_pad, warps; size_t smem; bool ring; bool ns2x; bool ns6; };

// SO_K2_RING=0 keeps the resident double buffer (smaller tiles) for every N, SO_K2_RING=all streams as soon as the 48-row
// tile does not fit twice (A/B measurements); default: stream when the 32-row tile does not fit twice either.
bool ring_allowed() {
    const char* v = std::getenv("SO_K2_RING");
    return !(v && std::string(v) == "0");
}
int ring_after_option() {
    const char* v = std::getenv("SO_K2_RING");
    return v && std::string(v) == "all" ? 0 : 1;
}

// 8 warps per CTA (four block rows per warp) by default; SO_K2_WARPS=16 selects the 16-warp variant (two block rows per
// warp, four warps per scheduler) for A/B measurements -- measured 13.55 vs 13.40 ms at config 4, pipe 84.7 % vs 85.5 %
// busy (profiles/r01_k2_variants.md): the extra warps do not buy back what the doubled B-fragment traffic costs.
int tma_warps() {
    const char* v = std::getenv("SO_K2_WARPS");
    return v && std::string(v) == "16" ? 16 : 8;
}

int plan_tma(so_handle* h, const GPState& g, TmaPlan& tp) {
    const int NB = g.NB;
    tp.kb_pad = NB;       // no padding to whole bulk-copy chunks: at NB = 33..35 that padding alone pushed the tile from 48 to 32 rows
    const int options[3] = {6, 4, 2};
    tp.ring = false;
    tp.ns2x = false;
    tp.ns6 = false;
    // SO_K2_SMALLN=1, N <= 128: eight warps with two block rows each, two CTAs per SM (see k_posterior_tma).  Measured on config 4's
    // grid (profiles/r02_k2_small_n.md): N = 32 +8 %, 64 -1 %, 96 +4 %, 128 -7 % against the four-rows-per-warp kernel -- the doubled
    // B-fragment traffic costs what the second CTA hides -- so it is an A/B switch, off by default.
    {
        const char* v = std::getenv("SO_K2_SMALLN");
        if (NB <= 16 && tma_warps() == 8 && v && std::string(v) == "1") {
            int rg = 1;
            while (rg < 8 && 2 * rg < NB) rg *= 2;
            const int cg = 8 / rg;
            for (int k = 0; k < 3; ++k) {
                const int bt = options[k];
                const TmaSmem L = tma_smem(tp.kb_pad, bt * cg, rg, 8 * bt * cg, kMaxOut - 1);
                if (2 * (L.total + 1024) <= (size_t)h->smem_optin) {
                    tp.warps = 8; tp.RG = rg; tp.CG = cg; tp.npass = (NB + 2 * rg - 1) / (2 * rg);
                    tp.BT = bt; tp.TB = bt * cg; tp.T = 8 * bt * cg; tp.smem = L.total; tp.ns2x = true;
                    return SO_OK;
                }
            }
        }
    }
    for (int warps = tma_warps(); warps >= 8; warps -= 8) {
        const int ns = warps == 16 ? 2 : 4;
        int rg = 1;
        while (rg < warps && ns * rg < NB) rg *= 2;       // row groups: enough warps to cover the block rows in one pass
        tp.warps = warps; tp.RG = rg; tp.CG = warps / rg;
        tp.npass = (NB + ns * rg - 1) / (ns * rg);
        for (int k = 0; k < 3; ++k) {
            const int bt = options[k];
            const TmaSmem L = tma_smem(tp.kb_pad, bt * tp.CG, tp.RG, 8 * bt * tp.CG);
            if (L.total <= (size_t)h->smem_optin) {
                tp.BT = bt; tp.TB = bt * tp.CG; tp.T = 8 * bt * tp.CG; tp.smem = L.total;
                // 32-row tile with 33..48 block rows: six block rows per warp, one pass (SO_K2_NS6=0: four rows, two passes)
                const char* v6 = std::getenv("SO_K2_NS6");
                if (bt == 4 && warps == 8 && rg == 8 && NB > 32 && NB <= 48 && !(v6 && std::string(v6) == "0")) {
                    tp.ns6 = true;
                    tp.npass = 1;
                }
                return SO_OK;
            }
            // neither a 48- nor a 32-row tile fits twice (N > 416): stream B through the k-chunk ring and keep the tile 48 rows
            // wide.  Measured at config 4's grid: N = 512 64.8 ms (ring) vs 69.8 (16-row double buffer); for N = 281..416 the
            // 32-row double buffer is faster than the ring (21.2 vs 28.7 ms at N = 288), so it stays.
            if (k == ring_after_option() && warps == 8 && rg == 8 && ring_allowed() &&
                ring_smem(6, 8, 48, kMaxOut - 1, (size_t)h->smem_optin).stages >= 4) {
                tp.BT = 6; tp.TB = 6; tp.T = 48; tp.smem = 0; tp.ring = true;
                return SO_OK;
            }
        }
    }
    return SO_ERR_CAPACITY;
}

template <int BT, int WARPS, int NSV>
int launch_tma_one(so_handle* h, const TmaParams& tp, size_t smem, cudaStream_t stream) {
    static int configured_for = -1;
    if (configured_for != h->device) {
        SO_CUDA(h, cudaFuncSetAttribute(k_posterior_tma<BT, WARPS, NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        configured_for = h->device;
    }
    const int64_t slots = (int64_t)h->num_sms * (NSV == 2 ? 2 : 1);
    const int grid = (int)(tp.p.ntiles < slots ? tp.p.ntiles : slots);
    k_posterior_tma<BT, WARPS, NSV><<<grid, WARPS * 32, smem, stream>>>(tp);
    SO_CHECK_LAUNCH(h, "k_posterior_tma");
    return SO_OK;
}

template <int WARPS, int NSV = 0>
int launch_tma(so_handle* h, int bt, const TmaParams& tp, size_t smem, cudaStream_t stream) {
    switch (bt) {
        case 6: return launch_tma_one<6, WARPS, NSV>(h, tp, smem, stream);
        case 4: return launch_tma_one<4, WARPS, NSV>(h, tp, smem, stream);
        default: return launch_tma_one<2, WARPS, NSV>(h, tp, smem, stream);
    }
}

// SO_K2_VARIANT=bulk forces the generate-then-contract kernel on the grid path too (A/B measurements).
bool force_bulk() {
    const char* v = std::getenv("SO_K2_VARIANT");
    return v && std::string(v) == "bulk";
}

int run_posterior(so_handle* h, int gp, const double* Xstar_d, bool grid, int64_t row0, int64_t M, double beta, double fmin,
                  double* mean_d, double* var_d, double* Q_d, int q_stride, int q_col, uint8_t* S_d, int safe_mode,
                  void* stream_, const ExtraOut* extra = nullptr) {
    if (!h) return SO_ERR_BAD_ARG;
    if (gp < 0 || gp >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "posterior: gp index out of range");
    GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "posterior: GP not fitted");
    if (M < 0) return so_fail(h, SO_ERR_BAD_ARG, "posterior: M < 0");
    if (M == 0) return SO_OK;
    if (Q_d && (q_stride < 2 || q_col < 0 || q_col + 2 > q_stride))
        return so_fail(h, SO_ERR_BAD_ARG, "posterior: Q column out of range");
    if (safe_mode < SO_SAFE_NONE || safe_mode > SO_SAFE_AND) return so_fail(h, SO_ERR_BAD_ARG, "posterior: bad safe_mode");
    if (grid) {
        if (!h->grid.defined) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid: no grid defined");
        if (!g.grid_ready) return so_fail(h, SO_ERR_NOT_FITTED, "posterior_grid: so_grid_prepare not called after so_fit");
        if (g.kind != SO_KERNEL_RBF) return so_fail(h, SO_ERR_UNSUPPORTED, "posterior_grid: separable tables need an RBF kernel");
        if (row0 < 0 || row0 + M > h->grid.rows) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid: rows outside the grid");
    } else if (!Xstar_d) {
        return so_fail(h, SO_ERR_BAD_ARG, "posterior_rows: null candidate pointer");
    }
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool tma = grid && g.tma_ready && !force_bulk();

    TmaParams tp;
    PostParams& p = tp.p;
    p.N = g.N; p.NB = g.NB; p.d = g.d; p.kind = g.kind;
    p.Afrag = g.Afrag; p.zvec = g.zvec; p.Xs = g.Xs;
    for (int j = 0; j < SO_MAX_DIM; ++j) p.inv_ls[j] = g.inv_ls[j];
    p.variance = g.variance;
    p.Xstar = Xstar_d; p.M = M; p.row0 = row0;
    p.gd = 0; p.E = g.E;
    for (int j = 0; j < kGridMaxDim; ++j) { p.gn[j] = 1; p.goff[j] = 0; p.gstride[j] = 1; }
    if (grid) {
        p.gd = h->grid.d;
        for (int j = 0; j < p.gd; ++j) { p.gn[j] = h->grid.n[j]; p.goff[j] = h->grid.off[j]; p.gstride[j] = h->grid.stride[j]; }
    }
    p.beta = beta; p.fmin = fmin;
    p.mean = mean_d; p.var = var_d; p.Q = Q_d; p.q_stride = q_stride; p.q_col = q_col; p.S = S_d; p.safe_mode = safe_mode;
    p.n_out = 1;
    for (int o = 0; o < kMaxOut - 1; ++o) { p.zvec_x[o] = nullptr; p.fmin_x[o] = 0.0; p.mean_x[o] = nullptr; p.var_x[o] = nullptr; p.q_col_x[o] = 0; }
    if (extra) {
        for (int o = 0; o < extra->n; ++o) {
            if (extra->gp[o] < 0 || extra->gp[o] >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "posterior_multi: gp index out of range");
            const GPState& e = h->gps[extra->gp[o]];
            if (!e.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "posterior_multi: GP not fitted");
            bool same = e.N == g.N && e.d == g.d && e.kind == g.kind && e.variance == g.variance && e.noise == g.noise;
            for (int j = 0; j < g.d && same; ++j) same = e.inv_ls[j] == g.inv_ls[j];
            if (!same) return so_fail(h, SO_ERR_BAD_ARG, "posterior_multi: the GPs do not share size, kernel and noise");
            if (Q_d && (extra->q_col[o] < 0 || extra->q_col[o] + 2 > q_stride))
                return so_fail(h, SO_ERR_BAD_ARG, "posterior_multi: Q column out of range");
            p.zvec_x[o] = e.zvec; p.fmin_x[o] = extra->fmin[o]; p.mean_x[o] = extra->mean[o]; p.var_x[o] = extra->var[o];
            p.q_col_x[o] = extra->q_col[o];
        }
        p.n_out = 1 + extra->n;
    }
    const int n_extra = p.n_out - 1;
    p.use_row_table = 0;
    int table_npass = 0;
    if (g.NB >= 32 && g.NB % 32 != 0 && g.NB <= 8 * 4 * kMaxPass) {       // RG == 8 and the closed-form pairing is unbalanced
        table_npass = plan_rows(g.NB, p.row_table);
        p.use_row_table = 1;
    }

    if (tma) {
        p.RG = g.tma_RG; p.CG = g.tma_CG; p.T = g.tma_T; p.TB = g.tma_BT * g.tma_CG;
        const int ns = g.tma_ns6 ? 6 : ((g.tma_warps == 16 || g.tma_ns2x) ? 2 : 4);
        p.npass = (g.NB + ns * p.RG - 1) / (ns * p.RG);
        if (ns == 6) { p.npass = plan_rows(g.NB, p.row_table, 6); p.use_row_table = 1; }
        else if (ns == 4 && p.RG == 8 && p.use_row_table) p.npass = table_npass; else p.use_row_table = 0;
        tp.PfFrag = g.PfFrag; tp.Aprime = g.Aprime; tp.a_stride = g.a_stride; tp.s0 = g.ap_s0;
        if (row0 / h->grid.fast_rows < g.ap_s0 || (row0 + M - 1) / h->grid.fast_rows >= g.ap_s1)
            return so_fail(h, SO_ERR_NOT_FITTED, "posterior_grid: rows outside the range given to so_grid_prepare_rows");
        tp.fast_rows = h->grid.fast_rows; tp.tpb = g.tma_tpb; tp.kb_pad = g.tma_kb_pad;
        // tiles are aligned to the slow blocks of the product grid: global tile = (row / F) * tpb + (row % F) / T
        const int64_t F = h->grid.fast_rows, last_row = row0 + M - 1;
        const int64_t t0 = (row0 / F) * g.tma_tpb + (row0 % F) / p.T;
        const int64_t t1 = (last_row / F) * g.tma_tpb + (last_row % F) / p.T;
        tp.first_tile = t0;
        p.ntiles = t1 - t0 + 1;
        if (g.tma_ring) {
            const RingSmem rs = ring_smem(p.TB, p.RG, p.T, n_extra, (size_t)h->smem_optin);
            static int ring_configured_for = -1;
            if (ring_configured_for != h->device) {
                SO_CUDA(h, cudaFuncSetAttribute(k_posterior_ring<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
                ring_configured_for = h->device;
            }
            const int nblk = (int)(p.ntiles < (int64_t)h->num_sms ? p.ntiles : (int64_t)h->num_sms);
            k_posterior_ring<6><<<nblk, kThreads, rs.total, stream>>>(tp, rs.stages);
            SO_CHECK_LAUNCH(h, "k_posterior_ring");
            return SO_OK;
        }
        const size_t smem = tma_smem(g.tma_kb_pad, p.TB, p.RG, p.T, n_extra).total;
        if (smem > (size_t)h->smem_optin)
            return so_fail(h, SO_ERR_CAPACITY, "posterior_multi: the tile does not fit with that many outputs; evaluate the GPs one by one");
        if (g.tma_ns6) return launch_tma_one<4, 8, 6>(h, tp, smem, stream);
        if (g.tma_ns2x) {
            if (2 * (smem + 1024) > (size_t)h->smem_optin)
                return so_fail(h, SO_ERR_CAPACITY, "posterior_multi: the tile does not fit with that many outputs; evaluate the GPs one by one");
            return launch_tma<8, 2>(h, g.tma_BT, tp, smem, stream);
        }
        return g.tma_warps == 16 ? launch_tma<16>(h, g.tma_BT, tp, smem, stream) : launch_tma<8>(h, g.tma_BT, tp, smem, stream);
    }
    LaunchPlan lp;
    int rc = plan_launch(h, g, M, grid, n_extra, lp);
    if (rc) return rc;
    p.RG = lp.RG; p.CG = lp.CG; p.T = lp.T; p.TB = lp.TB; p.npass = lp.npass;
    if (p.RG == 8 && p.use_row_table) p.npass = table_npass; else p.use_row_table = 0;
    p.ntiles = (M + p.T - 1) / p.T;
    if (grid) return launch_bt<SO_KERNEL_RBF, true>(h, p, lp, stream);
    switch (g.kind) {
        case SO_KERNEL_RBF: return launch_bt<SO_KERNEL_RBF, false>(h, p, lp, stream);
        case SO_KERNEL_MATERN32: return launch_bt<SO_KERNEL_MATERN32, false>(h, p, lp, stream);
        default: return launch_bt<SO_KERNEL_MATERN52, false>(h, p, lp, stream);
    }
}

}  // namespace

extern "C" int so_posterior_rows(so_handle* h, int gp, const double* Xstar_d, int64_t M, double beta, double fmin,
                                 double* mean_d, double* var_d, double* Q_d, int q_stride, int q_col, uint8_t* S_d,
                                 int safe_mode, void* stream) {
    return run_posterior(h, gp, Xstar_d, false, 0, M, beta, fmin, mean_d, var_d, Q_d, q_stride, q_col, S_d, safe_mode, stream);
}

extern "C" int so_posterior_grid(so_handle* h, int gp, int64_t row0, int64_t M, double beta, double fmin, double* mean_d,
                                 double* var_d, double* Q_d, int q_stride, int q_col, uint8_t* S_d, int safe_mode,
                                 void* stream) {
    return run_posterior(h, gp, nullptr, true, row0, M, beta, fmin, mean_d, var_d, Q_d, q_stride, q_col, S_d, safe_mode, stream);
}

static int run_multi(so_handle* h, int n, const int* gps_h, const double* Xstar_d, bool grid, int64_t row0, int64_t M, double beta,
                     const double* fmin_h, double* const* mean_dh, double* const* var_dh, double* Q_d, int q_stride,
                     const int* q_col_h, uint8_t* S_d, int safe_mode, void* stream) {
    if (!h || !gps_h || !fmin_h || !q_col_h) return SO_ERR_BAD_ARG;
    if (n < 1 || n > kMaxOut) return so_fail(h, SO_ERR_BAD_ARG, "posterior_multi: 1 <= n <= 4");
    ExtraOut ex;
    ex.n = n - 1;
    for (int o = 1; o < n; ++o) {
        ex.gp[o - 1] = gps_h[o]; ex.fmin[o - 1] = fmin_h[o]; ex.q_col[o - 1] = q_col_h[o];
        ex.mean[o - 1] = mean_dh ? mean_dh[o] : nullptr;
        ex.var[o - 1] = var_dh ? var_dh[o] : nullptr;
    }
    return run_posterior(h, gps_h[0], Xstar_d, grid, row0, M, beta, fmin_h[0], mean_dh ? mean_dh[0] : nullptr,
                         var_dh ? var_dh[0] : nullptr, Q_d, q_stride, q_col_h[0], S_d, safe_mode, stream, &ex);
}

extern "C" int so_posterior_rows_multi(so_handle* h, int n, const int* gps_h, const double* Xstar_d, int64_t M, double beta,
                                       const double* fmin_h, double* const* mean_dh, double* const* var_dh, double* Q_d,
                                       int q_stride, const int* q_col_h, uint8_t* S_d, int safe_mode, void* stream) {
    return run_multi(h, n, gps_h, Xstar_d, false, 0, M, beta, fmin_h, mean_dh, var_dh, Q_d, q_stride, q_col_h, S_d, safe_mode, stream);
}

extern "C" int so_posterior_grid_multi(so_handle* h, int n, const int* gps_h, int64_t row0, int64_t M, double beta,
                                       const double* fmin_h, double* const* mean_dh, double* const* var_dh, double* Q_d,
                                       int q_stride, const int* q_col_h, uint8_t* S_d, int safe_mode, void* stream) {
    return run_multi(h, n, gps_h, nullptr, true, row0, M, beta, fmin_h, mean_dh, var_dh, Q_d, q_stride, q_col_h, S_d, safe_mode, stream);
}

// Diagnostic, host only: the tile plans of the posterior kernels for a fit with NB block rows on a device with
// `smem_limit` bytes of opt-in shared memory and `num_sms` SMs, without touching a device.
//   out_h[0..9] = grid (TMA) kernel: status, BT, RG, CG, T, npass, ring (0/1), ring stages, shared-memory bytes, warps
//   out_h[10..17] = explicit-rows kernel (M candidates, dimension d): status, BT, RG, CG, T, npass, shared-memory bytes, 0
extern "C" int so_debug_tile_plans(int NB, int d, int64_t M, int n_extra, int64_t smem_limit, int num_sms, int64_t* out_h) {
    if (!out_h || NB < 1 || NB > 256 || d < 1 || d > SO_MAX_DIM || n_extra < 0 || n_extra >= kMaxOut || smem_limit < 0 || num_sms < 1)
        return SO_ERR_BAD_ARG;
    so_handle h;
    h.smem_optin = (int)smem_limit;
    h.num_sms = num_sms;
    GPState g;
    g.NB = NB; g.N = 8 * NB; g.d = d;
    for (int i = 0; i < 18; ++i) out_h[i] = 0;
    TmaPlan tp;
    const int rc = plan_tma(&h, g, tp);
    out_h[0] = rc;
    if (rc == SO_OK) {
        const int ns = tp.ns6 ? 6 : ((tp.warps == 16 || tp.ns2x) ? 2 : 4);
        int npass = (NB + ns * tp.RG - 1) / (ns * tp.RG);
        if (NB >= 32 && NB % 32 != 0 && tp.RG == 8 && ns >= 4) {
            short table[kMaxPass][8][kMaxSlots];
            npass = plan_rows(NB, table, ns);
        }
        out_h[1] = tp.BT; out_h[2] = tp.RG; out_h[3] = tp.CG; out_h[4] = tp.T; out_h[5] = npass; out_h[6] = tp.ring ? 1 : 0;
        if (tp.ring) {
            const RingSmem rs = ring_smem(tp.TB, tp.RG, tp.T, n_extra, (size_t)smem_limit);
            out_h[7] = rs.stages; out_h[8] = (int64_t)rs.total;
        } else {
            out_h[8] = (int64_t)tma_smem(tp.kb_pad, tp.TB, tp.RG, tp.T, n_extra).total;
        }
        out_h[9] = tp.warps;
        out_h[17] = ns;
    }
    LaunchPlan lp;
    const int rc2 = plan_launch(&h, g, M, false, n_extra, lp);
    out_h[10] = rc2;
    if (rc2 == SO_OK) {
        out_h[11] = lp.BT; out_h[12] = lp.RG; out_h[13] = lp.CG; out_h[14] = lp.T; out_h[15] = lp.npass; out_h[16] = (int64_t)lp.smem;
    }
    return SO_OK;
}

extern "C" int so_debug_row_plan(int NB, int16_t* table_h, int* npass_h) {
    if (!table_h || !npass_h || NB < 1 || NB > 8 * 4 * kMaxPass) return SO_ERR_BAD_ARG;
    short table[kMaxPass][8][kMaxSlots];
    *npass_h = plan_rows(NB, table);
    for (int p = 0; p < kMaxPass; ++p)
        for (int w = 0; w < 8; ++w)
            for (int s = 0; s < 4; ++s) table_h[(p * 8 + w) * 4 + s] = (int16_t)table[p][w][s];
    return SO_OK;
}

extern "C" int so_debug_row_plan_slots(int NB, int slots, int16_t* table_h, int* npass_h) {
    if (!table_h || !npass_h || NB < 1 || (slots != 4 && slots != 6) || (NB + 7) / 8 > slots * kMaxPass) return SO_ERR_BAD_ARG;
    short table[kMaxPass][8][kMaxSlots];
    *npass_h = plan_rows(NB, table, slots);
    for (int p = 0; p < kMaxPass; ++p)
        for (int w = 0; w < 8; ++w)
            for (int s = 0; s < kMaxSlots; ++s) table_h[(p * 8 + w) * kMaxSlots + s] = (int16_t)table[p][w][s];
    return SO_OK;
}

extern "C" int so_posterior_rows_simple(so_handle* h, int gp, const double* Xstar_d, int64_t M, double* mean_d,
                                        double* var_d, void* stream_) {
    if (!h || gp < 0 || gp >= h->max_gps || !Xstar_d || !mean_d || !var_d || M < 0) return SO_ERR_BAD_ARG;
    GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "posterior_rows_simple: GP not fitted");
    if (M == 0) return SO_OK;
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    double* inv_ls_d = nullptr;
    SO_CUDA(h, cudaMalloc(&inv_ls_d, sizeof(double) * SO_MAX_DIM));
    SO_CUDA(h, cudaMemcpyAsync(inv_ls_d, g.inv_ls, sizeof(double) * SO_MAX_DIM, cudaMemcpyHostToDevice, stream));
    k_posterior_simple<<<(unsigned)M, 256, sizeof(double) * (g.N + 2), stream>>>(g.Linv, g.alpha, g.Xs, Xstar_d, g.N, g.ld, g.d,
                                                                          g.kind, g.variance, inv_ls_d, M, mean_d, var_d);
    SO_CHECK_LAUNCH(h, "k_posterior_simple");
    SO_CUDA(h, cudaStreamSynchronize(stream));
    cudaFree(inv_ls_d);
    return SO_OK;
}

extern "C" int so_grid_define(so_handle* h, int d, const int32_t* n_h, const double* axis_values_h, void* stream_) {
    if (!h || !n_h || !axis_values_h) return SO_ERR_BAD_ARG;
    if (d < 1 || d > kGridMaxDim) return so_fail(h, SO_ERR_UNSUPPORTED, "so_grid_define: the grid fast path supports 1 <= d <= 6");
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    GridSpec& gs = h->grid;
    int total = 0;
    int64_t rows = 1;
    for (int j = 0; j < d; ++j) {
        if (n_h[j] < 1) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_define: axis with no points");
        gs.n[j] = n_h[j];
        gs.off[j] = total;
        total += n_h[j];
        rows *= n_h[j];
    }
    // reference row order (safeopt/utilities.py:50-54, meshgrid 'xy'): axis 1 slowest, then axis 0,
    // then axes 2..d-1 (fastest)
    if (d == 1) {
        gs.stride[0] = 1;
    } else {
        int64_t s = 1;
        for (int j = d - 1; j >= 2; --j) { gs.stride[j] = s; s *= n_h[j]; }
        gs.stride[0] = s; s *= n_h[0];
        gs.stride[1] = s;
    }
    if (total > gs.cap) {
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (gs.axis) cudaFree(gs.axis);
        gs.axis = nullptr;
        SO_CUDA(h, cudaMalloc(&gs.axis, sizeof(double) * total + sizeof(int) * total));
        gs.cap = total;
    }
    SO_CUDA(h, cudaMemcpyAsync(gs.axis, axis_values_h, sizeof(double) * total, cudaMemcpyHostToDevice, stream));
    std::vector<int> axis_of(total);
    for (int j = 0; j < d; ++j)
        for (int i = 0; i < n_h[j]; ++i) axis_of[gs.off[j] + i] = j;
    SO_CUDA(h, cudaMemcpyAsync(reinterpret_cast<int*>(gs.axis + gs.cap), axis_of.data(), sizeof(int) * total, cudaMemcpyHostToDevice, stream));
    SO_CUDA(h, cudaStreamSynchronize(stream));
    // product-table split: the fast group is the longest low-order suffix of the row order with <= 4096 rows
    {
        int order[SO_MAX_DIM];      // slowest ... fastest
        if (d == 1) order[0] = 0;
        else { order[0] = 1; order[1] = 0; for (int j = 2; j < d; ++j) order[j] = j; }
        for (int j = 0; j < d; ++j) gs.in_fast[j] = 0;
        int64_t fr = 1;
        for (int k = d - 1; k >= 0; --k) {
            const int j = order[k];
            if (k != d - 1 && fr * n_h[j] > 4096) break;
            fr *= n_h[j];
            gs.in_fast[j] = 1;
        }
        gs.fast_rows = fr;
        gs.slow_rows = rows / fr;
    }
    gs.d = d; gs.total = total; gs.rows = rows; gs.defined = true;
    for (auto& g : h->gps) { g.grid_ready = false; g.f32_ready = false; }
    return SO_OK;
}

// Two-level product tables of the grid path (Pfast: fast_rows x Npad, Pslow: slow_rows x Npad, back to back in g.P2).
// Only the slow rows [s_lo, s_hi) are (re)computed: a rank of an R-way run needs 1/R of them.
static int build_product_tables(so_handle* h, GPState& g, const double* inv_ls_d, cudaStream_t stream, int64_t s_lo, int64_t s_hi) {
    GridSpec& gs = h->grid;
    const int Npad = 8 * g.NB;
    const int64_t trows = gs.fast_rows + gs.slow_rows;
    const size_t need2 = (size_t)trows * Npad;
    if (need2 > g.capP2) {
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (g.P2) cudaFree(g.P2);
        g.P2 = nullptr;
        const size_t cap2 = (size_t)trows * g.capN;
        SO_CUDA(h, cudaMalloc(&g.P2, sizeof(double) * cap2));
        g.capP2 = cap2;
    }
    TableSpec ts;
    ts.d = gs.d;
    for (int j = 0; j < kGridMaxDim; ++j) {
        ts.n[j] = j < gs.d ? gs.n[j] : 1;
        ts.off[j] = j < gs.d ? gs.off[j] : 0;
        ts.stride[j] = j < gs.d ? gs.stride[j] : 1;
        ts.in_fast[j] = j < gs.d ? gs.in_fast[j] : 0;
    }
    ts.fast_rows = gs.fast_rows; ts.slow_rows = gs.slow_rows;
    k_grid_tables2<<<(unsigned)(gs.fast_rows + (s_hi - s_lo)), 128, 0, stream>>>(ts, gs.axis, g.Xs, g.P2, g.P2 + (size_t)gs.fast_rows * Npad,
                                                                                 g.N, Npad, g.d, g.variance, inv_ls_d, s_lo);
    SO_CHECK_LAUNCH(h, "k_grid_tables2");
    return SO_OK;
}

extern "C" int so_grid_prepare(so_handle* h, int gp, void* stream_) {
    if (!h) return SO_ERR_BAD_ARG;
    return so_grid_prepare_rows(h, gp, 0, h->grid.defined ? h->grid.rows : 0, stream_);
}

extern "C" int so_grid_prepare_rows(so_handle* h, int gp, int64_t row0, int64_t M, void* stream_) {
    if (!h || gp < 0 || gp >= h->max_gps) return SO_ERR_BAD_ARG;
    GPState& g = h->gps[gp];
    GridSpec& gs = h->grid;
    if (!gs.defined) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_prepare: no grid defined");
    if (row0 < 0 || M < 0 || row0 + M > gs.rows) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_prepare_rows: rows outside the grid");
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "so_grid_prepare: GP not fitted");
    if (g.kind != SO_KERNEL_RBF) return so_fail(h, SO_ERR_UNSUPPORTED, "so_grid_prepare: separable tables need an RBF kernel");
    if (g.d != gs.d) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_prepare: grid dimension differs from the GP input dimension");
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int Npad = 8 * g.NB;
    // ---- per-axis factor tables (expander kernel, fallback posterior kernel)
    const size_t need = (size_t)gs.total * g.capN + SO_MAX_DIM;
    if (need > g.capE) {
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (g.E) cudaFree(g.E);
        g.E = nullptr;
        SO_CUDA(h, cudaMalloc(&g.E, sizeof(double) * need));
        g.capE = need;
    }
    double* inv_ls_d = g.E + (size_t)gs.total * g.capN;
    SO_CUDA(h, cudaMemcpyAsync(inv_ls_d, g.inv_ls, sizeof(double) * SO_MAX_DIM, cudaMemcpyHostToDevice, stream));
    dim3 grd((Npad + 127) / 128, gs.total);
    k_grid_tables<<<grd, 128, 0, stream>>>(gs.axis, g.Xs, g.E, gs.total, g.N, Npad, g.d,
                                           reinterpret_cast<const int*>(gs.axis + gs.cap), g.variance, inv_ls_d);
    SO_CHECK_LAUNCH(h, "k_grid_tables");

    // ---- tables of the TMA kernel: product tables -> fragment-ordered fast table + scaled operands A'(s)
    g.tma_ready = false;
    TmaPlan pl;
    const int64_t trows = gs.fast_rows + gs.slow_rows;
    const size_t a_stride = (tri_blocks(g.NB) + 1) * 32;
    // scaled operands only for the slow indices the caller's rows touch (a rank's row block)
    const int64_t s_lo = M > 0 ? row0 / gs.fast_rows : 0;
    const int64_t s_hi = M > 0 ? (row0 + M - 1) / gs.fast_rows + 1 : 0;
    const int64_t n_slow = s_hi - s_lo;
    const size_t ap_elems = (size_t)n_slow * a_stride + 128;   // the A prefetch runs up to 3 blocks past a row
    // The scaled-operand table costs n_slow x ~N^2/2 x 8 bytes (676 MB at config 4, 4x that at N = 512): above the limit
    // (8 GB by default, SO_APRIME_LIMIT_MB overrides it) the grid path drops to the per-axis table kernel (k_posterior<GRID>,
    // slower: it spends fp64 multiplies on generating kernel rows) -- and says so, once.
    size_t ap_limit = (size_t)8 << 30;
    if (const char* v = std::getenv("SO_APRIME_LIMIT_MB")) ap_limit = (size_t)std::strtoull(v, nullptr, 10) << 20;
    const bool planned = plan_tma(h, g, pl) == SO_OK && trows <= 2147483647 && n_slow <= 65535 &&
                         (size_t)trows * Npad * sizeof(double) <= ((size_t)4 << 30);
    const bool fits = planned && ap_elems * sizeof(double2) <= ap_limit;
    if (!fits) {
        static bool said = false;
        if (!said) {
            said = true;
            fprintf(stderr, "safeopt_b200: grid path without the scaled-operand table (needs %.1f MB for %lld slow indices at N = %d, "
                            "limit %.1f MB%s): using the per-axis table kernel\n",
                    ap_elems * sizeof(double2) / 1048576.0, (long long)n_slow, g.N, ap_limit / 1048576.0,
                    planned ? "" : "; tile plan or table size out of range");
        }
    }
    if (fits) {
        int rc2 = build_product_tables(h, g, inv_ls_d, stream, s_lo, s_hi);
        if (rc2) return rc2;
        double* Pfast = g.P2;
        double* Pslow = g.P2 + (size_t)gs.fast_rows * Npad;

        const int tpb = (int)((gs.fast_rows + pl.T - 1) / pl.T);
        const size_t pf_elems = (size_t)tpb * pl.kb_pad * pl.TB * 32;
        if (pf_elems > g.capPfFrag) {
            SO_CUDA(h, cudaStreamSynchronize(stream));
            if (g.PfFrag) cudaFree(g.PfFrag);
            g.PfFrag = nullptr;
            SO_CUDA(h, cudaMalloc(&g.PfFrag, sizeof(double2) * pf_elems * 2));
            g.capPfFrag = pf_elems * 2;
        }
        if (ap_elems > g.capAprime) {
            SO_CUDA(h, cudaStreamSynchronize(stream));
            if (g.Aprime) cudaFree(g.Aprime);
            g.Aprime = nullptr;
            const size_t cap = ap_elems + ap_elems / 4;
            SO_CUDA(h, cudaMalloc(&g.Aprime, sizeof(double2) * cap));
            g.capAprime = cap;
        }
        const size_t pf_blocks = (pf_elems + 255) / 256;
        k_pffrag<<<(unsigned)(pf_blocks < 4096 ? pf_blocks : 4096), 256, 0, stream>>>(Pfast, g.PfFrag, gs.fast_rows, g.N, Npad, pl.T,
                                                                                      pl.TB, pl.kb_pad, tpb);
        SO_CHECK_LAUNCH(h, "k_pffrag");
        if (n_slow > 0) {
            dim3 grd2((unsigned)((a_stride + 255) / 256), (unsigned)n_slow);
            k_aprime<<<grd2, 256, 0, stream>>>(g.Afrag, Pslow, g.Aprime, g.NB, a_stride, s_lo);
            SO_CHECK_LAUNCH(h, "k_aprime");
        }
        SO_CUDA(h, cudaMemsetAsync(g.Aprime + (size_t)n_slow * a_stride, 0, sizeof(double2) * 128, stream));
        g.ap_s0 = s_lo; g.ap_s1 = s_hi;
        g.a_stride = a_stride; g.tma_T = pl.T; g.tma_tpb = tpb; g.tma_BT = pl.BT; g.tma_RG = pl.RG; g.tma_CG = pl.CG;
        g.tma_kb_pad = pl.kb_pad; g.tma_warps = pl.warps; g.tma_ring = pl.ring; g.tma_ns2x = pl.ns2x; g.tma_ns6 = pl.ns6; g.tma_ready = true;
    }
    g.grid_ready = true;
    return SO_OK;
}

extern "C" int so_grid_rows(so_handle* h, int64_t row0, int64_t M, double* X_d, void* stream_) {
    if (!h || !X_d || M < 0) return SO_ERR_BAD_ARG;
    GridSpec& gs = h->grid;
    if (!gs.defined) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_rows: no grid defined");
    if (row0 < 0 || row0 + M > gs.rows) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_rows: rows outside the grid");
    if (M == 0) return SO_OK;
    DeviceGuard guard(h->device);
    GridDecode gd;
    gd.d = gs.d;
    for (int j = 0; j < kGridMaxDim; ++j) {
        gd.n[j] = j < gs.d ? gs.n[j] : 1;
        gd.off[j] = j < gs.d ? gs.off[j] : 0;
        gd.stride[j] = j < gs.d ? gs.stride[j] : 1;
    }
    k_grid_rows<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(gd, gs.axis, row0, M, X_d);
    SO_CHECK_LAUNCH(h, "k_grid_rows");
    return SO_OK;
}

// ---------------------------------------------------------------- fp32 arithmetic mode (tcgen05 / TMEM), grid path
extern "C" int so_grid_prepare_f32(so_handle* h, int gp, int64_t row0, int64_t M, void* stream_) {
    if (!h || gp < 0 || gp >= h->max_gps) return SO_ERR_BAD_ARG;
    GPState& g = h->gps[gp];
    GridSpec& gs = h->grid;
    if (!gs.defined) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_prepare_f32: no grid defined");
    if (row0 < 0 || M < 0 || row0 + M > gs.rows) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_prepare_f32: rows outside the grid");
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "so_grid_prepare_f32: GP not fitted");
    if (g.kind != SO_KERNEL_RBF) return so_fail(h, SO_ERR_UNSUPPORTED, "so_grid_prepare_f32: separable tables need an RBF kernel");
    if (g.d != gs.d) return so_fail(h, SO_ERR_BAD_ARG, "so_grid_prepare_f32: grid dimension differs from the GP input dimension");
    const int nslab = f32_nslab(g.N), Np = nslab * kF32SlabK;
    if (Np > kF32MaxNp) return so_fail(h, SO_ERR_CAPACITY, "so_grid_prepare_f32: the fp32 tensor-core path holds N <= 256 (TMEM accumulator columns)");
    if (!g.E) return so_fail(h, SO_ERR_NOT_FITTED, "so_grid_prepare_f32: call so_grid_prepare(_rows) first (per-axis tables, hyper-parameters)");
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int Npad = 8 * g.NB;
    const double* inv_ls_d = g.E + (size_t)gs.total * g.capN;     // left there by so_grid_prepare_rows
    const int tpb = (int)((gs.fast_rows + kF32TileRows - 1) / kF32TileRows);
    const int64_t s_lo = M > 0 ? row0 / gs.fast_rows : 0;
    const int64_t s_hi = M > 0 ? (row0 + M - 1) / gs.fast_rows + 1 : 0;
    const int64_t n_slow = s_hi - s_lo;
    int rc = build_product_tables(h, g, inv_ls_d, stream, s_lo, s_hi);
    if (rc) return rc;
    if (n_slow > 65535) return so_fail(h, SO_ERR_CAPACITY, "so_grid_prepare_f32: more than 65535 slow indices in one row block");
    const size_t a_floats = (size_t)tpb * nslab * (kF32ASlabBytes / 4);
    const size_t b_stride = f32_b_bytes(Np);
    const size_t b_floats = (size_t)(n_slow > 0 ? n_slow : 1) * (b_stride / 4);
    if (b_floats * 4 > ((size_t)16 << 30)) return so_fail(h, SO_ERR_CAPACITY, "so_grid_prepare_f32: operand table above 16 GB");
    if (a_floats > g.capF32A) {
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (g.f32_A) cudaFree(g.f32_A);
        g.f32_A = nullptr;
        SO_CUDA(h, cudaMalloc(&g.f32_A, a_floats * 4 * 2));
        g.capF32A = a_floats * 2;
    }
    if (b_floats > g.capF32B) {
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (g.f32_B) cudaFree(g.f32_B);
        g.f32_B = nullptr;
        const size_t cap = b_floats + b_floats / 4;
        SO_CUDA(h, cudaMalloc(&g.f32_B, cap * 4));
        g.capF32B = cap;
    }
    const size_t a_blocks = (a_floats / 2 + 255) / 256;
    k_f32_pack_a<<<(unsigned)(a_blocks < 2048 ? a_blocks : 2048), 256, 0, stream>>>(g.P2, g.f32_A, gs.fast_rows, g.N, Npad, nslab, tpb);
    SO_CHECK_LAUNCH(h, "k_f32_pack_a");
    if (n_slow > 0) {
        const size_t plane0 = (size_t)8 * Np * 4;
        dim3 grd((unsigned)((plane0 + 255) / 256), (unsigned)n_slow);
        k_f32_pack_b<<<grd, 256, 0, stream>>>(g.Linv, g.ld, g.P2 + (size_t)gs.fast_rows * Npad, Npad, g.f32_B, b_stride / 4, s_lo, g.N, Np);
        SO_CHECK_LAUNCH(h, "k_f32_pack_b");
    }
    // transposed fp64 fast table of the mean kernel
    const int64_t Fpad = (gs.fast_rows + 127) / 128 * 128;
    const size_t pt = (size_t)Fpad * Npad;
    if (pt > g.capF32PfT) {
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (g.f32_PfT) cudaFree(g.f32_PfT);
        g.f32_PfT = nullptr;
        SO_CUDA(h, cudaMalloc(&g.f32_PfT, sizeof(double) * (size_t)Fpad * g.capN));
        g.capF32PfT = (size_t)Fpad * g.capN;
    }
    k_transpose_pfast<<<dim3((unsigned)(Fpad / 32), (unsigned)((g.N + 31) / 32)), dim3(32, 8), 0, stream>>>(g.P2, g.f32_PfT, gs.fast_rows, Fpad,
                                                                                                        g.N, Npad);
    SO_CHECK_LAUNCH(h, "k_transpose_pfast");
    g.f32_Fpad = Fpad;
    g.f32_Np = Np; g.f32_tpb = tpb; g.f32_s0 = s_lo; g.f32_s1 = s_hi; g.f32_ready = true;
    return SO_OK;
}

extern "C" int so_posterior_grid_f32(so_handle* h, int n, const int* gps_h, int64_t row0, int64_t M, double beta,
                                     const double* fmin_h, double* const* mean_dh, double* const* var_dh, double* Q_d,
                                     int q_stride, const int* q_col_h, uint8_t* S_d, int safe_mode, void* stream_) {
    if (!h || !gps_h || !fmin_h || !q_col_h) return SO_ERR_BAD_ARG;
    if (n < 1 || n > kMaxOut) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid_f32: 1 <= n <= 4");
    if (gps_h[0] < 0 || gps_h[0] >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid_f32: gp index out of range");
    GPState& g = h->gps[gps_h[0]];
    if (!g.fitted || !g.f32_ready) return so_fail(h, SO_ERR_NOT_FITTED, "posterior_grid_f32: so_grid_prepare_f32 not called after the fit");
    if (M < 0 || row0 < 0 || row0 + M > h->grid.rows) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid_f32: rows outside the grid");
    if (M == 0) return SO_OK;
    if (safe_mode < SO_SAFE_NONE || safe_mode > SO_SAFE_AND) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid_f32: bad safe_mode");
    const int64_t F = h->grid.fast_rows;
    if (row0 / F < g.f32_s0 || (row0 + M - 1) / F >= g.f32_s1)
        return so_fail(h, SO_ERR_NOT_FITTED, "posterior_grid_f32: rows outside the range given to so_grid_prepare_f32");
    F32Params fp;
    PostParams& p = fp.p;
    p.N = g.N; p.NB = g.NB; p.d = g.d; p.kind = g.kind; p.zvec = g.zvec; p.variance = g.variance;
    p.M = M; p.row0 = row0; p.beta = beta; p.Q = Q_d; p.q_stride = q_stride; p.S = S_d; p.safe_mode = safe_mode;
    p.n_out = n;
    p.fmin = fmin_h[0]; p.q_col = q_col_h[0];
    p.mean = mean_dh ? mean_dh[0] : nullptr; p.var = var_dh ? var_dh[0] : nullptr;
    for (int o = 0; o < kMaxOut - 1; ++o) { p.zvec_x[o] = nullptr; p.fmin_x[o] = 0.0; p.mean_x[o] = nullptr; p.var_x[o] = nullptr; p.q_col_x[o] = 0; }
    for (int o = 1; o < n; ++o) {
        if (gps_h[o] < 0 || gps_h[o] >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid_f32: gp index out of range");
        const GPState& e = h->gps[gps_h[o]];
        if (!e.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "posterior_grid_f32: GP not fitted");
        bool same = e.N == g.N && e.d == g.d && e.kind == g.kind && e.variance == g.variance && e.noise == g.noise;
        for (int j = 0; j < g.d && same; ++j) same = e.inv_ls[j] == g.inv_ls[j];
        if (!same) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid_f32: the GPs do not share size, kernel and noise");
        p.zvec_x[o - 1] = e.zvec; p.fmin_x[o - 1] = fmin_h[o]; p.q_col_x[o - 1] = q_col_h[o];
        p.mean_x[o - 1] = mean_dh ? mean_dh[o] : nullptr; p.var_x[o - 1] = var_dh ? var_dh[o] : nullptr;
    }
    for (int o = 0; o < n; ++o)
        if (Q_d && (q_col_h[o] < 0 || q_col_h[o] + 2 > q_stride)) return so_fail(h, SO_ERR_BAD_ARG, "posterior_grid_f32: Q column out of range");
    fp.Aop = reinterpret_cast<const unsigned char*>(g.f32_A);
    fp.Bop = reinterpret_cast<const unsigned char*>(g.f32_B);
    fp.Np = g.f32_Np; fp.nslab = g.f32_Np / kF32SlabK; fp.b_stride = f32_b_bytes(g.f32_Np);
    fp.s0 = g.f32_s0; fp.fast_rows = F; fp.tpb = g.f32_tpb;
    fp.s_lo = row0 / F; fp.s_hi = (row0 + M - 1) / F + 1;
    // one fast tile per CTA, `lanes` CTAs per fast tile walking the slow indices; A resident when two B stages still fit beside it
    int lanes = h->num_sms / fp.tpb;
    if (lanes < 1) lanes = 1;
    if ((int64_t)lanes > fp.s_hi - fp.s_lo) lanes = (int)(fp.s_hi - fp.s_lo);
    fp.lanes = lanes;
    const char* res_env = std::getenv("SO_F32_A_RESIDENT");           // "0": always stream A (A/B measurements)
    fp.a_resident = !(res_env && res_env[0] == '0') && f32_smem(fp.Np, 2, true).total <= (size_t)h->smem_optin ? 1 : 0;
    int stages = 4;
    while (stages > 1 && f32_smem(fp.Np, stages, fp.a_resident != 0).total > (size_t)h->smem_optin) --stages;
    if (f32_smem(fp.Np, stages, fp.a_resident != 0).total > (size_t)h->smem_optin) return so_fail(h, SO_ERR_CAPACITY, "posterior_grid_f32: shared memory");
    fp.stages = stages;
    p.ntiles = (int64_t)fp.tpb * (fp.s_hi - fp.s_lo);
    DeviceGuard guard(h->device);
    static int configured_for = -1;
    if (configured_for != h->device) {
        SO_CUDA(h, cudaFuncSetAttribute(k_posterior_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        configured_for = h->device;
    }
    // ---- the means in fp64 first (k_mean_grid), into the caller's planes or a scratch plane
    MeanParams mq;
    {
        size_t need = 0;
        for (int o = 0; o < n; ++o)
            if (!(mean_dh && mean_dh[o])) need += (size_t)M;
        if (need > h->f32_mean_cap) {
            SO_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream_));
            if (h->f32_mean_scratch) cudaFree(h->f32_mean_scratch);
            h->f32_mean_scratch = nullptr;
            SO_CUDA(h, cudaMalloc(&h->f32_mean_scratch, sizeof(double) * need));
            h->f32_mean_cap = need;
        }
        double* scratch = h->f32_mean_scratch;
        for (int o = 0; o < kMaxOut; ++o) { mq.alpha[o] = nullptr; mq.mean[o] = nullptr; fp.mean_in[o] = nullptr; }
        for (int o = 0; o < n; ++o) {
            double* dst = (mean_dh && mean_dh[o]) ? mean_dh[o] : scratch;
            if (!(mean_dh && mean_dh[o])) scratch += M;
            mq.mean[o] = dst;
            fp.mean_in[o] = dst;
            mq.alpha[o] = h->gps[gps_h[o]].alpha;
        }
        const int Npad = 8 * g.NB;
        mq.PfastT = g.f32_PfT; mq.Pslow = g.P2 + (size_t)F * Npad; mq.n_out = n; mq.N = g.N; mq.ldp = Npad;
        mq.TR = g.N <= 128 ? 128 : 64; mq.SB = kMeanCols / n; mq.N4 = (g.N + 3) / 4 * 4;
        mq.Fpad = g.f32_Fpad; mq.fast_rows = F; mq.row0 = row0; mq.M = M;
        mq.s_lo = row0 / F; mq.s_hi = (row0 + M - 1) / F + 1;
        const int64_t gx = (F + mq.TR - 1) / mq.TR;
        int64_t gy = ((int64_t)h->num_sms + gx - 1) / gx;                 // one CTA per SM: each keeps its slice of the fast table in shared memory
        const int64_t groups = (mq.s_hi - mq.s_lo + mq.SB - 1) / mq.SB;
        if (gy > groups) gy = groups;
        if (gy > 65535) gy = 65535;
        const size_t msm = mean_smem_bytes(mq.N4, mq.TR);
        static int mean_configured_for = -1;
        if (mean_configured_for != h->device) {
            SO_CUDA(h, cudaFuncSetAttribute(k_mean_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
            mean_configured_for = h->device;
        }
        k_mean_grid<<<dim3((unsigned)gx, (unsigned)gy), kMeanThreads, msm, (cudaStream_t)stream_>>>(mq);
        SO_CHECK_LAUNCH(h, "k_mean_grid");
    }
    // SO_F32_MULTICAST=1: the tpb CTAs that share a slow index form a cluster and multicast the B slabs (1 / tpb of the L2 -> SM
    // operand traffic, which bounds the kernel: 189 MB per launch at config 3).  Measured at config 3 (profiles/r02_f32_variants.md):
    // 107 us with multicast against 58 us without -- a stage is free only when all four CTAs have read it, and with three stages
    // that lock-step costs more than the saved traffic -- so it is an A/B switch, off by default.
    const char* mc_env = std::getenv("SO_F32_MULTICAST");
    fp.mc = (fp.a_resident && (fp.tpb == 2 || fp.tpb == 4 || fp.tpb == 8) && mc_env && mc_env[0] == '1') ? fp.tpb : 1;
    const size_t smem_bytes = f32_smem(fp.Np, fp.stages, fp.a_resident != 0).total;
    if (fp.mc > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(fp.tpb * fp.lanes); cfg.blockDim = dim3(kF32Threads); cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = (cudaStream_t)stream_;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = fp.mc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        SO_CUDA(h, cudaLaunchKernelEx(&cfg, k_posterior_f32, fp));
        return SO_OK;
    }
    k_posterior_f32<<<fp.tpb * fp.lanes, kF32Threads, smem_bytes, (cudaStream_t)stream_>>>(fp);
    SO_CHECK_LAUNCH(h, "k_posterior_f32");
    return SO_OK;
}
