// K1 -- training-side fit on the device, fp64 throughout.
// Stands in for GPy's ExactGaussianInference.inference reached from `gp.set_XY`
// (reference call sites safeopt/gp_opt.py:227, :267, :275): Ky = K(X,X) + (noise+1e-8) I,
// lower Cholesky, L^-1, alpha = Ky^-1 Y.  The explicit triangular inverse (not GPy's dpotri
// full inverse) is what the posterior contraction consumes: var = k** - |L^-1 k|^2 costs
// N^2/2 FMAs per candidate instead of N^2 and is the better-conditioned form (SURVEY.md
// section 7, hard part 1).
#include "fit_cluster.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

__device__ __forceinline__ double kernel_switch(int kind, double r2, double variance) {
    switch (kind) {
        case SO_KERNEL_RBF: return kernel_of_r2<SO_KERNEL_RBF>(r2, variance);
        case SO_KERNEL_MATERN32: return kernel_of_r2<SO_KERNEL_MATERN32>(r2, variance);
        default: return kernel_of_r2<SO_KERNEL_MATERN52>(r2, variance);
    }
}

struct InvLs { double v[SO_MAX_DIM]; };

__global__ void k_scale_x(const double* __restrict__ X, double* __restrict__ Xs, int N, int Npad, int d, InvLs il) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Npad * d) return;
    int n = idx / d, j = idx - n * d;
    Xs[idx] = n < N ? X[idx] * il.v[j] : 0.0;
}

// Ky (padded with an identity block so the factorisation of the padded matrix is trivial there).
__global__ void k_build_ky(const double* __restrict__ Xs, double* __restrict__ K, int N, int Npad, int ld, int d,
                           int kind, double variance, double diag_add) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Npad || j >= Npad) return;
    double v;
    if (i < N && j < N) {
        double r2 = 0.0;
        for (int c = 0; c < d; ++c) {
            double t = Xs[i * d + c] - Xs[j * d + c];
            r2 = fma(t, t, r2);
        }
        v = kernel_switch(kind, r2, variance);
        if (i == j) v += diag_add;
    } else {
        v = (i == j) ? 1.0 : 0.0;
    }
    K[(size_t)i * ld + j] = v;
}

// ---------------------------------------------------------------- blocked Cholesky + triangular inverse
// Right-looking, panel width 32, on the identity-padded Npad x Npad matrix.  Per panel P (rows/cols p0 .. p0+nb):
//   k_chol_panel : every CTA re-factors the 32x32 diagonal block in registers (one warp, shuffles; ~2 us, cheaper than
//                  a grid-wide dependency), then the CTAs share (b) L[R,P] = A[R,P] L_PP^-T by forward substitution,
//                  one thread per row, and (c) W[P, 0:p0] <- L_PP^-1 W[P, 0:p0], one thread per column.  CTA 0 also
//                  writes L_PP and W[P,P] = L_PP^-1 (substitution on the identity).
//   k_chol_update: 32x32 tiles over all SMs:  A[R,R] -= L[R,P] L[R,P]^T  (lower tiles)  and
//                  W[R, 0:p0+nb] -= L[R,P] W[P, 0:p0+nb]  -- forward substitution on the identity done right-looking,
//                  so W ends as L^-1.  Both updates have the same shape and run in the same launch.
// N = 256: 15 launches, ~0.1 ms (the previous single-CTA column-by-column factorisation + per-column inverse: 1.25 ms).
constexpr int kPB = 32;

__device__ __forceinline__ void chol_diag_block(const double* __restrict__ K, double* __restrict__ W, int ld, int p0, int nb,
                                                double (*sL)[kPB + 1], double* __restrict__ sRinv, bool write_back, int* status) {
    const int lane = threadIdx.x & 31;
    double row[kPB];
#pragma unroll
    for (int k = 0; k < kPB; ++k) {
        double v = (k == lane) ? 1.0 : 0.0;
        if (lane < nb && k <= lane) v = K[(size_t)(p0 + lane) * ld + p0 + k];
        row[k] = v;
    }
    int bad = 0;
#pragma unroll
    for (int j = 0; j < kPB; ++j) {
        double djj = __shfl_sync(0xffffffffu, row[j], j);
        if (!(djj > 0.0) || isinf(djj)) { bad = 1; djj = 1.0; }
        const double ljj = sqrt(djj);
        row[j] = (lane == j) ? ljj : row[j] / ljj;
#pragma unroll
        for (int k = j + 1; k < kPB; ++k) {
            const double lkj = __shfl_sync(0xffffffffu, row[j], k);
            row[k] = fma(-row[j], lkj, row[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kPB; ++k) {
        sL[lane][k] = (k <= lane) ? row[k] : 0.0;
        if (k == lane) sRinv[lane] = 1.0 / row[k];      // static register index
    }
    if (bad && lane == 0 && write_back) *status = SO_ERR_NOT_PD;
    __syncwarp();
    if (!write_back) return;
#pragma unroll
    for (int k = 0; k < kPB; ++k)
        if (lane < nb && k <= lane) const_cast<double*>(K)[(size_t)(p0 + lane) * ld + p0 + k] = row[k];
    // W[P,P] = L_PP^-1: lane c solves L_PP x = e_c
    double x[kPB];
#pragma unroll
    for (int i = 0; i < kPB; ++i) {
        double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < i; k += 2) {
            s0 = fma(-sL[i][k], x[k], s0);
            s1 = fma(-sL[i][k + 1], x[k + 1], s1);
        }
        if (i & 1) s0 = fma(-sL[i][i - 1], x[i - 1], s0);
        x[i] = (s0 + s1) / sL[i][i];
    }
#pragma unroll
    for (int i = 0; i < kPB; ++i)
        if (i < nb && lane < nb) W[(size_t)(p0 + i) * ld + p0 + lane] = x[i];
}

__global__ void __launch_bounds__(256) k_chol_panel(double* __restrict__ K, double* __restrict__ W, int Npad, int ld, int p0,
                                                    int nb, int* status) {
    __shared__ double sL[kPB][kPB + 1];
    __shared__ double sRinv[kPB];
    if (threadIdx.x < 32) chol_diag_block(K, W, ld, p0, nb, sL, sRinv, blockIdx.x == 0, status);
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = p0 + nb + t;
    if (r < Npad) {                 // only reached with nb == 32 (a narrower panel is the last one)
        double* arow = K + (size_t)r * ld + p0;
        double x[kPB];
#pragma unroll
        for (int c = 0; c < kPB; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(arow + c);
            x[c] = v.x; x[c + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < kPB; ++c) {
            double s0 = x[c], s1 = 0.0;
#pragma unroll
            for (int k = 0; k + 1 < c; k += 2) {
                s0 = fma(-sL[c][k], x[k], s0);
                s1 = fma(-sL[c][k + 1], x[k + 1], s1);
            }
            if (c & 1) s0 = fma(-sL[c][c - 1], x[c - 1], s0);
            x[c] = (s0 + s1) * sRinv[c];
        }
#pragma unroll
        for (int c = 0; c < kPB; c += 2) *reinterpret_cast<double2*>(arow + c) = make_double2(x[c], x[c + 1]);
    }
    if (t < p0) {                   // column t of W[P, 0:p0]
        double x[kPB];
#pragma unroll
        for (int i = 0; i < kPB; ++i) x[i] = i < nb ? W[(size_t)(p0 + i) * ld + t] : 0.0;
#pragma unroll
        for (int i = 0; i < kPB; ++i) {
            double s0 = x[i], s1 = 0.0;
#pragma unroll
            for (int k = 0; k + 1 < i; k += 2) {
                s0 = fma(-sL[i][k], x[k], s0);
                s1 = fma(-sL[i][k + 1], x[k + 1], s1);
            }
            if (i & 1) s0 = fma(-sL[i][i - 1], x[i - 1], s0);
            x[i] = (s0 + s1) * sRinv[i];
        }
#pragma unroll
        for (int i = 0; i < kPB; ++i)
            if (i < nb) W[(size_t)(p0 + i) * ld + t] = x[i];
    }
}

// One 32x32 output tile per CTA.  blockIdx.y = row block below the panel, blockIdx.x = column block (<= row block).
__global__ void __launch_bounds__(256) k_chol_update(double* __restrict__ K, double* __restrict__ W, int Npad, int ld, int p0) {
    const int pblk = p0 / kPB;
    const int rb = pblk + 1 + blockIdx.y, cb = blockIdx.x;
    if (cb > rb) return;
    __shared__ double sA[kPB][kPB + 1];     // L[rb rows, P]
    __shared__ double sB[kPB][kPB + 1];     // [k][j]: L[cb rows, P]^T (A tiles) or W[P, cb cols] (W tiles)
    const int tid = threadIdx.x;
    const bool a_tile = cb > pblk;
    for (int e = tid; e < kPB * kPB; e += 256) {
        const int i = e >> 5, k = e & 31;
        const int r = rb * kPB + i;
        sA[i][k] = r < Npad ? K[(size_t)r * ld + p0 + k] : 0.0;
        if (a_tile) {
            const int rc = cb * kPB + i;
            sB[k][i] = rc < Npad ? K[(size_t)rc * ld + p0 + k] : 0.0;
        } else {
            const int c = cb * kPB + k;      // here e = (row i of the panel, column k of the tile)
            sB[i][k] = (c < Npad && (cb < pblk || k <= i)) ? W[(size_t)(p0 + i) * ld + c] : 0.0;
        }
    }
    __syncthreads();
    const int i = tid >> 3, j0 = (tid & 7) * 4;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < kPB; ++k) {
        const double a = sA[i][k];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(a, sB[k][j0 + u], acc[u]);
    }
    const int r = rb * kPB + i;
    if (r >= Npad) return;
    double* dst = (a_tile ? K : W) + (size_t)r * ld + cb * kPB + j0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (cb * kPB + j0 + u >= Npad) break;
        if (!a_tile && cb == pblk) dst[u] = -acc[u];          // first touch of W[R, P]: starts from zero
        else dst[u] -= acc[u];
    }
}

// alpha = Linv^T (Linv y), one CTA; padding entries are zero.
__global__ void __launch_bounds__(1024) k_alpha(const double* __restrict__ Linv, const double* __restrict__ Y,
                                                double* __restrict__ alpha, double* __restrict__ zvec, int N, int Npad, int ld) {
    extern __shared__ double w[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = warp; i < N; i += nwarps) {
        const double* row = Linv + (size_t)i * ld;
        double part = 0.0;
        for (int k = lane; k <= i; k += 32) part = fma(row[k], Y[k], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) w[i] = part;
    }
    __syncthreads();
    for (int c = tid; c < Npad; c += blockDim.x) {
        double s = 0.0;
        if (c < N)
            for (int i = c; i < N; ++i) s = fma(Linv[(size_t)i * ld + c], w[i], s);
        alpha[c] = s;
        zvec[c] = c < N ? w[c] : 0.0;
    }
}

// Pack L^-1 into DMMA A-fragment order.  Block (i, kb), kb <= i, lives at index i(i+1)/2 + kb and
// holds 32 double2: lane l -> (Linv[8i + l/4][8kb + 2(l%4)], Linv[8i + l/4][8kb + 2(l%4) + 1]).
// The k permutation inside an 8-block (even slots in MMA step 0, odd slots in step 1) is shared
// with the B operand built by the posterior kernel, so a lane's two values are adjacent doubles.
__global__ void k_pack_afrag(const double* __restrict__ Linv, double2* __restrict__ Afrag, int N, int ld, int i0) {
    const int i = i0 + blockIdx.y, kb = blockIdx.x;
    if (kb > i) return;
    const int lane = threadIdx.x;
    const int r = 8 * i + (lane >> 2), c0 = 8 * kb + 2 * (lane & 3);
    double v0 = 0.0, v1 = 0.0;
    if (r < N) {
        if (c0 < N && c0 <= r) v0 = Linv[(size_t)r * ld + c0];
        if (c0 + 1 < N && c0 + 1 <= r) v1 = Linv[(size_t)r * ld + c0 + 1];
    }
    Afrag[((size_t)i * (i + 1) / 2 + kb) * 32 + lane] = make_double2(v0, v1);
}

template <typename T>
int grow(so_handle* h, T*& ptr, size_t count) {
    if (ptr) { cudaFree(ptr); ptr = nullptr; }
    SO_CUDA(h, cudaMalloc(&ptr, count * sizeof(T)));
    return SO_OK;
}

// Cluster size for the one-launch fit of N points (0 = use the kernel-per-panel path): 16 CTAs where the device can co-schedule
// such a cluster (non-portable size), else 8.
int fit_cluster_size(so_handle* h, int N) {
    if (N > kFcMaxN) return 0;
    const char* v = std::getenv("SO_FIT_CLUSTER");
    if (v && std::string(v) == "0") return 0;
    static int cached_dev = -1, cached = 0;
    if (cached_dev == h->device) return cached;
    cached_dev = h->device;
    cached = 0;
    int want = 16;
    if (v && std::atoi(v) > 0) want = std::atoi(v);
    if (cudaFuncSetAttribute(k_fit_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFcDynSmem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    for (int cs = want; cs >= 2; cs /= 2) {
        if (cs > 8 && cudaFuncSetAttribute(k_fit_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kFcThreads); cfg.dynamicSmemBytes = kFcDynSmem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, k_fit_cluster, &cfg) == cudaSuccess && n >= 1) { cached = cs; break; }
        cudaGetLastError();
    }
    if (std::getenv("SO_FIT_VERBOSE"))
        fprintf(stderr, "safeopt_b200: one-launch fit uses a cluster of %d CTAs%s\n", cached, cached ? "" : " (unavailable: kernel-per-panel fit)");
    return cached;
}

}  // namespace

static int fit_impl(so_handle* h, int gp, const double* X_h, const double* Y_h, int N, int d, int kernel_kind,
                    const double* lengthscale_h, double variance, double noise_var, void* stream_, bool async) {
    if (!h) return SO_ERR_BAD_ARG;
    if (gp < 0 || gp >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "so_fit: gp index out of range");
    if (!X_h || !Y_h || !lengthscale_h || N < 1) return so_fail(h, SO_ERR_BAD_ARG, "so_fit: null input or N < 1");
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "so_fit: input dimension must be in [1, 16]");
    if (kernel_kind < SO_KERNEL_RBF || kernel_kind > SO_KERNEL_MATERN52)
        return so_fail(h, SO_ERR_UNSUPPORTED, "so_fit: unknown kernel family");
    if (!(variance > 0.0) || !(noise_var >= 0.0)) return so_fail(h, SO_ERR_BAD_ARG, "so_fit: variance must be > 0, noise >= 0");
    if (N > 2048) return so_fail(h, SO_ERR_CAPACITY, "so_fit: N > 2048 not supported by the single-CTA factorisation");
    for (int j = 0; j < d; ++j)
        if (!(lengthscale_h[j] > 0.0)) return so_fail(h, SO_ERR_BAD_ARG, "so_fit: lengthscales must be > 0");

    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    GPState& g = h->gps[gp];
    const int NB = (N + 7) / 8, Npad = 8 * NB;
    if (Npad > g.capN || d != g.d) {
        // capacity grows geometrically so that the one-point-per-iteration BO loop rarely reallocates
        int cap = g.capN > 0 ? g.capN : 64;
        while (cap < Npad) cap = cap + cap / 2;
        cap = (cap + 63) / 64 * 64;
        const int capNB = cap / 8;
        SO_CUDA(h, cudaStreamSynchronize(stream));
        int rc;
        if ((rc = grow(h, g.X, (size_t)cap * d))) return rc;
        if ((rc = grow(h, g.Xs, (size_t)cap * d))) return rc;
        if ((rc = grow(h, g.Y, (size_t)cap))) return rc;
        if ((rc = grow(h, g.K, (size_t)cap * cap))) return rc;
        if ((rc = grow(h, g.Linv, (size_t)cap * cap))) return rc;
        if ((rc = grow(h, g.alpha, (size_t)cap))) return rc;
        if ((rc = grow(h, g.zvec, (size_t)cap))) return rc;
        if ((rc = grow(h, g.Afrag, (tri_blocks(capNB) + 4) * 32))) return rc;
        g.capN = cap;
        g.ld = cap;
    }
    const int ld = g.ld;
    g.fitted = false;
    g.grid_ready = false;
    g.f32_ready = false;
    g.N = N; g.d = d; g.kind = kernel_kind; g.NB = NB;
    g.variance = variance; g.noise = noise_var;
    InvLs il;
    for (int j = 0; j < SO_MAX_DIM; ++j) { il.v[j] = j < d ? 1.0 / lengthscale_h[j] : 0.0; g.inv_ls[j] = il.v[j]; }

    const int csize_try = fit_cluster_size(h, N);
    if (csize_try > 0 && (size_t)N * (d + 1) * sizeof(double) <= h->fit_stage_bytes) {
        // pinned staging (one slot per GP, so that asynchronous fits of several GPs do not overwrite each other's inputs)
        double* st = reinterpret_cast<double*>(static_cast<unsigned char*>(h->fit_stage_h) + (size_t)gp * h->fit_stage_bytes);
        std::memcpy(st, X_h, sizeof(double) * N * d);
        std::memcpy(st + (size_t)N * d, Y_h, sizeof(double) * N);
        SO_CUDA(h, cudaMemcpyAsync(g.X, st, sizeof(double) * N * d, cudaMemcpyHostToDevice, stream));
        SO_CUDA(h, cudaMemcpyAsync(g.Y, st + (size_t)N * d, sizeof(double) * N, cudaMemcpyHostToDevice, stream));
    } else {
        SO_CUDA(h, cudaMemcpyAsync(g.X, X_h, sizeof(double) * N * d, cudaMemcpyHostToDevice, stream));
        SO_CUDA(h, cudaMemcpyAsync(g.Y, Y_h, sizeof(double) * N, cudaMemcpyHostToDevice, stream));
    }
    // N <= 512: the whole fit in one launch of one thread-block cluster (fit_cluster.cuh); SO_FIT_CLUSTER=0 keeps the
    // kernel-per-panel version below (also used for larger N)
    const int csize = csize_try;
    if (csize > 0) {
        FitClusterParams fp;
        fp.X = g.X; fp.Y = g.Y; fp.Xs = g.Xs; fp.K = g.K; fp.W = g.Linv; fp.alpha = g.alpha; fp.zvec = g.zvec; fp.Afrag = g.Afrag;
        fp.N = N; fp.Npad = Npad; fp.NP = (Npad + kFcB - 1) / kFcB * kFcB; fp.ld = ld; fp.d = d; fp.kind = kernel_kind; fp.NB = NB;
        fp.variance = variance; fp.diag_add = noise_var + SO_JITTER;
        for (int j = 0; j < SO_MAX_DIM; ++j) fp.inv_ls[j] = il.v[j];
        fp.status = h->fit_status_d + gp;                   // mapped pinned word per GP: no status copy after the kernel
        h->fit_status_h[gp] = SO_OK;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csize); cfg.blockDim = dim3(kFcThreads); cfg.dynamicSmemBytes = kFcDynSmem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
#ifdef SO_FIT_STAMPS
        static long long* stamps_d = nullptr;
        if (!stamps_d) { cudaMalloc(&stamps_d, 256 * sizeof(long long)); }
        cudaMemsetAsync(stamps_d, 0, 256 * sizeof(long long), stream);
        fp.stamps = stamps_d;
#endif
        SO_CUDA(h, cudaLaunchKernelEx(&cfg, k_fit_cluster, fp));
#ifdef SO_FIT_STAMPS
        if (std::getenv("SO_FIT_VERBOSE")) {
            long long st_h[256];
            cudaStreamSynchronize(stream);
            cudaMemcpy(st_h, stamps_d, sizeof(st_h), cudaMemcpyDeviceToHost);
            const int nblk = fp.NP / kFcB;
            std::fprintf(stderr, "[fit stamps N=%d] phase0 %lld sync %lld |", N, st_h[1] - st_h[0], st_h[2] - st_h[1]);
            for (int pi = 0; pi < nblk; ++pi) {
                const long long* q = st_h + 2 + 6 * pi;
                std::fprintf(stderr, " p%d: factor+inv %lld solve %lld sync %lld trail %lld sync %lld |", pi, q[2] - q[0],
                             q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5]);
            }
            std::fprintf(stderr, " tail %lld total %lld cycles\n", st_h[3 + 6 * nblk] - st_h[2 + 6 * nblk], st_h[3 + 6 * nblk] - st_h[0]);
        }
#endif
        g.fitted = true;
        if (async) return SO_OK;                            // the caller reads so_fit_status after its next synchronisation
        SO_CUDA(h, cudaStreamSynchronize(stream));
        if (h->fit_status_h[gp] != SO_OK) {
            g.fitted = false;
            return so_fail(h, SO_ERR_NOT_PD, "so_fit: K + (noise + 1e-8) I is not positive definite");
        }
        return SO_OK;
    }
    k_scale_x<<<(Npad * d + 255) / 256, 256, 0, stream>>>(g.X, g.Xs, N, Npad, d, il);
    dim3 blk(16, 16), grd((Npad + 15) / 16, (Npad + 15) / 16);
    k_build_ky<<<grd, blk, 0, stream>>>(g.Xs, g.K, N, Npad, ld, d, kernel_kind, variance, noise_var + SO_JITTER);
    SO_CUDA(h, cudaMemsetAsync(h->d_status, 0, sizeof(int), stream));
    static_assert(SO_OK == 0, "status word is cleared with memset");
    SO_CUDA(h, cudaMemsetAsync(g.Linv, 0, sizeof(double) * (size_t)Npad * ld, stream));
    for (int p0 = 0; p0 < Npad; p0 += kPB) {
        const int nb = Npad - p0 < kPB ? Npad - p0 : kPB;
        const int below = Npad - p0 - nb;
        const int work = below > p0 ? below : p0;
        k_chol_panel<<<work > 0 ? (work + 255) / 256 : 1, 256, 0, stream>>>(g.K, g.Linv, Npad, ld, p0, nb, h->d_status);
        if (below > 0) {
            const int nrb = (below + kPB - 1) / kPB, ncb = (Npad + kPB - 1) / kPB;
            k_chol_update<<<dim3(ncb, nrb), 256, 0, stream>>>(g.K, g.Linv, Npad, ld, p0);
        }
    }
    k_alpha<<<1, 1024, sizeof(double) * Npad, stream>>>(g.Linv, g.Y, g.alpha, g.zvec, N, Npad, ld);
    SO_CUDA(h, cudaMemsetAsync(g.Afrag, 0, sizeof(double2) * (tri_blocks(NB) + 4) * 32, stream));
    k_pack_afrag<<<dim3(NB, NB), 32, 0, stream>>>(g.Linv, g.Afrag, N, ld, 0);
    SO_CHECK_LAUNCH(h, "so_fit kernels");
    SO_CUDA(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SO_CUDA(h, cudaStreamSynchronize(stream));
    if (*h->h_status != SO_OK)
        return so_fail(h, SO_ERR_NOT_PD, "so_fit: K + (noise + 1e-8) I is not positive definite");
    g.fitted = true;
    return SO_OK;
}

extern "C" int so_fit(so_handle* h, int gp, const double* X_h, const double* Y_h, int N, int d, int kernel_kind,
                      const double* lengthscale_h, double variance, double noise_var, void* stream) {
    return fit_impl(h, gp, X_h, Y_h, N, d, kernel_kind, lengthscale_h, variance, noise_var, stream, false);
}

extern "C" int so_fit_async(so_handle* h, int gp, const double* X_h, const double* Y_h, int N, int d, int kernel_kind,
                            const double* lengthscale_h, double variance, double noise_var, void* stream) {
    return fit_impl(h, gp, X_h, Y_h, N, d, kernel_kind, lengthscale_h, variance, noise_var, stream, true);
}

extern "C" int so_fit_status(so_handle* h, int gp) {
    if (!h || gp < 0 || gp >= h->max_gps) return SO_ERR_BAD_ARG;
    const int st = h->fit_status_h[gp];
    if (st != SO_OK) {
        h->gps[gp].fitted = false;
        return so_fail(h, st, "so_fit_async: K + (noise + 1e-8) I is not positive definite");
    }
    return SO_OK;
}

// GPs that share inputs, kernel and noise (SafeOpt's constraint GPs usually do, gp_opt.py:121-130) share K, L and L^-1: the
// factorisation of `src_gp` is copied device-to-device and only z = L^-1 y, alpha = L^-T z are computed for the new targets.
extern "C" int so_fit_like(so_handle* h, int gp, int src_gp, const double* Y_h, void* stream_) {
    if (!h || !Y_h) return SO_ERR_BAD_ARG;
    if (gp < 0 || gp >= h->max_gps || src_gp < 0 || src_gp >= h->max_gps || gp == src_gp)
        return so_fail(h, SO_ERR_BAD_ARG, "so_fit_like: gp index out of range");
    GPState& s = h->gps[src_gp];
    GPState& g = h->gps[gp];
    if (!s.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "so_fit_like: source GP not fitted");
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int N = s.N, d = s.d, NB = s.NB, Npad = 8 * NB;
    if (Npad > g.capN || d != g.d) {
        const int cap = s.capN, capNB = cap / 8;
        SO_CUDA(h, cudaStreamSynchronize(stream));
        int rc;
        if ((rc = grow(h, g.X, (size_t)cap * d))) return rc;
        if ((rc = grow(h, g.Xs, (size_t)cap * d))) return rc;
        if ((rc = grow(h, g.Y, (size_t)cap))) return rc;
        if ((rc = grow(h, g.K, (size_t)cap * cap))) return rc;
        if ((rc = grow(h, g.Linv, (size_t)cap * cap))) return rc;
        if ((rc = grow(h, g.alpha, (size_t)cap))) return rc;
        if ((rc = grow(h, g.zvec, (size_t)cap))) return rc;
        if ((rc = grow(h, g.Afrag, (tri_blocks(capNB) + 4) * 32))) return rc;
        g.capN = cap;
        g.ld = cap;
    }
    g.fitted = false;
    g.grid_ready = false;
    g.f32_ready = false;
    g.tma_ready = false;
    g.N = N; g.d = d; g.kind = s.kind; g.NB = NB; g.variance = s.variance; g.noise = s.noise;
    for (int j = 0; j < SO_MAX_DIM; ++j) g.inv_ls[j] = s.inv_ls[j];
    SO_CUDA(h, cudaMemcpyAsync(g.X, s.X, sizeof(double) * N * d, cudaMemcpyDeviceToDevice, stream));
    SO_CUDA(h, cudaMemcpyAsync(g.Xs, s.Xs, sizeof(double) * Npad * d, cudaMemcpyDeviceToDevice, stream));
    SO_CUDA(h, cudaMemcpy2DAsync(g.K, sizeof(double) * g.ld, s.K, sizeof(double) * s.ld, sizeof(double) * Npad, Npad, cudaMemcpyDeviceToDevice, stream));
    SO_CUDA(h, cudaMemcpy2DAsync(g.Linv, sizeof(double) * g.ld, s.Linv, sizeof(double) * s.ld, sizeof(double) * Npad, Npad, cudaMemcpyDeviceToDevice, stream));
    SO_CUDA(h, cudaMemcpyAsync(g.Afrag, s.Afrag, sizeof(double2) * (tri_blocks(NB) + 4) * 32, cudaMemcpyDeviceToDevice, stream));
    SO_CUDA(h, cudaMemcpyAsync(g.Y, Y_h, sizeof(double) * N, cudaMemcpyHostToDevice, stream));
    k_alpha<<<1, 1024, sizeof(double) * Npad, stream>>>(g.Linv, g.Y, g.alpha, g.zvec, N, Npad, g.ld);
    SO_CHECK_LAUNCH(h, "so_fit_like kernels");
    g.fitted = true;
    return SO_OK;
}

extern "C" int so_fit_export(so_handle* h, int gp, double* L_h, double* Linv_h, double* alpha_h) {
    if (!h || gp < 0 || gp >= h->max_gps) return SO_ERR_BAD_ARG;
    GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "so_fit_export: GP not fitted");
    DeviceGuard guard(h->device);
    SO_CUDA(h, cudaDeviceSynchronize());
    const int N = g.N, ld = g.ld;
    if (L_h) {
        SO_CUDA(h, cudaMemcpy2D(L_h, sizeof(double) * N, g.K, sizeof(double) * ld, sizeof(double) * N, N, cudaMemcpyDeviceToHost));
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) L_h[(size_t)i * N + j] = 0.0;
    }
    if (Linv_h)
        SO_CUDA(h, cudaMemcpy2D(Linv_h, sizeof(double) * N, g.Linv, sizeof(double) * ld, sizeof(double) * N, N, cudaMemcpyDeviceToHost));
    if (alpha_h) SO_CUDA(h, cudaMemcpy(alpha_h, g.alpha, sizeof(double) * N, cudaMemcpyDeviceToHost));
    return SO_OK;
}

// ---------------------------------------------------------------- f4: one-point append / removal
// `add_new_data_point` (safeopt/gp_opt.py:230-255) grows the data by one row per iteration; GPy refits from
// scratch (O(N^3)).  With L and W = L^-1 resident, the bordered factorisation is O(N^2):
//   l   = W k_new                     (new row of L)          k_append_row    one warp per entry
//   lnn = sqrt(k(x,x) + noise + 1e-8 - |l|^2)
//   W[N, c] = -(sum_{i>=c} l_i W[i, c]) / lnn , W[N,N] = 1/lnn   k_append_finish one thread per column
//   z_N = W[N, :] y ;  alpha += W[N, :]^T z_N                  k_append_alpha
// followed by re-packing the one fragment block row that changed.  Removing the last point keeps the
// leading blocks of L and W as they are (they do not depend on later rows).
namespace {

__global__ void __launch_bounds__(256) k_append_row(const double* __restrict__ W, const double* __restrict__ Xs,
                                                    double* __restrict__ Lrow, int N, int ld, int d, int kind, double variance) {
    extern __shared__ double kv[];                      // k(x_new, X_i), i < N (every CTA builds its own copy)
    const double* xn = Xs + (size_t)N * d;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double r2 = 0.0;
        for (int c = 0; c < d; ++c) {
            const double t = xn[c] - Xs[(size_t)i * d + c];
            r2 = fma(t, t, r2);
        }
        kv[i] = kernel_switch(kind, r2, variance);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= N) return;
    const double* row = W + (size_t)i * ld;
    double part = 0.0;
    for (int k = lane; k <= i; k += 32) part = fma(row[k], kv[k], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) Lrow[i] = part;
}

// Deterministic block-wide sum (same tree in every CTA, so every CTA derives the same lnn).
__device__ __forceinline__ double block_sum_256(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(256) k_append_finish(double* __restrict__ K, double* __restrict__ W, int N, int ld, double kss,
                                                       int* status) {
    __shared__ double red[8];
    const double* l = K + (size_t)N * ld;               // new row of L, entries 0..N-1
    double part = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) part = fma(l[i], l[i], part);
    const double s = block_sum_256(part, red);
    double dnn = kss - s;
    if (!(dnn > 0.0) || isinf(dnn)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *status = SO_ERR_NOT_PD;
        dnn = 1.0;
    }
    const double lnn = sqrt(dnn);
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) {
        double acc0 = 0.0, acc1 = 0.0;
        int i = c;
        for (; i + 1 < N; i += 2) {
            acc0 = fma(l[i], W[(size_t)i * ld + c], acc0);
            acc1 = fma(l[i + 1], W[(size_t)(i + 1) * ld + c], acc1);
        }
        if (i < N) acc0 = fma(l[i], W[(size_t)i * ld + c], acc0);
        W[(size_t)N * ld + c] = -(acc0 + acc1) / lnn;
    } else if (c == N) {
        W[(size_t)N * ld + N] = 1.0 / lnn;
        K[(size_t)N * ld + N] = lnn;
    }
}

__global__ void __launch_bounds__(256) k_append_alpha(const double* __restrict__ W, const double* __restrict__ Y,
                                                      double* __restrict__ alpha, double* __restrict__ zvec, int N, int ld) {
    __shared__ double red[8];
    const double* w = W + (size_t)N * ld;
    double part = 0.0;
    for (int c = threadIdx.x; c <= N; c += 256) part = fma(w[c], Y[c], part);
    const double zn = block_sum_256(part, red);
    for (int c = threadIdx.x; c <= N; c += 256) alpha[c] = c < N ? fma(w[c], zn, alpha[c]) : w[c] * zn;
    if (threadIdx.x == 0) zvec[N] = zn;
}

// Rows [r0, r1) of the per-point arrays become padding: zero coordinates / targets / alpha / z.
__global__ void k_clear_rows(double* __restrict__ Xs, double* __restrict__ Y, double* __restrict__ alpha,
                             double* __restrict__ zvec, int r0, int r1, int d) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = r0 + idx / (d + 3), j = idx % (d + 3);
    if (n >= r1) return;
    if (j < d) Xs[(size_t)n * d + j] = 0.0;
    else if (j == d) Y[n] = 0.0;
    else if (j == d + 1) alpha[n] = 0.0;
    else zvec[n] = 0.0;
}

}  // namespace

extern "C" int so_fit_append(so_handle* h, int gp, const double* x_new_h, double y_new, void* stream_) {
    if (!h || !x_new_h) return SO_ERR_BAD_ARG;
    if (gp < 0 || gp >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "so_fit_append: gp index out of range");
    GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "so_fit_append: GP not fitted");
    const int N = g.N, d = g.d;
    const int NB1 = (N + 8) / 8, Npad1 = 8 * NB1;
    if (Npad1 > g.capN || N + 1 > 2048) return so_fail(h, SO_ERR_CAPACITY, "so_fit_append: capacity exhausted, refit with so_fit");
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int ld = g.ld;
    g.fitted = false;
    g.grid_ready = false;
    g.f32_ready = false;
    g.tma_ready = false;
    if (NB1 != g.NB) {
        // a new block of 8 rows starts: its 7 other rows are padding (finite zeros), its fragment blocks start empty
        const int cells = 8 * (d + 3);
        k_clear_rows<<<(cells + 127) / 128, 128, 0, stream>>>(g.Xs, g.Y, g.alpha, g.zvec, N, Npad1, d);
        SO_CUDA(h, cudaMemsetAsync(g.Afrag + tri_blocks(g.NB) * 32, 0, sizeof(double2) * ((size_t)NB1 + 4) * 32, stream));
    }
    double xs[SO_MAX_DIM + 1];
    for (int j = 0; j < d; ++j) xs[j] = x_new_h[j] * g.inv_ls[j];
    xs[d] = y_new;
    // small synchronous-safe staging: cudaMemcpyAsync from pageable memory copies before returning
    SO_CUDA(h, cudaMemcpyAsync(g.X + (size_t)N * d, x_new_h, sizeof(double) * d, cudaMemcpyHostToDevice, stream));
    SO_CUDA(h, cudaMemcpyAsync(g.Xs + (size_t)N * d, xs, sizeof(double) * d, cudaMemcpyHostToDevice, stream));
    SO_CUDA(h, cudaMemcpyAsync(g.Y + N, xs + d, sizeof(double), cudaMemcpyHostToDevice, stream));
    SO_CUDA(h, cudaMemsetAsync(h->d_status, 0, sizeof(int), stream));
    SO_CUDA(h, cudaMemsetAsync(g.K + (size_t)N * ld, 0, sizeof(double) * Npad1, stream));
    SO_CUDA(h, cudaMemsetAsync(g.Linv + (size_t)N * ld, 0, sizeof(double) * Npad1, stream));
    k_append_row<<<(N + 7) / 8, 256, sizeof(double) * N, stream>>>(g.Linv, g.Xs, g.K + (size_t)N * ld, N, ld, d, g.kind, g.variance);
    const double kss = g.variance + g.noise + SO_JITTER;      // Kdiag = variance for every stationary family
    k_append_finish<<<(N + 1 + 255) / 256, 256, 0, stream>>>(g.K, g.Linv, N, ld, kss, h->d_status);
    k_append_alpha<<<1, 256, 0, stream>>>(g.Linv, g.Y, g.alpha, g.zvec, N, ld);
    k_pack_afrag<<<dim3(NB1, 1), 32, 0, stream>>>(g.Linv, g.Afrag, N + 1, ld, NB1 - 1);
    SO_CHECK_LAUNCH(h, "so_fit_append kernels");
    SO_CUDA(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SO_CUDA(h, cudaStreamSynchronize(stream));
    if (*h->h_status != SO_OK)
        return so_fail(h, SO_ERR_NOT_PD, "so_fit_append: bordered matrix is not positive definite; refit with so_fit");
    g.N = N + 1;
    g.NB = NB1;
    g.fitted = true;
    return SO_OK;
}

extern "C" int so_fit_remove_last(so_handle* h, int gp, void* stream_) {
    if (!h) return SO_ERR_BAD_ARG;
    if (gp < 0 || gp >= h->max_gps) return so_fail(h, SO_ERR_BAD_ARG, "so_fit_remove_last: gp index out of range");
    GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "so_fit_remove_last: GP not fitted");
    if (g.N < 2) return so_fail(h, SO_ERR_BAD_ARG, "so_fit_remove_last: at least one point must remain");
    DeviceGuard guard(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int N1 = g.N - 1, d = g.d, ld = g.ld;
    const int NB1 = (N1 + 7) / 8, Npad1 = 8 * NB1;
    g.grid_ready = false;
    g.f32_ready = false;
    g.tma_ready = false;
    // row N1 becomes padding; alpha / z are rebuilt from the leading block of L^-1 (exactly what a refit would use)
    k_clear_rows<<<(d + 3 + 127) / 128, 128, 0, stream>>>(g.Xs, g.Y, g.alpha, g.zvec, N1, N1 + 1, d);
    k_alpha<<<1, 1024, sizeof(double) * Npad1, stream>>>(g.Linv, g.Y, g.alpha, g.zvec, N1, Npad1, ld);
    k_pack_afrag<<<dim3(NB1, 1), 32, 0, stream>>>(g.Linv, g.Afrag, N1, ld, NB1 - 1);
    if (NB1 != g.NB)        // the dropped block row: the posterior kernels prefetch a few blocks past the end, keep them finite
        SO_CUDA(h, cudaMemsetAsync(g.Afrag + tri_blocks(NB1) * 32, 0, sizeof(double2) * ((size_t)g.NB + 4) * 32, stream));
    SO_CHECK_LAUNCH(h, "so_fit_remove_last kernels");
    g.N = N1;
    g.NB = NB1;
    return SO_OK;
}
