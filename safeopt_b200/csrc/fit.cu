// K1 -- training-side fit on the device, fp64 throughout.
// Stands in for GPy's ExactGaussianInference.inference reached from `gp.set_XY`
// (reference call sites safeopt/gp_opt.py:227, :267, :275): Ky = K(X,X) + (noise+1e-8) I,
// lower Cholesky, L^-1, alpha = Ky^-1 Y.  The explicit triangular inverse (not GPy's dpotri
// full inverse) is what the posterior contraction consumes: var = k** - |L^-1 k|^2 costs
// N^2/2 FMAs per candidate instead of N^2 and is the better-conditioned form (SURVEY.md
// section 7, hard part 1).
#include "common.cuh"

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
__global__ void k_build_ky(const double* __restrict__ Xs, double* __restrict__ K, int N, int Npad, int d,
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
    K[(size_t)i * Npad + j] = v;
}

// In-place right-looking Cholesky of the leading N x N block (lower triangle), one CTA.
__global__ void __launch_bounds__(1024) k_cholesky(double* __restrict__ K, int N, int ld, int* status) {
    extern __shared__ double col[];
    __shared__ int bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int j = 0; j < N; ++j) {
        if (tid == 0) {
            double dj = K[(size_t)j * ld + j];
            if (!(dj > 0.0) || isinf(dj)) bad = 1; else K[(size_t)j * ld + j] = sqrt(dj);
        }
        __syncthreads();
        if (bad) break;
        const double ljj = K[(size_t)j * ld + j];
        for (int i = j + 1 + tid; i < N; i += nt) {
            double v = K[(size_t)i * ld + j] / ljj;
            K[(size_t)i * ld + j] = v;
            col[i] = v;
        }
        __syncthreads();
        for (int i = j + 1 + warp; i < N; i += nwarps) {
            const double ci = col[i];
            double* row = K + (size_t)i * ld;
            for (int k = j + 1 + lane; k <= i; k += 32) row[k] = fma(-ci, col[k], row[k]);
        }
        __syncthreads();
    }
    if (tid == 0) *status = bad ? SO_ERR_NOT_PD : SO_OK;
}

// L^-1 by forward substitution, one warp per column (columns are independent).
__global__ void __launch_bounds__(256) k_trinv(const double* __restrict__ L, double* __restrict__ Linv, int N, int ld) {
    extern __shared__ double xs_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= N) return;
    double* xs = xs_all + (size_t)warp * ld;
    for (int i = c; i < N; ++i) {
        const double* row = L + (size_t)i * ld;
        double part = 0.0;
        for (int k = c + lane; k < i; k += 32) part = fma(row[k], xs[k], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        double xi = ((i == c ? 1.0 : 0.0) - part) / row[i];
        if (lane == 0) xs[i] = xi;
        __syncwarp();
    }
    for (int i = c + lane; i < N; i += 32) Linv[(size_t)i * ld + c] = xs[i];
}

// alpha = Linv^T (Linv y), one CTA; padding entries are zero.
__global__ void __launch_bounds__(1024) k_alpha(const double* __restrict__ Linv, const double* __restrict__ Y,
                                                double* __restrict__ alpha, double* __restrict__ zvec, int N, int Npad) {
    extern __shared__ double w[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = warp; i < N; i += nwarps) {
        const double* row = Linv + (size_t)i * Npad;
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
            for (int i = c; i < N; ++i) s = fma(Linv[(size_t)i * Npad + c], w[i], s);
        alpha[c] = s;
        zvec[c] = c < N ? w[c] : 0.0;
    }
}

// Pack L^-1 into DMMA A-fragment order.  Block (i, kb), kb <= i, lives at index i(i+1)/2 + kb and
// holds 32 double2: lane l -> (Linv[8i + l/4][8kb + 2(l%4)], Linv[8i + l/4][8kb + 2(l%4) + 1]).
// The k permutation inside an 8-block (even slots in MMA step 0, odd slots in step 1) is shared
// with the B operand built by the posterior kernel, so a lane's two values are adjacent doubles.
__global__ void k_pack_afrag(const double* __restrict__ Linv, double2* __restrict__ Afrag, int N, int Npad) {
    const int i = blockIdx.y, kb = blockIdx.x;
    if (kb > i) return;
    const int lane = threadIdx.x;
    const int r = 8 * i + (lane >> 2), c0 = 8 * kb + 2 * (lane & 3);
    double v0 = 0.0, v1 = 0.0;
    if (r < N) {
        if (c0 < N && c0 <= r) v0 = Linv[(size_t)r * Npad + c0];
        if (c0 + 1 < N && c0 + 1 <= r) v1 = Linv[(size_t)r * Npad + c0 + 1];
    }
    Afrag[((size_t)i * (i + 1) / 2 + kb) * 32 + lane] = make_double2(v0, v1);
}

template <typename T>
int grow(so_handle* h, T*& ptr, size_t count) {
    if (ptr) { cudaFree(ptr); ptr = nullptr; }
    SO_CUDA(h, cudaMalloc(&ptr, count * sizeof(T)));
    return SO_OK;
}

}  // namespace

extern "C" int so_fit(so_handle* h, int gp, const double* X_h, const double* Y_h, int N, int d, int kernel_kind,
                      const double* lengthscale_h, double variance, double noise_var, void* stream_) {
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
    }
    g.fitted = false;
    g.grid_ready = false;
    g.N = N; g.d = d; g.kind = kernel_kind; g.NB = NB;
    g.variance = variance; g.noise = noise_var;
    InvLs il;
    for (int j = 0; j < SO_MAX_DIM; ++j) { il.v[j] = j < d ? 1.0 / lengthscale_h[j] : 0.0; g.inv_ls[j] = il.v[j]; }

    SO_CUDA(h, cudaMemcpyAsync(g.X, X_h, sizeof(double) * N * d, cudaMemcpyHostToDevice, stream));
    SO_CUDA(h, cudaMemcpyAsync(g.Y, Y_h, sizeof(double) * N, cudaMemcpyHostToDevice, stream));
    k_scale_x<<<(Npad * d + 255) / 256, 256, 0, stream>>>(g.X, g.Xs, N, Npad, d, il);
    dim3 blk(16, 16), grd((Npad + 15) / 16, (Npad + 15) / 16);
    k_build_ky<<<grd, blk, 0, stream>>>(g.Xs, g.K, N, Npad, d, kernel_kind, variance, noise_var + SO_JITTER);
    k_cholesky<<<1, 1024, sizeof(double) * Npad, stream>>>(g.K, N, Npad, h->d_status);
    SO_CUDA(h, cudaMemsetAsync(g.Linv, 0, sizeof(double) * (size_t)Npad * Npad, stream));
    {
        const int warps = 8;
        size_t smem = sizeof(double) * (size_t)warps * Npad;
        static bool attr_set = false;
        if (!attr_set) {
            SO_CUDA(h, cudaFuncSetAttribute(k_trinv, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        k_trinv<<<(N + warps - 1) / warps, warps * 32, smem, stream>>>(g.K, g.Linv, N, Npad);
    }
    k_alpha<<<1, 1024, sizeof(double) * Npad, stream>>>(g.Linv, g.Y, g.alpha, g.zvec, N, Npad);
    SO_CUDA(h, cudaMemsetAsync(g.Afrag, 0, sizeof(double2) * (tri_blocks(NB) + 4) * 32, stream));
    k_pack_afrag<<<dim3(NB, NB), 32, 0, stream>>>(g.Linv, g.Afrag, N, Npad);
    SO_CHECK_LAUNCH(h, "so_fit kernels");
    SO_CUDA(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SO_CUDA(h, cudaStreamSynchronize(stream));
    if (*h->h_status != SO_OK)
        return so_fail(h, SO_ERR_NOT_PD, "so_fit: K + (noise + 1e-8) I is not positive definite");
    g.fitted = true;
    return SO_OK;
}

extern "C" int so_fit_export(so_handle* h, int gp, double* L_h, double* Linv_h, double* alpha_h) {
    if (!h || gp < 0 || gp >= h->max_gps) return SO_ERR_BAD_ARG;
    GPState& g = h->gps[gp];
    if (!g.fitted) return so_fail(h, SO_ERR_NOT_FITTED, "so_fit_export: GP not fitted");
    DeviceGuard guard(h->device);
    SO_CUDA(h, cudaDeviceSynchronize());
    const int N = g.N, Npad = 8 * g.NB;
    if (L_h) {
        SO_CUDA(h, cudaMemcpy2D(L_h, sizeof(double) * N, g.K, sizeof(double) * Npad, sizeof(double) * N, N, cudaMemcpyDeviceToHost));
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) L_h[(size_t)i * N + j] = 0.0;
    }
    if (Linv_h)
        SO_CUDA(h, cudaMemcpy2D(Linv_h, sizeof(double) * N, g.Linv, sizeof(double) * Npad, sizeof(double) * N, N, cudaMemcpyDeviceToHost));
    if (alpha_h) SO_CUDA(h, cudaMemcpy(alpha_h, g.alpha, sizeof(double) * N, cudaMemcpyDeviceToHost));
    return SO_OK;
}
