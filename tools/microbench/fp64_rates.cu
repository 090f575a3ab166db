// fp64 pipe calibration for B200 (sm_100a): DFMA vs DMMA.8x8x4 vs both, fp64 exp(), cuBLAS DGEMM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_rates fp64_rates.cu -lcublas
// The DGEMM rate printed here is the denominator for the fp64 roofline in bench.py.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved in the same warp (RATIO dfma per dmma)
template <int NACC, int RATIO>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC], f[NACC * RATIO];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NACC * RATIO; ++i) f[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            dmma884(c0[i], c1[i], a, b);
#pragma unroll
            for (int r = 0; r < RATIO; ++r) f[i * RATIO + r] = fma(f[i * RATIO + r], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NACC * RATIO; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Warp-specialised: even warps DMMA, odd warps DFMA
template <int NACC>
__global__ void __launch_bounds__(256) k_split(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    if ((threadIdx.x >> 5) & 4) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double a) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -1e-3 * (threadIdx.x + i + 1);
    double s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += exp(x[i]); x[i] *= a; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory fragments (LDS.128 of A and B per 2 k-steps), a x b register blocking
template <int RA, int RB>
__global__ void __launch_bounds__(256) k_dmma_smem(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double c0[RA * RB], c1[RA * RB];
#pragma unroll
    for (int i = 0; i < RA * RB; ++i) { c0[i] = 0; c1[i] = 0; }
    for (int it = 0; it < iters; ++it) {
        const double2* pa = reinterpret_cast<const double2*>(sm) + ((it * 7 + warp) & 63) * 32 + lane;
        const double2* pb = reinterpret_cast<const double2*>(sm) + 4096 + ((it * 5) & 63) * 32 + lane;
        double2 A[RA], B[RB];
#pragma unroll
        for (int i = 0; i < RA; ++i) A[i] = pa[i * 32];
#pragma unroll
        for (int j = 0; j < RB; ++j) B[j] = pb[j * 32];
#pragma unroll
        for (int i = 0; i < RA; ++i)
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                dmma884(c0[i * RB + j], c1[i * RB + j], A[i].x, B[j].x);
                dmma884(c0[i * RB + j], c1[i * RB + j], A[i].y, B[j].y);
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < RA * RB; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256 * 4));
    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        {
            float ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 16 * iters * 256.0 * grid;
            printf("{\"test\": \"dfma\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 256 * 16 * iters * 8.0 * grid;
            printf("{\"test\": \"dmma884\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9);
        }
    }
    {
        int grid = sms * 2;
        float ms = time_ms([&] { k_mixed<8, 1><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
        double fl_mma = 2.0 * 256 * 8 * iters * 8.0 * grid, fl_fma = 2.0 * 8 * iters * 256.0 * grid;
        printf("{\"test\": \"mixed_1dfma_per_dmma\", \"ms\": %.3f, \"tflops_dmma\": %.2f, \"tflops_dfma\": %.2f}\n", ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
        ms = time_ms([&] { k_mixed<8, 4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
        fl_fma *= 4;
        printf("{\"test\": \"mixed_4dfma_per_dmma\", \"ms\": %.3f, \"tflops_dmma\": %.2f, \"tflops_dfma\": %.2f}\n", ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
        ms = time_ms([&] { k_split<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
        double fm = 2.0 * 256 * 16 * iters * 4.0 * grid, ff = 2.0 * 32 * iters * 128.0 * grid;
        printf("{\"test\": \"split_warps_half_dmma_half_dfma\", \"ms\": %.3f, \"tflops_dmma\": %.2f, \"tflops_dfma\": %.2f}\n", ms, fm / ms * 1e-9, ff / ms * 1e-9);
    }
    {
        int grid = sms * 4;
        float ms = time_ms([&] { k_exp<<<grid, 256>>>(out, 2000, 0.9999999); });
        double n = 8.0 * 2000 * 256.0 * grid;
        printf("{\"test\": \"exp_f64\", \"ms\": %.3f, \"gexp_per_s\": %.2f}\n", ms, n / ms * 1e-6);
    }
    {
        CK(cudaFuncSetAttribute(k_dmma_smem<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        CK(cudaFuncSetAttribute(k_dmma_smem<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        CK(cudaFuncSetAttribute(k_dmma_smem<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        CK(cudaFuncSetAttribute(k_dmma_smem<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        int it2 = 4000;
        float ms = time_ms([&] { k_dmma_smem<4, 4><<<sms, 256, 16384 * 8>>>(out, it2); });
        printf("{\"test\": \"dmma_smem_4x4_1cta\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 256 * 32 * it2 * 8.0 * sms / ms * 1e-9);
        ms = time_ms([&] { k_dmma_smem<4, 2><<<sms, 256, 16384 * 8>>>(out, it2); });
        printf("{\"test\": \"dmma_smem_4x2_1cta\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 256 * 16 * it2 * 8.0 * sms / ms * 1e-9);
        ms = time_ms([&] { k_dmma_smem<2, 2><<<sms, 256, 16384 * 8>>>(out, it2); });
        printf("{\"test\": \"dmma_smem_2x2_1cta\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 256 * 8 * it2 * 8.0 * sms / ms * 1e-9);
        ms = time_ms([&] { k_dmma_smem<4, 8><<<sms, 256, 16384 * 8>>>(out, it2); });
        printf("{\"test\": \"dmma_smem_4x8_1cta\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 256 * 64 * it2 * 8.0 * sms / ms * 1e-9);
    }
    {
        cublasHandle_t h; cublasCreate(&h);
        for (int n : {2048, 4096, 8192}) {
            double *A, *B, *C;
            CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
            CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
            double one = 1.0, zero = 0.0;
            float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 4);
            printf("{\"test\": \"cublas_dgemm\", \"n\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", n, ms, 2.0 * n * n * (double)n / ms * 1e-9);
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        cublasDestroy(h);
    }
    return 0;
}
