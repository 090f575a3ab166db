// Host-side accuracy check of so_exp_neg (same source as the device function).  nvcc -O2 -o /tmp/t tools/test_fastexp.cu && /tmp/t
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../safeopt_b200/csrc/fastexp.cuh"
static double ulp_err(double got, double ref) {
    if (ref == 0.0) return got == 0.0 ? 0.0 : 1e9;
    int e; frexp(ref, &e);
    return fabs(got - ref) / ldexp(1.0, e - 53);
}
__global__ void k_dev(const double* x, double* y, double* yl, int n) {
    __shared__ double T[64];
    const double tab[64] = {SO_EXP_TABLE_VALUES};
    if (threadIdx.x < 64) T[threadIdx.x] = tab[threadIdx.x];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { y[i] = so_exp_neg(x[i], T); yl[i] = exp(x[i]); }
}

static int device_check() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no GPU: device check skipped\n"); return 0; }
    const int n = 1 << 20;
    double *hx = new double[n], *hy = new double[n], *hl = new double[n], *dx, *dy, *dl;
    srand(7);
    for (int i = 0; i < n; ++i) { double u = rand() / (double)RAND_MAX; hx[i] = (i & 1) ? -40.0 * u * u : -708.0 * u; }
    hx[0] = 0.0; hx[1] = -0.0; hx[2] = -1e-300; hx[3] = -0.5 * 0.02020202020202022 * 0.02020202020202022;
    cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&dl, n * 8);
    cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
    k_dev<<<n / 256, 256>>>(dx, dy, dl, n);
    cudaMemcpy(hy, dy, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hl, dl, n * 8, cudaMemcpyDeviceToHost);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("device error\n"); return 1; }
    double w = 0, wx = 0, wl = 0;
    for (int i = 0; i < n; ++i) {
        if (hx[i] <= -700.0) continue;
        double e = ulp_err(hy[i], exp(hx[i])), el = ulp_err(hl[i], exp(hx[i]));
        if (e > w) { w = e; wx = hx[i]; }
        if (el > wl) wl = el;
    }
    printf("device: so_exp_neg max error %.3f ulp at x=%.17g; libdevice exp max error %.3f ulp\n", w, wx, wl);
    return w <= 1.0 ? 0 : 1;
}

int main() {
    const double T[64] = {SO_EXP_TABLE_VALUES};
    double worst = 0, worst_x = 0;
    srand(1);
    long n = 0;
    for (int rep = 0; rep < 4000000; ++rep) {
        double u = rand() / (double)RAND_MAX;
        double x;
        switch (rep & 3) {
            case 0: x = -708.0 * u; break;
            case 1: x = -50.0 * u * u; break;
            case 2: x = -u * 1e-3; break;
            default: x = -12.5 * u; break;
        }
        double g = so_exp_neg(x, T), r = exp(x);
        if (x <= -700.0) continue;
        double e = ulp_err(g, r);
        if (e > worst) { worst = e; worst_x = x; }
        ++n;
    }
    printf("points %ld  max error %.3f ulp at x=%.17g  (exp_neg(0)=%.17g, exp_neg(-1e-300)=%.17g, exp_neg(-800)=%g)\n", n, worst, worst_x,
           so_exp_neg(0.0, T), so_exp_neg(-1e-300, T), so_exp_neg(-800.0, T));
    if (worst > 1.0) return 1;
    return device_check();
}
