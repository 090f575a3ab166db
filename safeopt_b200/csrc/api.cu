// Handle lifetime and error reporting for the C ABI (include/safeopt_b200.h).
#include "common.cuh"

void xchg_destroy(so_handle* h);

extern "C" {

int so_abi_version(void) { return SO_ABI_VERSION; }

const char* so_status_string(int status) {
    switch (status) {
        case SO_OK: return "ok";
        case SO_ERR_BAD_ARG: return "bad argument";
        case SO_ERR_UNSUPPORTED: return "unsupported kernel family or shape";
        case SO_ERR_NOT_PD: return "covariance matrix not positive definite";
        case SO_ERR_CUDA: return "CUDA error";
        case SO_ERR_NOT_FITTED: return "GP not fitted";
        case SO_ERR_CAPACITY: return "capacity exceeded";
        case SO_ERR_NO_DEVICE: return "no CUDA device";
        case SO_ERR_TIMEOUT: return "cross-rank exchange timed out";
        default: return "unknown status";
    }
}

int so_create(int device, int max_gps, so_handle** out) {
    if (!out || max_gps < 1 || max_gps > 64) return SO_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SO_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return SO_ERR_BAD_ARG;
    so_handle* h = new so_handle();
    h->device = device;
    h->max_gps = max_gps;
    h->gps.resize(max_gps);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return SO_ERR_CUDA; }
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaMalloc(&h->d_status, sizeof(int)) != cudaSuccess ||
        cudaHostAlloc(&h->h_status, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&h->status_mapped_d, h->h_status, 0) != cudaSuccess ||
        cudaMallocHost(&h->fit_stage_h, (size_t)max_gps * 512 * (SO_MAX_DIM + 1) * sizeof(double)) != cudaSuccess ||
        cudaHostAlloc(&h->fit_status_h, (size_t)max_gps * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&h->fit_status_d, h->fit_status_h, 0) != cudaSuccess ||
        cudaMalloc(&h->ws_partials, (size_t)SO_WS_MAX_BLOCKS * 64) != cudaSuccess ||
        cudaMalloc(&h->ws_counter, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(h->ws_counter, 0, sizeof(unsigned int)) != cudaSuccess) {
        delete h;
        return SO_ERR_CUDA;
    }
    h->fit_stage_bytes = (size_t)512 * (SO_MAX_DIM + 1) * sizeof(double);
    for (int i = 0; i < max_gps; ++i) h->fit_status_h[i] = SO_OK;
    *out = h;
    return SO_OK;
}

static void free_gp(GPState& g) {
    cudaFree(g.X); cudaFree(g.Xs); cudaFree(g.Y); cudaFree(g.K); cudaFree(g.Linv);
    cudaFree(g.alpha); cudaFree(g.zvec); cudaFree(g.Afrag); cudaFree(g.E); cudaFree(g.P2); cudaFree(g.PfFrag); cudaFree(g.Aprime); cudaFree(g.f32_A); cudaFree(g.f32_B); cudaFree(g.f32_PfT);
    g = GPState();
}

int so_destroy(so_handle* h) {
    if (!h) return SO_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    for (auto& g : h->gps) free_gp(g);
    cudaFree(h->grid.axis);
    cudaFree(h->d_status);
    cudaFreeHost(h->h_status);
    cudaFreeHost(h->fit_stage_h);
    cudaFreeHost(h->fit_status_h);
    cudaFree(h->ws_partials);
    cudaFree(h->ws_counter);
    cudaFree(h->ws_z);
    cudaFree(h->f32_mean_scratch);
    xchg_destroy(h);
    cudaFree(h->fused_bar); cudaFree(h->fused_part); cudaFree(h->fused_ncand); cudaFree(h->fused_dbg);
    cudaFreeHost(h->fused_result_h);
    delete h;
    return SO_OK;
}

const char* so_last_error(const so_handle* h) { return h ? h->err.c_str() : "null handle"; }

int so_num_sms(const so_handle* h) { return h ? h->num_sms : 0; }

}  // extern "C"
