// Host side of the cross-rank record exchange (see xchg.cuh): allocation of this rank's buffer, its CUDA IPC handle,
// and the mapping of the peers' buffers.  The handles travel between the processes through whatever the host runtime
// uses for plumbing (torch.distributed all_gather in safeopt_b200/distributed.py); no data-path bytes go that way.
#include "xchg.cuh"
#include <cstring>

static_assert(sizeof(cudaIpcMemHandle_t) == SO_XCHG_HANDLE_BYTES, "IPC handle size is part of the ABI");

int xchg_ensure_local(so_handle* h) {
    if (h->xchg_local) return SO_OK;
    DeviceGuard guard(h->device);
    XchgBuf* buf = nullptr;
    SO_CUDA(h, cudaMalloc(&buf, sizeof(XchgBuf)));
    SO_CUDA(h, cudaMemset(buf, 0, sizeof(XchgBuf)));
    SO_CUDA(h, cudaMalloc(&h->xchg_epochs, 8 * sizeof(unsigned long long)));
    SO_CUDA(h, cudaMemset(h->xchg_epochs, 0, 8 * sizeof(unsigned long long)));
    h->xchg_local = buf;
    h->xchg_world = 1;
    h->xchg_rank = 0;
    h->xchg_peer[0] = buf;
    return SO_OK;
}

XchgView xchg_view(const so_handle* h) {
    XchgView v;
    v.local = h->xchg_local;
    for (int r = 0; r < kXchgMaxWorld; ++r) v.peer[r] = r < h->xchg_world ? h->xchg_peer[r] : nullptr;
    v.world = h->xchg_world;
    v.rank = h->xchg_rank;
    return v;
}

static void xchg_close_peers(so_handle* h) {
    for (int r = 0; r < kXchgMaxWorld; ++r) {
        if (h->xchg_opened[r] && h->xchg_peer[r]) cudaIpcCloseMemHandle(h->xchg_peer[r]);
        h->xchg_opened[r] = false;
        h->xchg_peer[r] = nullptr;
    }
}

void xchg_destroy(so_handle* h) {
    xchg_close_peers(h);
    cudaFree(h->xchg_local);
    cudaFree(h->xchg_epochs);
    h->xchg_local = nullptr;
    h->xchg_epochs = nullptr;
}

extern "C" int so_xchg_export(so_handle* h, void* ipc_handle_h) {
    if (!h || !ipc_handle_h) return SO_ERR_BAD_ARG;
    int rc = xchg_ensure_local(h);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    cudaIpcMemHandle_t hd;
    SO_CUDA(h, cudaIpcGetMemHandle(&hd, h->xchg_local));
    std::memcpy(ipc_handle_h, &hd, sizeof(hd));
    return SO_OK;
}

extern "C" int so_xchg_connect(so_handle* h, int world, int rank, const void* ipc_handles_h) {
    if (!h || !ipc_handles_h) return SO_ERR_BAD_ARG;
    if (world < 1 || world > kXchgMaxWorld || rank < 0 || rank >= world)
        return so_fail(h, SO_ERR_BAD_ARG, "so_xchg_connect: 1 <= world <= 16, 0 <= rank < world");
    int rc = xchg_ensure_local(h);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    xchg_close_peers(h);
    const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(ipc_handles_h);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { h->xchg_peer[r] = h->xchg_local; continue; }
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            xchg_close_peers(h);
            h->xchg_peer[0] = h->xchg_local; h->xchg_world = 1; h->xchg_rank = 0;
            cudaGetLastError();
            return so_fail(h, SO_ERR_CUDA, std::string("so_xchg_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        }
        h->xchg_peer[r] = static_cast<XchgBuf*>(p);
        h->xchg_opened[r] = true;
    }
    h->xchg_world = world;
    h->xchg_rank = rank;
    return SO_OK;
}

extern "C" int so_xchg_world(const so_handle* h) { return h ? h->xchg_world : 0; }
