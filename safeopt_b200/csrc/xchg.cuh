// Cross-rank record exchange over peer-mapped memory (NVLink / NVSwitch), used INSIDE the set-pass and swarm kernels
// instead of an NCCL all-gather between kernels (the only data that crosses GPUs are 64..144-byte records).
//
// Every rank owns one XchgBuf (cudaMalloc, exported with cudaIpcGetMemHandle, opened by the peers with
// cudaIpcOpenMemHandle; so_xchg_export / so_xchg_connect).  A producer kernel writes its record into slot [rank] of EVERY
// rank's buffer (plain stores through the peer mapping), then a system-scope release store of the epoch stamp into the
// matching flag; a consumer spins with system-scope acquire loads on the flags in its OWN buffer until all `world` stamps
// carry the current epoch.  Epochs are counted on the device (one counter per protocol, bumped by the kernel itself) so a
// CUDA graph can replay the kernels; slots are double-buffered by epoch parity so that a rank running ahead never overwrites
// a record a slower rank has not read yet (a slot is rewritten two epochs later, and a rank only reaches epoch e+1 after every
// rank has published epoch e, i.e. after every rank finished reading epoch e-1).
#pragma once
#include "common.cuh"

constexpr int kXchgMaxWorld = 16;

struct XchgSets {                                   // one parity of the set-pass protocol
    so_safe_record safe[kXchgMaxWorld];
    so_max_record max[kXchgMaxWorld];
    long long ncand[kXchgMaxWorld];
    unsigned long long flag[3][kXchgMaxWorld];      // epoch stamps of the three phases
};

struct XchgSwarm {                                  // one parity of the swarm best-record protocol
    double rec[kXchgMaxWorld][SO_SWARM_REC_DOUBLES];
    unsigned long long flag[kXchgMaxWorld];
};

struct XchgBuf {
    XchgSets sets[2];
    XchgSwarm swarm[2];
};

struct XchgView {                                   // kernel parameter: where the ranks' buffers are mapped in this process
    XchgBuf* local;
    XchgBuf* peer[kXchgMaxWorld];                   // peer[rank] == local
    int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Spin until flag[r] >= epoch (threads r < world poll one flag each; the caller synchronises the block afterwards).  With one
// rank the flag was written by this GPU: device scope suffices.  A peer that never arrives (crashed rank) must not hang the
// GPU: after ~kXchgTimeoutCycles of this SM's clock the wait gives up and returns false; the kernel then stamps an error into
// its status word and the host raises.  (The SM cycle counter is used, not %globaltimer, whose reads cost microseconds.)
constexpr long long kXchgTimeoutCycles = 40ll * 1000ll * 1000ll * 1000ll;       // ~20 s at 2 GHz
__device__ __forceinline__ bool xchg_wait(const unsigned long long* flags, int r, unsigned long long epoch, bool multi_gpu) {
    long long t0 = 0;
    unsigned spins = 0;
    while ((multi_gpu ? ld_acquire_sys(flags + r) : ld_acquire_gpu(flags + r)) < epoch) {
        if ((++spins & 0xfffu) == 0) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > kXchgTimeoutCycles) return false;
        }
    }
    return true;
}
