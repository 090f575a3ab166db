// K3 -- the set logic of SafeOpt.compute_sets as HBM-streaming passes over Q (fp64, row-major,
// 2G columns) and the byte masks.  Each pass is one coalesced sweep: a thread takes 16 consecutive rows,
// reads their 16 mask bytes with one 128-bit load and touches Q only for rows that are in the set
// (16*G + 1..3 bytes per row at most); warp-shuffle + shared-memory block reduction, and the last block
// to finish combines the per-block partials in parallel, so results are deterministic (first-row
// tie-breaks follow NumPy's argmax, gp_opt.py:635,:644,:710).
//   so_sets_reduce_safe : gp_opt.py:504 (any S), :512 (max l0[S]), :634-636, :708-712
//   so_sets_maximizers  : gp_opt.py:511-513 and the M part of :642-644
//   so_sets_candidates  : gp_opt.py:531-536 and the sort key of :551
// The *_chain variants take the scalar that links two passes (max l0[S]; max width over M) from the
// records of the previous pass in device memory -- one record per rank, as all-gathered -- so the three
// passes run back to back on the stream and the host reads all results with a single copy.
#include "xchg.cuh"
#include <cstdlib>
#include <cstring>

int xchg_ensure_local(so_handle* h);
XchgView xchg_view(const so_handle* h);

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = SO_WS_MAX_BLOCKS;
constexpr int kRows = 16;                      // rows per thread and step: one uint4 of mask bytes

struct SafePartial { long long n; double max_l; long long arg_l; double max_u; long long arg_u; };
struct MaxPartial { long long n; double max_w0; double best; long long best_row; };
struct Scal { double v[64]; };

__device__ __forceinline__ void take_max_first(double& v, long long& r, double ov, long long orow) {
    // keep the larger value; on ties keep the smaller (global) row; row < 0 means "empty"
    if (orow >= 0 && (r < 0 || ov > v || (ov == v && orow < r))) { v = ov; r = orow; }
}

__device__ __forceinline__ void merge(SafePartial& a, const SafePartial& o) {
    a.n += o.n;
    take_max_first(a.max_l, a.arg_l, o.max_l, o.arg_l);
    take_max_first(a.max_u, a.arg_u, o.max_u, o.arg_u);
}
__device__ __forceinline__ void merge(MaxPartial& a, const MaxPartial& o) {
    a.n += o.n;
    a.max_w0 = o.max_w0 > a.max_w0 ? o.max_w0 : a.max_w0;
    take_max_first(a.best, a.best_row, o.best, o.best_row);
}
__device__ __forceinline__ SafePartial shfl_xor(const SafePartial& a, int o) {
    SafePartial r;
    r.n = __shfl_xor_sync(0xffffffffu, a.n, o);
    r.max_l = __shfl_xor_sync(0xffffffffu, a.max_l, o);
    r.arg_l = __shfl_xor_sync(0xffffffffu, a.arg_l, o);
    r.max_u = __shfl_xor_sync(0xffffffffu, a.max_u, o);
    r.arg_u = __shfl_xor_sync(0xffffffffu, a.arg_u, o);
    return r;
}
__device__ __forceinline__ MaxPartial shfl_xor(const MaxPartial& a, int o) {
    MaxPartial r;
    r.n = __shfl_xor_sync(0xffffffffu, a.n, o);
    r.max_w0 = __shfl_xor_sync(0xffffffffu, a.max_w0, o);
    r.best = __shfl_xor_sync(0xffffffffu, a.best, o);
    r.best_row = __shfl_xor_sync(0xffffffffu, a.best_row, o);
    return r;
}

// Block reduction of `acc` (all operations are order-independent: integer sums, max with lowest-row ties);
// the result is valid in thread 0.
template <typename P>
__device__ __forceinline__ P block_reduce(P acc, P* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) merge(acc, shfl_xor(acc, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned nw = blockDim.x >> 5;                  // 8 warps (one-pass kernels, fused) or 32 (single-block fused)
        P t = sm[threadIdx.x < nw ? threadIdx.x : 0];
        if (threadIdx.x >= nw) t.n = 0;                       // duplicates of entry 0 do not change a max; their counts must not be added twice
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) merge(t, shfl_xor(t, o));
        acc = t;
    }
    __syncthreads();
    return acc;
}

// Publishes this block's partial; the last block to arrive reduces all partials (in parallel) and returns true
// in thread 0 with the total in `acc`.
template <typename P>
__device__ __forceinline__ bool grid_reduce(P& acc, P* __restrict__ part, unsigned int* __restrict__ counter, P identity) {
    __shared__ P sm[32];
    __shared__ bool last;
    acc = block_reduce(acc, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = acc;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    P t = identity;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) merge(t, part[b]);
    acc = block_reduce(t, sm);
    if (threadIdx.x == 0) *counter = 0;
    return threadIdx.x == 0;
}

// 16 mask bytes of rows [r0, r0+16): one 128-bit load when the block is whole and aligned, bytes otherwise
// (rows >= M read as 0).
__device__ __forceinline__ void load_mask16(const uint8_t* __restrict__ m, int64_t r0, int64_t M, bool aligned, uint8_t (&b)[kRows]) {
    if (aligned && r0 + kRows <= M) {
        const uint4 v = *reinterpret_cast<const uint4*>(m + r0);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < kRows; ++k) b[k] = (uint8_t)((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
    } else {
#pragma unroll
        for (int k = 0; k < kRows; ++k) b[k] = (r0 + k < M) ? m[r0 + k] : (uint8_t)0;
    }
}
__device__ __forceinline__ bool any16(const uint8_t (&b)[kRows]) {
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < kRows; ++k) acc |= b[k];
    return acc != 0;
}
__device__ __forceinline__ void store_mask16(uint8_t* __restrict__ m, int64_t r0, int64_t M, bool aligned, const uint8_t (&b)[kRows]) {
    if (aligned && r0 + kRows <= M) {
        unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < kRows; ++k) w[k >> 2] |= (unsigned)b[k] << (8 * (k & 3));
        *reinterpret_cast<uint4*>(m + r0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
#pragma unroll
        for (int k = 0; k < kRows; ++k)
            if (r0 + k < M) m[r0 + k] = b[k];
    }
}

// The three streaming scans, shared by the one-pass kernels and the fused kernel.
__device__ __forceinline__ SafePartial scan_safe(const double* __restrict__ Q, int q_stride, int64_t M, int64_t row0,
                                                 const uint8_t* __restrict__ S) {
    SafePartial acc = {0, -INFINITY, -1, -INFINITY, -1};
    const bool aligned = (reinterpret_cast<uintptr_t>(S) & 15) == 0;
    const int64_t nchunks = (M + kRows - 1) / kRows;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r0 = c * kRows;
        uint8_t s[kRows];
        load_mask16(S, r0, M, aligned, s);
        if (!any16(s)) continue;
        // all loads of the chunk are issued before the first use: one memory latency per chunk instead of one per safe row
        double2 lu[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k)
            lu[k] = s[k] ? *reinterpret_cast<const double2*>(Q + (size_t)(r0 + k) * q_stride) : make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
            if (!s[k]) continue;
            const int64_t r = r0 + k;
            acc.n += 1;
            take_max_first(acc.max_l, acc.arg_l, lu[k].x, row0 + r);
            take_max_first(acc.max_u, acc.arg_u, lu[k].y, row0 + r);
        }
    }
    return acc;
}

__device__ __forceinline__ MaxPartial scan_maximizers(const double* __restrict__ Q, int G, int64_t M, int64_t row0,
                                                      const uint8_t* __restrict__ S, double max_l0, const Scal& scaling,
                                                      uint8_t* __restrict__ Mmask) {
    MaxPartial acc = {0, -INFINITY, -INFINITY, -1};
    const int qs = 2 * G;
    const bool aligned = ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(Mmask)) & 15) == 0;
    const int64_t nchunks = (M + kRows - 1) / kRows;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r0 = c * kRows;
        uint8_t s[kRows], m[kRows];
        load_mask16(S, r0, M, aligned, s);
#pragma unroll
        for (int k = 0; k < kRows; ++k) m[k] = 0;
        if (any16(s)) {
            double2 lus[kRows];
#pragma unroll
            for (int k = 0; k < kRows; ++k)
                lus[k] = s[k] ? *reinterpret_cast<const double2*>(Q + (size_t)(r0 + k) * qs) : make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                if (!s[k]) continue;
                const int64_t r = r0 + k;
                const double* q = Q + (size_t)r * qs;
                const double2 lu = lus[k];
                if (lu.y >= max_l0) {
                    m[k] = 1;
                    acc.n += 1;
                    const double w0 = lu.y - lu.x;
                    acc.max_w0 = w0 > acc.max_w0 ? w0 : acc.max_w0;
                    double val = w0 / scaling.v[0];
                    for (int i = 1; i < G; ++i) {
                        const double wi = (q[2 * i + 1] - q[2 * i]) / scaling.v[i];
                        val = wi > val ? wi : val;
                    }
                    take_max_first(acc.best, acc.best_row, val, row0 + r);
                }
            }
        }
        store_mask16(Mmask, r0, M, aligned, m);
    }
    return acc;
}

__device__ __forceinline__ void scan_candidates(const double* __restrict__ Q, int G, int64_t M, int64_t row0,
                                                const uint8_t* __restrict__ S, const uint8_t* __restrict__ Mmask, double max_var,
                                                const Scal& scaling, const Scal& thr, uint8_t* __restrict__ cmask,
                                                double* __restrict__ ckey, int64_t* __restrict__ crow, int64_t cap,
                                                unsigned long long* __restrict__ n_cand) {
    const int qs = 2 * G;
    const bool aligned = ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(Mmask) | reinterpret_cast<uintptr_t>(cmask)) & 15) == 0;
    const int64_t nchunks = (M + kRows - 1) / kRows;
    const unsigned lane = threadIdx.x & 31;
    // whole warps iterate together (the append below uses full-mask warp collectives)
    for (int64_t cbase = (int64_t)blockIdx.x * blockDim.x; cbase < nchunks; cbase += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = cbase + threadIdx.x;
        const int64_t r0 = c * kRows;
        uint8_t s[kRows], m[kRows], cnd[kRows];
        unsigned bits = 0;
#pragma unroll
        for (int k = 0; k < kRows; ++k) cnd[k] = 0;
        if (c < nchunks) {
            load_mask16(S, r0, M, aligned, s);
            if (any16(s)) {
                load_mask16(Mmask, r0, M, aligned, m);
#pragma unroll
                for (int k = 0; k < kRows; ++k) {
                    if (!s[k] || m[k]) continue;
                    const double* q = Q + (size_t)(r0 + k) * qs;
                    double smax = -INFINITY;
                    bool over = false;
                    for (int i = 0; i < G; ++i) {
                        const double w = q[2 * i + 1] - q[2 * i];
                        const double ws = w / scaling.v[i];
                        smax = ws > smax ? ws : smax;
                        over = over || (w > thr.v[i]);
                    }
                    if (smax > max_var && over) { cnd[k] = 1; bits |= 1u << k; }
                }
            }
            if (cmask) store_mask16(cmask, r0, M, aligned, cnd);
        }
        // warp-aggregated append: exclusive scan of the per-lane counts, one atomic per warp
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long slot0 = 0;
        if (lane == 31) slot0 = atomicAdd(n_cand, (unsigned long long)total);
        slot0 = __shfl_sync(0xffffffffu, slot0, 31);
        unsigned long long slot = slot0 + (unsigned long long)(incl - cnt);
        while (bits) {
            const int k = __ffs(bits) - 1;
            bits &= bits - 1;
            if ((int64_t)slot < cap) {
                const double* q = Q + (size_t)(r0 + k) * qs;
                double wmax = -INFINITY;                  // sort key of gp_opt.py:551: unscaled width
                for (int i = 0; i < G; ++i) {
                    const double w = q[2 * i + 1] - q[2 * i];
                    wmax = w > wmax ? w : wmax;
                }
                ckey[slot] = wmax;
                crow[slot] = row0 + r0 + k;
            }
            ++slot;
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_reduce_safe(const double* __restrict__ Q, int q_stride, int64_t M, int64_t row0,
                                                         const uint8_t* __restrict__ S, SafePartial* __restrict__ part,
                                                         unsigned int* __restrict__ counter, so_safe_record* __restrict__ out) {
    const SafePartial identity = {0, -INFINITY, -1, -INFINITY, -1};
    SafePartial acc = scan_safe(Q, q_stride, M, row0, S);
    if (grid_reduce(acc, part, counter, identity)) {
        out->n_safe = acc.n; out->max_l0 = acc.max_l; out->argmax_l0 = acc.arg_l; out->max_u0 = acc.max_u; out->argmax_u0 = acc.arg_u;
        out->reserved[0] = out->reserved[1] = out->reserved[2] = 0;
    }
}

__global__ void __launch_bounds__(kThreads) k_maximizers(const double* __restrict__ Q, int G, int64_t M, int64_t row0,
                                                        const uint8_t* __restrict__ S, double max_l0,
                                                        const so_safe_record* __restrict__ safe_recs, int n_recs, Scal scaling,
                                                        uint8_t* __restrict__ Mmask, MaxPartial* __restrict__ part,
                                                        unsigned int* __restrict__ counter, so_max_record* __restrict__ out) {
    if (safe_recs) {                            // chained: max over the ranks' records of max l0[S] (-inf where a rank has none)
        max_l0 = -INFINITY;
        for (int r = 0; r < n_recs; ++r) max_l0 = safe_recs[r].max_l0 > max_l0 ? safe_recs[r].max_l0 : max_l0;
    }
    const MaxPartial identity = {0, -INFINITY, -INFINITY, -1};
    MaxPartial acc = scan_maximizers(Q, G, M, row0, S, max_l0, scaling, Mmask);
    if (grid_reduce(acc, part, counter, identity)) {
        out->n_max = acc.n; out->max_width0 = acc.max_w0; out->best_value = acc.best; out->best_row = acc.best_row;
        out->reserved[0] = out->reserved[1] = out->reserved[2] = out->reserved[3] = 0;
    }
}

__global__ void __launch_bounds__(kThreads) k_candidates(const double* __restrict__ Q, int G, int64_t M, int64_t row0,
                                                        const uint8_t* __restrict__ S, const uint8_t* __restrict__ Mmask,
                                                        double max_var, const so_max_record* __restrict__ max_recs, int n_recs,
                                                        Scal scaling, Scal thr, uint8_t* __restrict__ cmask,
                                                        double* __restrict__ ckey, int64_t* __restrict__ crow, int64_t cap,
                                                        unsigned long long* __restrict__ n_cand) {
    if (max_recs) {                             // chained: gp_opt.py:513 from the ranks' maximiser records
        double w = -INFINITY;
        for (int r = 0; r < n_recs; ++r) w = max_recs[r].max_width0 > w ? max_recs[r].max_width0 : w;
        max_var = w / scaling.v[0];
    }
    scan_candidates(Q, G, M, row0, S, Mmask, max_var, scaling, thr, cmask, ckey, crow, cap, n_cand);
}

// ---------------------------------------------------------------- fused: the three passes and their cross-rank exchanges in ONE launch
// compute_sets needs max l0[S] before M and max width(M) before the candidates -- two true data dependencies, each a
// reduction over every rank's rows.  Chained as separate kernels with an NCCL all-gather of the 64-byte record after each
// (round 1) the chain cost 0.35 ms of a 2 ms step on eight GPUs.  Here one cooperative kernel does all of it: scan -> grid
// barrier -> block 0 combines the partials and publishes this rank's record into every rank's exchange buffer over NVLink
// (xchg.cuh) -> every block waits for the `world` stamps of the phase -> next scan.  When the kernel ends, the records of
// all ranks for all three phases sit in `result` (device) in the layout the host parses with one copy.
struct FusedParams {
    const double* Q;
    int G;
    int with_candidates;
    int64_t M, row0;
    const uint8_t* S;
    uint8_t* Mmask;
    Scal scaling, thr;
    double* ckey;
    int64_t* crow;
    int64_t cap;
    SafePartial* partA;
    MaxPartial* partB;
    unsigned int* bar;                   // {arrivals, generation}
    unsigned long long* ncand;
    unsigned long long* epoch;
    XchgView x;
    unsigned char* result;               // [world x safe record][world x max record][world x count][status][epoch]
    long long* dbg;                      // optional: SM-clock stamps of block 0 at the phase boundaries (so_debug_fused_times)
};

#define SO_FUSED_STAMP(i) do { if (p.dbg && threadIdx.x == 0) p.dbg[i] = (long long)globaltimer_dbg(); } while (0)

__device__ __forceinline__ unsigned long long globaltimer_dbg() {     // diagnostics only (SO_FUSED_DEBUG_TIMES): comparable across SMs
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

// All `world` stamps of a phase carry `epoch`?  (block-wide; returns the same answer in every thread)
__device__ __forceinline__ bool wait_phase(const unsigned long long* flags, int world, unsigned long long epoch) {
    int ok = 1;
    if ((int)threadIdx.x < world) ok = xchg_wait(flags, threadIdx.x, epoch, world > 1) ? 1 : 0;
    return __syncthreads_and(ok) != 0;
}

// Block 0, thread 0: stamp `epoch` into flag[rank] of every rank's buffer.  The record stores before it were issued by the same
// thread, so the release orders them (system scope across GPUs, device scope when this GPU is alone).
__device__ __forceinline__ void publish_flags(const XchgView& x, int par, int phase, unsigned long long epoch) {
    if (x.world == 1) { st_release_gpu(&x.local->sets[par].flag[phase][0], epoch); return; }
    // ONE system-scope fence orders the record stores to all peers before all stamps (a release store per peer would wait for
    // its own acknowledgement round trip each: ~3 us x world, 24 us per phase on eight GPUs); the stamps are then posted writes
    __threadfence_system();
    for (int r = 0; r < x.world; ++r) st_relaxed_sys(&x.peer[r]->sets[par].flag[phase][x.rank], epoch);
}

// SINGLE: one CTA of 1024 threads does everything (row blocks up to kSingleMaxRows): the grid barriers become __syncthreads and
// no partial leaves the SM -- on small grids the multi-block kernel is nothing but ~40 serialised L2 round trips (31 us at 40 k
// rows, measured with SO_FUSED_DEBUG_TIMES), the single-block one a handful.
constexpr int kSingleThreads = 1024;
constexpr int64_t kSingleMaxRows = 400000;

// Publishes this block's partial and takes a ticket; returns true (in every thread) in the LAST block to arrive, which then
// combines the partials and publishes the record -- the other blocks go straight to waiting for the stamp, so a phase costs
// one arrival + one stamp instead of a grid barrier followed by a stamp.  The ticket counter is reset by the last block before
// it publishes (nobody can reach the next phase's ticket before seeing that stamp).
template <typename P>
__device__ __forceinline__ bool arrive_last(const P& acc, P* __restrict__ part, unsigned int* __restrict__ counter) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = acc;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
        if (s_last) { __threadfence(); *counter = 0; }
    }
    __syncthreads();
    return s_last != 0;
}

template <bool SINGLE>
__global__ void __launch_bounds__(SINGLE ? kSingleThreads : kThreads) k_sets_fused(const __grid_constant__ FusedParams p) {
    __shared__ SafePartial smA[32];
    __shared__ MaxPartial smB[32];
    const unsigned long long epoch = *p.epoch + 1;          // bumped by the last block at the very end
    const int par = (int)(epoch & 1), world = p.x.world, rank = p.x.rank;
    XchgSets* mine = &p.x.local->sets[par];
    int status = SO_OK;

    SO_FUSED_STAMP(0);
    // ---- phase A: safe-set record
    {
        SafePartial acc = block_reduce(scan_safe(p.Q, 2 * p.G, p.M, p.row0, p.S), smA);
        SO_FUSED_STAMP(1);
        if (arrive_last(acc, p.partA, p.bar)) {
            SafePartial t = {0, -INFINITY, -1, -INFINITY, -1};
            for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) merge(t, p.partA[b]);
            t = block_reduce(t, smA);
            if (threadIdx.x == 0) {
                *p.ncand = 0;                                // first touched in phase C, two stamps from here
                so_safe_record rec;
                rec.n_safe = t.n; rec.max_l0 = t.max_l; rec.argmax_l0 = t.arg_l; rec.max_u0 = t.max_u; rec.argmax_u0 = t.arg_u;
                rec.reserved[0] = rec.reserved[1] = rec.reserved[2] = 0;
                for (int r = 0; r < world; ++r) p.x.peer[r]->sets[par].safe[rank] = rec;
                publish_flags(p.x, par, 0, epoch);
            }
        }
    }
    SO_FUSED_STAMP(3);
    if (!wait_phase(mine->flag[0], world, epoch)) status = SO_ERR_TIMEOUT;
    SO_FUSED_STAMP(4);
    double max_l0 = -INFINITY;
    for (int r = 0; r < world; ++r) {
        const double v = __ldcg(&mine->safe[r].max_l0);
        max_l0 = v > max_l0 ? v : max_l0;
    }

    // ---- phase B: maximisers
    {
        MaxPartial acc = block_reduce(scan_maximizers(p.Q, p.G, p.M, p.row0, p.S, max_l0, p.scaling, p.Mmask), smB);
        SO_FUSED_STAMP(5);
        if (arrive_last(acc, p.partB, p.bar)) {
            MaxPartial t = {0, -INFINITY, -INFINITY, -1};
            for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) merge(t, p.partB[b]);
            t = block_reduce(t, smB);
            if (threadIdx.x == 0) {
                so_max_record rec;
                rec.n_max = t.n; rec.max_width0 = t.max_w0; rec.best_value = t.best; rec.best_row = t.best_row;
                rec.reserved[0] = rec.reserved[1] = rec.reserved[2] = rec.reserved[3] = 0;
                for (int r = 0; r < world; ++r) p.x.peer[r]->sets[par].max[rank] = rec;
                publish_flags(p.x, par, 1, epoch);
            }
        }
    }
    SO_FUSED_STAMP(7);
    if (!wait_phase(mine->flag[1], world, epoch)) status = SO_ERR_TIMEOUT;
    SO_FUSED_STAMP(8);

    // ---- phase C: expander candidates (skipped for full_sets: every safe row is a candidate there)
    if (p.with_candidates) {
        double w = -INFINITY;
        for (int r = 0; r < world; ++r) {
            const double v = __ldcg(&mine->max[r].max_width0);
            w = v > w ? v : w;
        }
        scan_candidates(p.Q, p.G, p.M, p.row0, p.S, p.Mmask, w / p.scaling.v[0], p.scaling, p.thr, nullptr, p.ckey, p.crow, p.cap,
                        p.ncand);
    }
    SO_FUSED_STAMP(9);
    {
        SafePartial none = {0, -INFINITY, -1, -INFINITY, -1};
        if (!arrive_last(none, p.partA, p.bar)) return;     // everybody but the last block is done
    }
    SO_FUSED_STAMP(10);
    if (threadIdx.x == 0) {
        const long long n = p.with_candidates ? (long long)*reinterpret_cast<volatile unsigned long long*>(p.ncand) : 0;
        for (int r = 0; r < world; ++r) p.x.peer[r]->sets[par].ncand[rank] = n;
        publish_flags(p.x, par, 2, epoch);
    }
    // the kernel may only end once every rank's phase-C record has landed here: the host reads `result` next
    if (!wait_phase(mine->flag[2], world, epoch)) status = SO_ERR_TIMEOUT;
    SO_FUSED_STAMP(11);
    unsigned long long* res = reinterpret_cast<unsigned long long*>(p.result);
    const unsigned long long* src_safe = reinterpret_cast<const unsigned long long*>(mine->safe);
    const unsigned long long* src_max = reinterpret_cast<const unsigned long long*>(mine->max);
    for (int i = threadIdx.x; i < world * 8; i += blockDim.x) {
        res[i] = __ldcg(src_safe + i);
        res[world * 8 + i] = __ldcg(src_max + i);
    }
    for (int r = threadIdx.x; r < world; r += blockDim.x)
        res[world * 16 + r] = (unsigned long long)__ldcg(reinterpret_cast<const unsigned long long*>(mine->ncand) + r);
    if (threadIdx.x == 0) res[world * 17] = (unsigned long long)(long long)status;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned long long*>(res + world * 17 + 1) = epoch;     // the host polls this word (mapped pinned memory)
        *p.epoch = epoch;
    }
    SO_FUSED_STAMP(12);
}

int grid_for(const so_handle* h, int64_t M) {
    int64_t blocks = ((M + kRows - 1) / kRows + kThreads - 1) / kThreads;
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

static int run_maximizers(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d, double max_l0,
                          const so_safe_record* safe_recs_d, int n_recs, const double* scaling_h, uint8_t* Mmask_d,
                          so_max_record* rec_d, void* stream) {
    if (!h || !Q_d || !S_d || !scaling_h || !Mmask_d || !rec_d || n_gps < 1 || n_gps > 64 || M < 0) return SO_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    Scal sc;
    for (int i = 0; i < 64; ++i) sc.v[i] = i < n_gps ? scaling_h[i] : 1.0;
    k_maximizers<<<grid_for(h, M), kThreads, 0, (cudaStream_t)stream>>>(Q_d, n_gps, M, row0, S_d, max_l0, safe_recs_d, n_recs, sc,
                                                                        Mmask_d, (MaxPartial*)h->ws_partials, h->ws_counter, rec_d);
    SO_CHECK_LAUNCH(h, "k_maximizers");
    return SO_OK;
}

extern "C" int so_sets_maximizers(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                                  double max_l0, const double* scaling_h, uint8_t* Mmask_d, so_max_record* rec_d, void* stream) {
    return run_maximizers(h, Q_d, n_gps, M, row0, S_d, max_l0, nullptr, 0, scaling_h, Mmask_d, rec_d, stream);
}

extern "C" int so_sets_maximizers_chain(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                                        const so_safe_record* safe_recs_d, int n_recs, const double* scaling_h,
                                        uint8_t* Mmask_d, so_max_record* rec_d, void* stream) {
    if (!safe_recs_d || n_recs < 1) return SO_ERR_BAD_ARG;
    return run_maximizers(h, Q_d, n_gps, M, row0, S_d, 0.0, safe_recs_d, n_recs, scaling_h, Mmask_d, rec_d, stream);
}

static int run_candidates(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                          const uint8_t* Mmask_d, double max_var, const so_max_record* max_recs_d, int n_recs,
                          const double* scaling_h, const double* thr_h, uint8_t* cand_mask_d, double* cand_key_d,
                          int64_t* cand_row_d, int64_t cap, int64_t* n_cand_d, void* stream) {
    if (!h || !Q_d || !S_d || !Mmask_d || !scaling_h || !thr_h || !n_cand_d || n_gps < 1 || n_gps > 64 || M < 0 || cap < 0)
        return SO_ERR_BAD_ARG;
    if (cap > 0 && (!cand_key_d || !cand_row_d)) return SO_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    Scal sc, th;
    for (int i = 0; i < 64; ++i) { sc.v[i] = i < n_gps ? scaling_h[i] : 1.0; th.v[i] = i < n_gps ? thr_h[i] : 0.0; }
    SO_CUDA(h, cudaMemsetAsync(n_cand_d, 0, sizeof(int64_t), (cudaStream_t)stream));
    k_candidates<<<grid_for(h, M), kThreads, 0, (cudaStream_t)stream>>>(Q_d, n_gps, M, row0, S_d, Mmask_d, max_var, max_recs_d, n_recs,
                                                                        sc, th, cand_mask_d, cand_key_d, cand_row_d, cap,
                                                                        reinterpret_cast<unsigned long long*>(n_cand_d));
    SO_CHECK_LAUNCH(h, "k_candidates");
    return SO_OK;
}

extern "C" int so_sets_candidates(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                                  const uint8_t* Mmask_d, double max_var, const double* scaling_h, const double* thr_h,
                                  uint8_t* cand_mask_d, double* cand_key_d, int64_t* cand_row_d, int64_t cap,
                                  int64_t* n_cand_d, void* stream) {
    return run_candidates(h, Q_d, n_gps, M, row0, S_d, Mmask_d, max_var, nullptr, 0, scaling_h, thr_h, cand_mask_d, cand_key_d,
                          cand_row_d, cap, n_cand_d, stream);
}

extern "C" int so_sets_candidates_chain(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                                        const uint8_t* Mmask_d, const so_max_record* max_recs_d, int n_recs,
                                        const double* scaling_h, const double* thr_h, uint8_t* cand_mask_d, double* cand_key_d,
                                        int64_t* cand_row_d, int64_t cap, int64_t* n_cand_d, void* stream) {
    if (!max_recs_d || n_recs < 1) return SO_ERR_BAD_ARG;
    return run_candidates(h, Q_d, n_gps, M, row0, S_d, Mmask_d, 0.0, max_recs_d, n_recs, scaling_h, thr_h, cand_mask_d, cand_key_d,
                          cand_row_d, cap, n_cand_d, stream);
}

// ---------------------------------------------------------------- fused entry points
static int fused_setup(so_handle* h) {
    if (h->fused_part) return SO_OK;
    int rc = xchg_ensure_local(h);
    if (rc) return rc;
    SO_CUDA(h, cudaMalloc(&h->fused_part, (size_t)2 * SO_WS_MAX_BLOCKS * 64));
    SO_CUDA(h, cudaMalloc(&h->fused_bar, 2 * sizeof(unsigned int)));
    SO_CUDA(h, cudaMemset(h->fused_bar, 0, 2 * sizeof(unsigned int)));
    SO_CUDA(h, cudaMalloc(&h->fused_ncand, sizeof(unsigned long long)));
    // the combined records land in MAPPED pinned host memory: the kernel's last store is the epoch stamp, which the host polls
    // (no copy, no stream synchronisation on the step's critical path)
    SO_CUDA(h, cudaHostAlloc(&h->fused_result_h, SO_SETS_RESULT_BYTES(kXchgMaxWorld), cudaHostAllocMapped));
    memset(h->fused_result_h, 0, SO_SETS_RESULT_BYTES(kXchgMaxWorld));
    SO_CUDA(h, cudaHostGetDevicePointer(&h->fused_result_d, h->fused_result_h, 0));
    if (std::getenv("SO_FUSED_DEBUG_TIMES")) {
        SO_CUDA(h, cudaMalloc(&h->fused_dbg, 16 * sizeof(long long)));
        SO_CUDA(h, cudaMemset(h->fused_dbg, 0, 16 * sizeof(long long)));
    }
    int per_sm = 0;
    SO_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sets_fused<false>, kThreads, 0));
    if (per_sm < 1) return so_fail(h, SO_ERR_CUDA, "so_sets_fused: the fused kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;          // a few resident CTAs per SM saturate HBM; more only lengthens the grid barrier
    h->fused_grid = per_sm * h->num_sms;
    if (h->fused_grid > SO_WS_MAX_BLOCKS) h->fused_grid = SO_WS_MAX_BLOCKS;
    return SO_OK;
}

static int fused_collect(so_handle* h, void* result_h) {
    const size_t bytes = SO_SETS_RESULT_BYTES(h->xchg_world);
    memcpy(result_h, h->fused_result_h, bytes);
    const long long status = reinterpret_cast<const long long*>(h->fused_result_h)[17 * h->xchg_world];
    if (status == SO_ERR_TIMEOUT)
        return so_fail(h, SO_ERR_TIMEOUT, "so_sets_fused: a rank never published its record (peer exchange timed out)");
    return SO_OK;
}

extern "C" int so_sets_fused_result(so_handle* h, void* result_h, void* stream_) {
    if (!h || !result_h) return SO_ERR_BAD_ARG;
    if (!h->fused_part) return so_fail(h, SO_ERR_BAD_ARG, "so_sets_fused_result: so_sets_fused has not run");
    DeviceGuard guard(h->device);
    SO_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream_));
    h->fused_inflight = 0;
    return fused_collect(h, result_h);
}

extern "C" int so_sets_fused(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                             const double* scaling_h, const double* thr_h, int with_candidates, uint8_t* Mmask_d,
                             double* cand_key_d, int64_t* cand_row_d, int64_t cap, void* result_h, void* stream_) {
    // a rank may hold no rows at all (more ranks than row blocks): it still takes part in the exchange, with empty records
    if (!h || !scaling_h || !thr_h || n_gps < 1 || n_gps > 64 || M < 0 || cap < 0) return SO_ERR_BAD_ARG;
    if (M > 0 && (!Q_d || !S_d || !Mmask_d)) return SO_ERR_BAD_ARG;
    if (with_candidates && M > 0 && cap > 0 && (!cand_key_d || !cand_row_d)) return SO_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    int rc = fused_setup(h);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    FusedParams p;
    p.Q = Q_d; p.G = n_gps; p.with_candidates = with_candidates ? 1 : 0; p.M = M; p.row0 = row0; p.S = S_d; p.Mmask = Mmask_d;
    for (int i = 0; i < 64; ++i) { p.scaling.v[i] = i < n_gps ? scaling_h[i] : 1.0; p.thr.v[i] = i < n_gps ? thr_h[i] : 0.0; }
    p.ckey = cand_key_d; p.crow = cand_row_d; p.cap = cap;
    p.partA = static_cast<SafePartial*>(h->fused_part);
    p.partB = reinterpret_cast<MaxPartial*>(static_cast<unsigned char*>(h->fused_part) + (size_t)SO_WS_MAX_BLOCKS * 64);
    p.bar = h->fused_bar; p.ncand = h->fused_ncand; p.epoch = h->xchg_epochs;
    p.x = xchg_view(h);
    p.result = static_cast<unsigned char*>(h->fused_result_d);
    p.dbg = h->fused_dbg;
    volatile unsigned long long* stamp = reinterpret_cast<volatile unsigned long long*>(h->fused_result_h) + 17 * h->xchg_world + 1;
    const unsigned long long expect = *stamp + 1;      // meaningful only when no earlier launch is still in flight (checked below)
    const char* force = std::getenv("SO_SETS_SINGLE");  // "0": never the single-block kernel, "1": always (A/B measurements)
    // measured (tools/time_sets.py): one CTA is latency-bound on its dependent mask -> Q loads (53 us at 40 k rows, 208 us at 250 k,
    // against 31 us for the multi-block kernel), so it is an A/B switch only
    const bool single = force && force[0] == '1' && M <= kSingleMaxRows;
    if (single) {
        k_sets_fused<true><<<1, kSingleThreads, 0, stream>>>(p);
        SO_CHECK_LAUNCH(h, "k_sets_fused<single>");
    } else {
        int grid = grid_for(h, M);
        if (grid > h->fused_grid) grid = h->fused_grid;
        void* args[] = {&p};
        SO_CUDA(h, cudaLaunchCooperativeKernel((const void*)k_sets_fused<false>, dim3(grid), dim3(kThreads), args, 0, stream));
    }
    if (!result_h) { h->fused_inflight += 1; return SO_OK; }
    if (h->fused_inflight > 0) return so_sets_fused_result(h, result_h, stream_);
    // poll the epoch stamp the kernel writes last into the mapped result; look at the stream now and then so that a failed
    // launch or a faulting kernel surfaces as an error instead of a hang
    for (unsigned long long spins = 1;; ++spins) {
        if (*stamp == expect) break;
        if ((spins & 0x3fffu) == 0) {
            const cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaSuccess) { if (*stamp == expect) break; return so_sets_fused_result(h, result_h, stream_); }
            if (q != cudaErrorNotReady) return so_fail(h, SO_ERR_CUDA, std::string("so_sets_fused: ") + cudaGetErrorString(q));
        }
    }
    return fused_collect(h, result_h);
}

// Diagnostic: SM-clock stamps of block 0 at the phase boundaries of the last so_sets_fused (needs SO_FUSED_DEBUG_TIMES=1 in the
// environment when the handle first runs the fused kernel).  out_h: 16 int64.
extern "C" int so_debug_fused_times(so_handle* h, int64_t* out_h) {
    if (!h || !out_h) return SO_ERR_BAD_ARG;
    if (!h->fused_dbg) return so_fail(h, SO_ERR_BAD_ARG, "so_debug_fused_times: start the process with SO_FUSED_DEBUG_TIMES=1");
    DeviceGuard guard(h->device);
    SO_CUDA(h, cudaDeviceSynchronize());
    SO_CUDA(h, cudaMemcpy(out_h, h->fused_dbg, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return SO_OK;
}
