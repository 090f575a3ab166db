// K5/K6 -- SafeOptSwarm on the device.
//   so_swarm_fitness     : the epilogue of SafeOptSwarm._compute_particle_fitness
//                          (safeopt/gp_opt.py:926-1013) on per-GP posterior planes produced by
//                          so_posterior_rows; includes _compute_penalty (:874-899), scipy's expit
//                          (:960) and norm.pdf(., scale=0.2) (:1000).
//   so_swarm_step        : velocity / position update with clipping (safeopt/swarm.py:98-130).
//   so_swarm_update_best : personal / global best (safeopt/swarm.py:132-146).
// All three are elementwise, HBM-streaming kernels (tens of bytes per particle); operations are
// written with explicit roundings where the reference's NumPy expression would not contract.
#include "xchg.cuh"

int xchg_ensure_local(so_handle* h);
XchgView xchg_view(const so_handle* h);

namespace {

constexpr int kThreads = 256;
struct Vec64 { double v[64]; };
struct Vec16 { double v[SO_MAX_DIM]; };

__device__ __forceinline__ double penalty_of(double slack) {
    // safeopt/gp_opt.py:891-899
    double p = slack < 0.0 ? slack : 0.0;
    if (slack < 0.0 && slack > -0.001) p *= 2.0;
    if (slack <= -0.001 && slack > -0.1) p *= 5.0;
    if (slack <= -0.1 && slack > -1.0) p *= 10.0;
    if (slack < -1.0) p = -300.0 * (p * p);
    return p;
}

__device__ __forceinline__ double expit_f64(double x) {
    if (x >= 0.0) return 1.0 / (1.0 + exp(-x));
    const double e = exp(x);
    return e / (1.0 + e);
}

__global__ void __launch_bounds__(kThreads) k_swarm_fitness(int kind, int G, int64_t P, const double* __restrict__ mean,
                                                           const double* __restrict__ var, double beta, Vec64 fmin,
                                                           Vec64 scaling, double best_lower_bound,
                                                           double* __restrict__ values, uint8_t* __restrict__ safe_out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    double sd = sqrt(var[i]);
    double lower = __dsub_rn(mean[i], __dmul_rn(beta, sd));
    const double upper = __dadd_rn(mean[i], __dmul_rn(beta, sd));
    if (kind == SO_SWARM_GREEDY) {
        values[i] = lower;
        safe_out[i] = 1;
        return;
    }
    double val = sd / scaling.v[0];
    double interest = 1.0;
    if (kind == SO_SWARM_EXPANDERS) interest = (double)G;
    else if (kind == SO_SWARM_MAXIMIZERS) interest = expit_f64(10.0 * (upper - best_lower_bound) / scaling.v[0]);
    bool safe = true;
    double total_pen = 0.0;
    for (int g = 0; g < G; ++g) {
        if (g > 0) {
            sd = sqrt(var[(size_t)g * P + i]);
            lower = __dsub_rn(mean[(size_t)g * P + i], __dmul_rn(beta, sd));
            const double vs = sd / scaling.v[g];
            val = vs > val ? vs : val;
        }
        if (fmin.v[g] == -INFINITY) continue;
        double slack = lower - fmin.v[g];
        safe = safe && (slack >= 0.0);
        if (kind == SO_SWARM_SAFE_SET) continue;
        slack = slack / scaling.v[g];
        total_pen += penalty_of(slack);
        if (kind == SO_SWARM_EXPANDERS) {
            const double y = slack / 0.2;
            interest *= exp(-(y * y) / 2.0) / 2.5066282746310002 / 0.2;   // scipy norm.pdf(slack, scale=0.2)
        }
    }
    if (kind == SO_SWARM_SAFE_SET) {
        values[i] = lower;
    } else {
        values[i] = __dmul_rn(__dadd_rn(val, total_pen), interest);
    }
    safe_out[i] = safe ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads) k_swarm_step(int64_t P, int d, double* __restrict__ pos, double* __restrict__ vel,
                                                        const double* __restrict__ best_pos, const double* __restrict__ gbest,
                                                        const double* __restrict__ r, double inertia, Vec16 vscale, Vec16 lo,
                                                        Vec16 hi, int has_bounds) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= P * d) return;
    const int j = (int)(e % d);
    const double x = pos[e];
    const double d_self = __dsub_rn(best_pos[e], x);
    const double d_glob = __dsub_rn(gbest[j], x);
    const double r1 = r[e], r2 = r[(size_t)P * d + e];
    double v = __dmul_rn(vel[e], inertia);
    const double pull = __ddiv_rn(__dadd_rn(__dmul_rn(r1, d_self), __dmul_rn(r2, d_glob)), vscale.v[j]);
    v = __dadd_rn(v, pull);
    const double vmax = 10.0 * vscale.v[j];
    v = v < -vmax ? -vmax : (v > vmax ? vmax : v);
    double xn = __dadd_rn(x, v);
    if (has_bounds) xn = xn < lo.v[j] ? lo.v[j] : (xn > hi.v[j] ? hi.v[j] : xn);
    vel[e] = v;
    pos[e] = xn;
}

// ---- counter-based uniform randoms (Philox4x32-10): element (global particle, dimension) of iteration `it` gets the same two
// doubles whatever the sharding, so a swarm split over R ranks follows the single-GPU trajectory bit for bit.
__device__ __forceinline__ void philox4x32_10(unsigned (&c)[4], unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double u53(unsigned hi, unsigned lo) {          // uniform in [0, 1), 53 random bits
    const unsigned long long u = (((unsigned long long)hi << 32) | lo) >> 11;
    return (double)u * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ void swarm_rand2(unsigned long long seed, unsigned long long it, unsigned long long elem, double& r1, double& r2) {
    unsigned c[4] = {(unsigned)elem, (unsigned)(elem >> 32), (unsigned)it, (unsigned)(it >> 32)};
    philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
    r1 = u53(c[0], c[1]);
    r2 = u53(c[2], c[3]);
}

__global__ void __launch_bounds__(kThreads) k_swarm_rand(int64_t P, int d, int64_t p0, unsigned long long seed, unsigned long long it,
                                                        double* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= P * d) return;
    double r1, r2;
    swarm_rand2(seed, it, (unsigned long long)(p0 * d + e), r1, r2);
    out[e] = r1;
}

// k_swarm_step with the randoms drawn in the kernel and the inertia / iteration number read from device memory
// (state[0] = inertia, state[1] = inertia increment per iteration, state[2] = iteration number; advanced by
// k_swarm_update_best at the end of the iteration) -- nothing in the launch changes from one iteration to the next, so the
// whole PSO iteration can be replayed from a CUDA graph.
__global__ void __launch_bounds__(kThreads) k_swarm_step_dev(int64_t P, int d, int64_t p0, double* __restrict__ pos,
                                                            double* __restrict__ vel, const double* __restrict__ best_pos,
                                                            const double* __restrict__ gbest, const double* __restrict__ state,
                                                            unsigned long long seed, Vec16 vscale, Vec16 lo, Vec16 hi, int has_bounds) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= P * d) return;
    const int j = (int)(e % d);
    const double inertia = state[0];
    const unsigned long long it = (unsigned long long)state[2];
    double r1, r2;
    swarm_rand2(seed, it, (unsigned long long)(p0 * d + e), r1, r2);
    const double x = pos[e];
    const double d_self = __dsub_rn(best_pos[e], x);
    const double d_glob = __dsub_rn(gbest[j], x);
    double v = __dmul_rn(vel[e], inertia);
    const double pull = __ddiv_rn(__dadd_rn(__dmul_rn(r1, d_self), __dmul_rn(r2, d_glob)), vscale.v[j]);
    v = __dadd_rn(v, pull);
    const double vmax = 10.0 * vscale.v[j];
    v = v < -vmax ? -vmax : (v > vmax ? vmax : v);
    double xn = __dadd_rn(x, v);
    if (has_bounds) xn = xn < lo.v[j] ? lo.v[j] : (xn > hi.v[j] ? hi.v[j] : xn);
    vel[e] = v;
    pos[e] = xn;
}

struct BestPartial { double v; long long idx; };

__global__ void __launch_bounds__(kThreads) k_swarm_update_best(int64_t P, int d, const double* __restrict__ pos,
                                                               const double* __restrict__ values, const uint8_t* __restrict__ safe,
                                                               double* __restrict__ best_pos, double* __restrict__ best_values,
                                                               BestPartial* __restrict__ part, unsigned int* __restrict__ counter,
                                                               int64_t* __restrict__ best_idx, int64_t p0, double* __restrict__ rec,
                                                               XchgView x, unsigned long long* __restrict__ epoch_p,
                                                               double* __restrict__ gbest, double* __restrict__ grec,
                                                               double* __restrict__ state) {
    BestPartial acc = {-INFINITY, -1};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < P; i += (int64_t)gridDim.x * kThreads) {
        double bv = best_values[i];
        const double v = values[i];
        if (v > bv && safe[i]) {
            bv = v;
            best_values[i] = v;
            for (int j = 0; j < d; ++j) best_pos[(size_t)i * d + j] = pos[(size_t)i * d + j];
        }
        // np.argmax: first index of the maximum (NaN handling not replicated)
        if (acc.idx < 0 || bv > acc.v) { acc.v = bv; acc.idx = i; }
    }
    __shared__ BestPartial sm[kThreads / 32];
    __shared__ bool last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, acc.v, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, acc.idx, o);
        if (oi >= 0 && (acc.idx < 0 || ov > acc.v || (ov == acc.v && oi < acc.idx))) { acc.v = ov; acc.idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        BestPartial t = sm[0];
        for (int w = 1; w < kThreads / 32; ++w) {
            const BestPartial o = sm[w];
            if (o.idx >= 0 && (t.idx < 0 || o.v > t.v || (o.v == t.v && o.idx < t.idx))) t = o;
        }
        part[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        BestPartial t = {-INFINITY, -1};
        for (unsigned b = 0; b < gridDim.x; ++b) {
            const BestPartial o = part[b];
            if (o.idx >= 0 && (t.idx < 0 || o.v > t.v || (o.v == t.v && o.idx < t.idx))) t = o;
        }
        *best_idx = t.idx;
        *counter = 0;
        if (rec) {
            // every personal best of this launch was written before the owning block took its ticket
            rec[0] = t.v;
            rec[1] = (double)(p0 + t.idx);
            for (int j = 0; j < d; ++j) rec[2 + j] = best_pos[(size_t)t.idx * d + j];
        }
        if (state) { state[0] += state[1]; state[2] += 1.0; }      // `inertia += step` of swarm.py:117, next iteration's randoms
        if (gbest) {
            // exchange in the kernel: this rank's record goes into every rank's buffer (NVLink stores + a release stamp), then
            // the `world` stamps of this epoch are awaited and the global best is combined -- largest value, ties to the
            // lowest global particle index (np.argmax over the unsharded swarm, swarm.py:146)
            const unsigned long long epoch = *epoch_p + 1;
            const int par = (int)(epoch & 1);
            for (int r = 0; r < x.world; ++r) {
                double* dst = x.peer[r]->swarm[par].rec[x.rank];
                dst[0] = t.v;
                dst[1] = (double)(p0 + t.idx);
                for (int j = 0; j < d; ++j) dst[2 + j] = best_pos[(size_t)t.idx * d + j];
            }
            const XchgSwarm* mine = &x.local->swarm[par];
            bool ok = true;
            if (x.world > 1) {          // alone, the record just written is the whole exchange
                __threadfence_system();          // one fence, then posted stamp stores (not one acknowledged release per peer)
                for (int r = 0; r < x.world; ++r) st_relaxed_sys(&x.peer[r]->swarm[par].flag[x.rank], epoch);
                for (int r = 0; r < x.world; ++r) ok = xchg_wait(mine->flag, r, epoch, true) && ok;
            }
            int best = -1;
            double bv = 0.0, bi = 0.0;
            for (int r = 0; r < x.world; ++r) {
                const double v = x.world > 1 ? __ldcg(&mine->rec[r][0]) : mine->rec[r][0], i = x.world > 1 ? __ldcg(&mine->rec[r][1]) : mine->rec[r][1];
                if (i < 0.0) continue;
                if (best < 0 || v > bv || (v == bv && i < bi)) { best = r; bv = v; bi = i; }
            }
            if (best >= 0) {
                for (int j = 0; j < d; ++j) gbest[j] = x.world > 1 ? __ldcg(&mine->rec[best][2 + j]) : mine->rec[best][2 + j];
                if (grec) { grec[0] = bv; grec[1] = bi; grec[2] = ok ? 0.0 : (double)SO_ERR_TIMEOUT; }
            }
            *epoch_p = epoch;
        }
    }
}

// Global best over the R per-rank records {value, global index, position[d]} (stride SO_SWARM_REC_DOUBLES):
// largest value, ties to the lowest global particle index (np.argmax over the unsharded swarm, swarm.py:146).
__global__ void k_swarm_combine_best(const double* __restrict__ recs, int R, int d, double* __restrict__ gbest,
                                     double* __restrict__ grec) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int best = -1;
    for (int r = 0; r < R; ++r) {
        const double v = recs[(size_t)r * SO_SWARM_REC_DOUBLES], i = recs[(size_t)r * SO_SWARM_REC_DOUBLES + 1];
        if (i < 0.0) continue;
        if (best < 0 || v > recs[(size_t)best * SO_SWARM_REC_DOUBLES] ||
            (v == recs[(size_t)best * SO_SWARM_REC_DOUBLES] && i < recs[(size_t)best * SO_SWARM_REC_DOUBLES + 1])) best = r;
    }
    if (best < 0) return;
    const double* b = recs + (size_t)best * SO_SWARM_REC_DOUBLES;
    for (int j = 0; j < d; ++j) gbest[j] = b[2 + j];
    if (grec) { grec[0] = b[0]; grec[1] = b[1]; }
}

}  // namespace

extern "C" int so_swarm_fitness(so_handle* h, int kind, int n_gps, int64_t P, const double* mean_d, const double* var_d,
                                double beta, const double* fmin_h, const double* scaling_h, double best_lower_bound,
                                double* values_d, uint8_t* safe_d, void* stream) {
    if (!h || !mean_d || !var_d || !fmin_h || !scaling_h || !values_d || !safe_d || P < 0) return SO_ERR_BAD_ARG;
    if (n_gps < 1 || n_gps > 64) return so_fail(h, SO_ERR_BAD_ARG, "swarm_fitness: 1 <= n_gps <= 64");
    if (kind < SO_SWARM_GREEDY || kind > SO_SWARM_SAFE_SET) return so_fail(h, SO_ERR_BAD_ARG, "swarm_fitness: bad kind");
    if (P == 0) return SO_OK;
    DeviceGuard guard(h->device);
    Vec64 fm, sc;
    for (int i = 0; i < 64; ++i) { fm.v[i] = i < n_gps ? fmin_h[i] : 0.0; sc.v[i] = i < n_gps ? scaling_h[i] : 1.0; }
    k_swarm_fitness<<<(unsigned)((P + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        kind, n_gps, P, mean_d, var_d, beta, fm, sc, best_lower_bound, values_d, safe_d);
    SO_CHECK_LAUNCH(h, "k_swarm_fitness");
    return SO_OK;
}

extern "C" int so_swarm_step(so_handle* h, int64_t P, int d, double* pos_d, double* vel_d, const double* best_pos_d,
                             const double* global_best_d, const double* r_d, double inertia, const double* velocity_scale_h,
                             const double* bounds_h, void* stream) {
    if (!h || !pos_d || !vel_d || !best_pos_d || !global_best_d || !r_d || !velocity_scale_h || P < 0) return SO_ERR_BAD_ARG;
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "swarm_step: 1 <= d <= 16");
    if (P == 0) return SO_OK;
    DeviceGuard guard(h->device);
    Vec16 vs, lo, hi;
    for (int j = 0; j < SO_MAX_DIM; ++j) {
        vs.v[j] = j < d ? velocity_scale_h[j] : 1.0;
        lo.v[j] = (bounds_h && j < d) ? bounds_h[2 * j] : 0.0;
        hi.v[j] = (bounds_h && j < d) ? bounds_h[2 * j + 1] : 0.0;
    }
    k_swarm_step<<<(unsigned)((P * d + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        P, d, pos_d, vel_d, best_pos_d, global_best_d, r_d, inertia, vs, lo, hi, bounds_h ? 1 : 0);
    SO_CHECK_LAUNCH(h, "k_swarm_step");
    return SO_OK;
}

extern "C" int so_swarm_update_best(so_handle* h, int64_t P, int d, const double* pos_d, const double* values_d,
                                    const uint8_t* safe_d, double* best_pos_d, double* best_values_d, int64_t* best_idx_d,
                                    int64_t p0, double* rec_d, void* stream) {
    if (!h || !pos_d || !values_d || !safe_d || !best_pos_d || !best_values_d || !best_idx_d || P < 1) return SO_ERR_BAD_ARG;
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "swarm_update_best: 1 <= d <= 16");
    DeviceGuard guard(h->device);
    int64_t blocks = (P + kThreads - 1) / kThreads;
    if (blocks > SO_WS_MAX_BLOCKS) blocks = SO_WS_MAX_BLOCKS;
    XchgView none;
    none.local = nullptr; none.world = 0; none.rank = 0;
    for (int r = 0; r < kXchgMaxWorld; ++r) none.peer[r] = nullptr;
    k_swarm_update_best<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        P, d, pos_d, values_d, safe_d, best_pos_d, best_values_d, (BestPartial*)h->ws_partials, h->ws_counter, best_idx_d, p0, rec_d,
        none, nullptr, nullptr, nullptr, nullptr);
    SO_CHECK_LAUNCH(h, "k_swarm_update_best");
    return SO_OK;
}

extern "C" int so_swarm_update_best_x(so_handle* h, int64_t P, int d, const double* pos_d, const double* values_d,
                                      const uint8_t* safe_d, double* best_pos_d, double* best_values_d, int64_t* best_idx_d,
                                      int64_t p0, double* global_best_d, double* global_rec_d, double* state_d, void* stream) {
    if (!h || !pos_d || !values_d || !safe_d || !best_pos_d || !best_values_d || !best_idx_d || !global_best_d || P < 1)
        return SO_ERR_BAD_ARG;
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "swarm_update_best_x: 1 <= d <= 16");
    DeviceGuard guard(h->device);
    int rc = xchg_ensure_local(h);
    if (rc) return rc;
    int64_t blocks = (P + kThreads - 1) / kThreads;
    if (blocks > SO_WS_MAX_BLOCKS) blocks = SO_WS_MAX_BLOCKS;
    k_swarm_update_best<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        P, d, pos_d, values_d, safe_d, best_pos_d, best_values_d, (BestPartial*)h->ws_partials, h->ws_counter, best_idx_d, p0, nullptr,
        xchg_view(h), h->xchg_epochs + 1, global_best_d, global_rec_d, state_d);
    SO_CHECK_LAUNCH(h, "k_swarm_update_best_x");
    return SO_OK;
}

extern "C" int so_swarm_rand(so_handle* h, int64_t P, int d, int64_t p0, uint64_t seed, uint64_t counter, double* out_d, void* stream) {
    if (!h || !out_d || P < 0 || p0 < 0) return SO_ERR_BAD_ARG;
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "swarm_rand: 1 <= d <= 16");
    if (P == 0) return SO_OK;
    DeviceGuard guard(h->device);
    k_swarm_rand<<<(unsigned)((P * d + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, d, p0, seed, counter, out_d);
    SO_CHECK_LAUNCH(h, "k_swarm_rand");
    return SO_OK;
}

extern "C" int so_swarm_step_dev(so_handle* h, int64_t P, int d, int64_t p0, double* pos_d, double* vel_d, const double* best_pos_d,
                                 const double* global_best_d, const double* state_d, uint64_t seed, const double* velocity_scale_h,
                                 const double* bounds_h, void* stream) {
    if (!h || !pos_d || !vel_d || !best_pos_d || !global_best_d || !state_d || !velocity_scale_h || P < 0 || p0 < 0) return SO_ERR_BAD_ARG;
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "swarm_step_dev: 1 <= d <= 16");
    if (P == 0) return SO_OK;
    DeviceGuard guard(h->device);
    Vec16 vs, lo, hi;
    for (int j = 0; j < SO_MAX_DIM; ++j) {
        vs.v[j] = j < d ? velocity_scale_h[j] : 1.0;
        lo.v[j] = (bounds_h && j < d) ? bounds_h[2 * j] : 0.0;
        hi.v[j] = (bounds_h && j < d) ? bounds_h[2 * j + 1] : 0.0;
    }
    k_swarm_step_dev<<<(unsigned)((P * d + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        P, d, p0, pos_d, vel_d, best_pos_d, global_best_d, state_d, seed, vs, lo, hi, bounds_h ? 1 : 0);
    SO_CHECK_LAUNCH(h, "k_swarm_step_dev");
    return SO_OK;
}

extern "C" int so_swarm_combine_best(so_handle* h, const double* recs_d, int n_ranks, int d, double* global_best_d,
                                     double* global_rec_d, void* stream) {
    if (!h || !recs_d || !global_best_d || n_ranks < 1) return SO_ERR_BAD_ARG;
    if (d < 1 || d > SO_MAX_DIM) return so_fail(h, SO_ERR_UNSUPPORTED, "swarm_combine_best: 1 <= d <= 16");
    DeviceGuard guard(h->device);
    k_swarm_combine_best<<<1, 32, 0, (cudaStream_t)stream>>>(recs_d, n_ranks, d, global_best_d, global_rec_d);
    SO_CHECK_LAUNCH(h, "k_swarm_combine_best");
    return SO_OK;
}
