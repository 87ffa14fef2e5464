// sem_pf.cu -- bootstrap particle filter on the device (replaces pmcmc.py:123-233).
//
// pf_persistent runs the whole filter in ONE cooperative launch (grid.sync() is the resampling barrier); pf_init +
// pf_step run the same phases with one kernel launch per observation step p (the launch boundary is the barrier):
//   pf_step(p):  [resample]   ancestor a_j = first i with cdf_{p-1}(i) > u_j * total      (pmcmc.py:187-193)
//                [gather]     x = X[p-1][:, a_j]   (fused into the load, no separate gather pass, :195-199)
//                [propagate]  exact Gillespie SSA over one observation interval            (gillespie_algo.py)
//                [store]      X[p][:, j] = x, coalesced SoA int32                          (pmcmc.py:222-231)
//                [weigh]      logw_j = min_c log pmf/pdf(Y[p][c] | x_c)                    (pmcmc.py:178-181, used at p+1)
//                [scan]       CTA-local max m_b and inclusive scan L of exp(logw - m_b) (warp shuffles + smem)
//                [finalize]   the last CTA to finish combines the per-CTA (m_b, s_b) partials: global max M,
//                             CTA prefixes, total, and log_zetas[p+1] = log_zetas[p] + M + log(total) - log(N)
//                             (pmcmc.py:183 in the log domain); flags collapse (pmcmc.py:191-192).
//   cdf_p(i) for i in CTA b  =  pfx[b] + scale[b] * L[i],  scale[b] = exp(m_b - M): no second pass over the weights.
// pf_step(0) initialises X_0 (given, or I_0 ~ Poisson(mu) per pmcmc.py:156-169) and weighs it against Y[0].
// Timing convention of the reference is kept: step p weighs X[p-1] against Y[p-1]; the last state is
// never weighed.
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "sem_common.cuh"
#include "sem_host.h"

namespace sem {

struct PfDev {
    int N, T, Cobs, obs_kind, resampler, nb, ppb, hist_rows, model, n_filters, ntheta, init_poisson, pfx_in_smem;
    double probs, dt;
    PhiloxKey key;
    uint32_t filter_id0;
    double mu[SEM_MAX_GROUPS], npop[SEM_MAX_GROUPS];
    const double *Y, *theta;
    const int32_t *X0;
    const double *res_u, *ssa_u;
    const long long *ssa_off;
    int32_t *X_hist, *ancestry, *status;
    double *log_zetas;
    unsigned long long *n_events;
    // workspace (double-buffered by step parity)
    double *L[2];        // [F][N]   CTA-local inclusive scan of exp(logw - m_b)
    double *pfx[2];      // [F][nb]  exclusive prefix of scale_b * s_b
    double *scale[2];    // [F][nb]
    double *total[2];    // [F]
    double2 *part;       // [2][F][nb]  (m_b, s_b), by step parity
    unsigned int *counter;  // [F]
    // particle-sharded filter (one shard of a larger filter, see sem_shard_*): global index of particle 0, the
    // pre-gathered children records [N][C+1] (state, global ancestor) and the (M, total) summary of the local weights
    int j0, sharded;
    int path_exact;
    double *wtab;        // [T-1][Cobs][wt_n + 1] log-weight of a compartment count, or null (see weight_table_fill)
    int wt_n;
    double *iter_out;    // [F][SEM_ITER_HEADER + T*C] packed result of one PMCMC iteration, or null
    int split_main;      // pf_persistent: > 0 = particles [split_main, ppb) of a CTA are shared by two warps each (see there)
    const int32_t *X_in;
    double *summary;
};

constexpr int kMaxThreads = 768;
constexpr int kMaxThreadsUnif = 352;      // the uniformized step keeps more live state: two 352-thread CTAs per SM, <= 93 registers

// CTA-wide max / inclusive scan: warp shuffles, one shared-memory slot per warp, and a second shuffle pass over the
// (at most 32) warp results done redundantly by every warp -- no serial loop over the warps.
__device__ __forceinline__ double block_max(double v, double *sm, int tid, int nwarps) {
    v = warp_max_d(v);
    __syncthreads();
    if ((tid & 31) == 0) sm[tid >> 5] = v;
    __syncthreads();
    return warp_max_d((tid & 31) < nwarps ? sm[tid & 31] : -CUDART_INF);
}

// inclusive scan over the CTA; returns this thread's inclusive value, *total = CTA sum
__device__ __forceinline__ double block_incl_scan(double v, double *sm, int tid, int nwarps, double *total) {
    const int lane = tid & 31, w = tid >> 5;
    v = warp_incl_scan_d(v, lane);
    __syncthreads();
    if (lane == 31) sm[w] = v;
    __syncthreads();
    const double ws = warp_incl_scan_d(lane < nwarps ? sm[lane] : 0.0, lane);   // inclusive scan of the warp totals
    *total = __shfl_sync(0xffffffffu, ws, nwarps - 1);
    const double off = __shfl_sync(0xffffffffu, ws, w > 0 ? w - 1 : 0);
    return w > 0 ? v + off : v;
}

// log-weight of one observed column given the compartment count (pmcmc.py:179,181)
__device__ __forceinline__ double column_logw(const PfDev &P, double y, double xc, const double2 *tab) {
    return (P.obs_kind == SEM_OBS_BINOMIAL) ? binom_logpmf_obs(binom_obs(y, tab), xc, P.probs, tab) : norm_logpdf(y, xc, P.probs, tab);
}

// The weight of a column depends on the particle only through an integer count in [0, total population], and Y is
// known up front: all (T-1) x Cobs x (pop+1) values are tabulated once per launch (weight_table_fill; 3*10^6 evaluations
// for the headline instead of 3*10^7 per pass) and the per-particle weight becomes Cobs L2-resident loads.  Same
// function, same values: results are bit-identical to the direct evaluation.
__device__ __forceinline__ void weight_table_fill(const PfDev &P, size_t first, size_t stride, const double2 *tab) {
    const size_t per_col = (size_t)P.wt_n + 1, total = (size_t)(P.T - 1) * P.Cobs * per_col;
    for (size_t i = first; i < total; i += stride) {
        const size_t pc = i / per_col;
        const double y = P.Y[pc];                            // Y[p][c], p = pc / Cobs
        P.wtab[i] = (y != y) ? 0.0 : column_logw(P, y, (double)(i - pc * per_col), tab);
    }
}

__global__ void __launch_bounds__(256) weight_table_kernel(const __grid_constant__ PfDev P) {
    __shared__ double2 s_tab[kLogTabSize];
    load_logtab(s_tab);
    __syncthreads();
    weight_table_fill(P, blockIdx.x * (size_t)blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x, s_tab);
}

template <class Model>
__device__ __forceinline__ double particle_logw(const PfDev &P, const double *x, const double *Yrow, const double *wrow, const double2 *tab) {
    double lw = CUDART_INF;
#pragma unroll
    for (int c = 0; c < Model::C; c++) {
        if (c < P.Cobs) {
            double xc = x[c];
            if (P.model == SEM_MODEL_SIR_SUBGROUPS2) {      // observes the group sum of each compartment (pmcmc.py:172-173)
                xc = 0.0;
#pragma unroll
                for (int g = 0; g < Model::G; g++) xc += x[3 * g + (c % 3)];
            }
            const double y = Yrow[c];
            if (y != y) continue;                            // extension: a NaN entry of Y marks an unobserved column
            double l;
            if (wrow && xc >= 0.0 && xc <= (double)P.wt_n) l = wrow[(size_t)c * (P.wt_n + 1) + (int)xc];
            else l = column_logw(P, y, xc, tab);
            lw = (l < lw || l != l) ? l : lw;                // min over columns; NaN sticks
        }
    }
    return lw == CUDART_INF ? 0.0 : lw;                      // nothing observed at this time: weight 1
}

#ifdef SEM_PHASES
__device__ unsigned long long g_phase[16 * 256];
#define PHASE(k) do { if (tid == 0 && b == 0 && p < 256) { g_phase[p * 16 + (k)] = (unsigned long long)clock64(); } } while (0)
__device__ unsigned long long g_warp_end[256 * 32];          // CTA 0: when each warp left the SSA loop, and its work
__device__ unsigned int g_warp_work[256 * 32];
#define WARP_END(work) do { const unsigned int wk_ = __reduce_max_sync(0xffffffffu, (unsigned int)(work)); \
    if ((tid & 31) == 0 && b == 0 && p < 256) { g_warp_end[p * 32 + (tid >> 5)] = (unsigned long long)clock64(); g_warp_work[p * 32 + (tid >> 5)] = wk_; } } while (0)
#else
#define PHASE(k)
#define WARP_END(work)
#endif

// Weigh the CTA's particles against Y[p] and CTA-local scan: writes L[par] and the CTA partial (m_b, s_b).
template <class Model>
__device__ __forceinline__ void weigh_local(const PfDev &P, const int p, const int f, const int b, const int tid,
                                            const bool active, const int j, const double *x, double *sm,
                                            const double2 *tab) {
    const int N = P.N, par = p & 1;
    double lw = -CUDART_INF;
    if (active) {
        lw = particle_logw<Model>(P, x, P.Y + (size_t)p * P.Cobs, P.wtab ? P.wtab + (size_t)p * P.Cobs * (P.wt_n + 1) : nullptr, tab);
        if (lw != lw) lw = CUDART_INF;                       // NaN -> +inf marker => collapse in the combine
    }
    PHASE(8);
    const int nwarps = (blockDim.x + 31) >> 5;
    const double mb = block_max(lw, sm, tid, nwarps);
    PHASE(9);
    const double e = (active && mb > -CUDART_INF && mb < CUDART_INF) ? exp(lw - mb) : 0.0;
    double sb;
    const double incl = block_incl_scan(e, sm, tid, nwarps, &sb);
    PHASE(10);
    if (active) P.L[par][(size_t)f * N + j] = incl;
    if (tid == 0) P.part[((size_t)par * P.n_filters + f) * P.nb + b] = make_double2(mb, sb);
}

// Combine the CTA partials of parity `par` into (M, total) and per-CTA (prefix, scale) written to pfx_out/scale_out
// (shared or global memory).  Called by every thread of a CTA.
__device__ __forceinline__ void combine_partials(const PfDev &P, const int f, const int par, const int tid, double *sm,
                                                 double *pfx_out, double *scale_out, double &M_out, double &total_out) {
    const int nwarps = (blockDim.x + 31) >> 5;
    const double2 *part = P.part + ((size_t)par * P.n_filters + f) * P.nb;
    if (P.nb <= (int)blockDim.x) {                           // one partial per thread: a single L2 round trip
        double2 ps = make_double2(-CUDART_INF, 0.0);
        if (tid < P.nb) ps = __ldcg(&part[tid]);
        const double M = block_max(ps.x, sm, tid, nwarps);
        const bool finiteM = (M > -CUDART_INF && M < CUDART_INF);
        const double sc = (tid < P.nb && finiteM && ps.x > -CUDART_INF) ? exp(ps.x - M) : 0.0, val = sc * ps.y;
        double tot;
        const double incl = block_incl_scan(val, sm, tid, nwarps, &tot);
        if (tid < P.nb) { pfx_out[tid] = 0.0 + (incl - val); scale_out[tid] = sc; }
        M_out = M; total_out = 0.0 + tot;
        return;
    }
    double M = -CUDART_INF;
    for (int i = tid; i < P.nb; i += blockDim.x) M = fmax(M, __ldcg(&part[i].x));
    M = block_max(M, sm, tid, nwarps);
    double carry = 0.0;
    const bool finiteM = (M > -CUDART_INF && M < CUDART_INF);
    for (int i0 = 0; i0 < P.nb; i0 += blockDim.x) {
        const int i = i0 + tid;
        double sc = 0.0, val = 0.0;
        if (i < P.nb && finiteM) {
            const double mi = __ldcg(&part[i].x), si = __ldcg(&part[i].y);
            sc = (mi > -CUDART_INF) ? exp(mi - M) : 0.0;
            val = sc * si;
        }
        double chunk;
        const double incl2 = block_incl_scan(val, sm, tid, nwarps, &chunk);
        if (i < P.nb) { pfx_out[i] = carry + (incl2 - val); scale_out[i] = sc; }
        carry += chunk;
    }
    M_out = M; total_out = carry;
}

// Weigh the CTA's particles against Y[p], CTA-local scan, and (last CTA to arrive) the step's global combine.
template <class Model>
__device__ __forceinline__ void weigh_scan_finalize(const PfDev &P, const int p, const int f, const int b, const int tid,
                                                    const bool active, const int j, const double *x, double *sm,
                                                    const double2 *tab, bool *is_last) {
    const int N = P.N, par = p & 1;
    weigh_local<Model>(P, p, f, b, tid, active, j, x, sm, tab);          // weigh against Y[p] (pmcmc.py:178-181) + CTA scan

    // ------------------------------------------------------------------------ last CTA finalizes the step
    __threadfence();
    __syncthreads();
    if (tid == 0) *is_last = (atomicAdd(&P.counter[f], 1u) == (unsigned)(P.nb - 1));
    __syncthreads();
    if (!*is_last) return;
    __threadfence();
    double M, carry;
    combine_partials(P, f, par, tid, sm, P.pfx[par] + (size_t)f * P.nb, P.scale[par] + (size_t)f * P.nb, M, carry);
    const bool finiteM = (M > -CUDART_INF && M < CUDART_INF);
    if (tid == 0) {
        P.total[par][f] = carry;
        if (P.sharded) {                                     // the host combines the shards' (M, total) summaries
            P.summary[0] = M; P.summary[1] = carry;
            P.counter[f] = 0;
            return;
        }
        double *lz = P.log_zetas + (size_t)f * P.T;
        if (!finiteM || !(carry > 0.0)) {
            P.status[f] = p + 1;                              // np.random.choice raises at step p+1 (pmcmc.py:191-192)
            for (int q = p + 1; q < P.T; q++) lz[q] = -CUDART_INF;
        } else {
            lz[p + 1] = lz[p] + M + log(carry) - log((double)N);    // zetas[p+1] = zetas[p] * mean(w), pmcmc.py:183
        }
        P.counter[f] = 0;
    }
}

// Step 0: X_0 (pmcmc.py:156-170), given or I_0 ~ Poisson(mu), then weigh against Y[0].
template <class Model>
__global__ void __launch_bounds__(kMaxThreads) pf_init(const __grid_constant__ PfDev P) {
    __shared__ double sm[32];
    __shared__ double2 s_tab[kLogTabSize];
    __shared__ bool is_last;
    const int f = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
    const int N = P.N, j = b * P.ppb + tid;
    const bool active = tid < P.ppb && j < N;
    const uint32_t fid = P.filter_id0 + f;
    load_logtab(s_tab);                                      // the weights' logarithms (made visible by the barriers of the scan)
    double x[Model::C];
    if (active) {
        if (!P.init_poisson) {
#pragma unroll
            for (int c = 0; c < Model::C; c++) x[c] = (double)P.X0[(size_t)c * N + j];
        } else {
#pragma unroll
            for (int c = 0; c < Model::C; c++) x[c] = 0.0;
#pragma unroll
            for (int g = 0; g < Model::G; g++) {
                PairSource<false> src; src.init(P.key, (uint32_t)(P.j0 + j), (uint32_t)g, stream_word(DOM_INIT, fid));
                const double i0 = poisson_draw(src, P.mu[g]);
                constexpr bool seir = (Model::C == 4);
                x[seir ? 2 : 3 * g + 1] = i0;
                x[seir ? 0 : 3 * g] = P.npop[g] - i0;
            }
        }
        P.ancestry[(size_t)f * P.hist_rows * N + j] = 0;
        int32_t *Xr = P.X_hist + (size_t)f * P.hist_rows * Model::C * N;
#pragma unroll
        for (int c = 0; c < Model::C; c++) Xr[(size_t)c * N + j] = (int32_t)x[c];
    }
    __syncthreads();                                         // s_tab complete
    if (P.T > 1) weigh_scan_finalize<Model>(P, 0, f, b, tid, active, j, x, sm, s_tab, &is_last);
}

// The step's single systematic-resampling uniform (one Philox call, the same for every thread)
__device__ __forceinline__ double systematic_u0(const PfDev &P, const int p, const uint32_t fid) {
    const uint4 w = philox4x32_10(0u, 0u, (uint32_t)p, stream_word(DOM_RESAMPLE, fid), P.key);
    return bits_to_d12(w.x, w.y) - 1.0;
}

// Ancestor of slot j at step p (pmcmc.py:187-193): first particle whose cdf exceeds u_j * total, by a two-level
// search: CTA prefixes (shared or global memory), then the CTA's local scan L of the previous step.
template <bool REPLAY>
__device__ __forceinline__ int select_ancestor(const PfDev &P, const int p, const int f, const int j, const uint32_t fid,
                                               const double *pfx, const double *scale, const double total) {
    const int N = P.N, par = p & 1;
    double u;
    if (REPLAY) u = P.res_u[(size_t)(p - 1) * N + j];
    else if (P.resampler == SEM_RESAMPLE_SYSTEMATIC) {
        u = __ddiv_rn(__dadd_rn((double)j, systematic_u0(P, p, fid)), (double)N);
    } else {
        const uint4 w = philox4x32_10(0u, (uint32_t)j, (uint32_t)p, stream_word(DOM_RESAMPLE, fid), P.key);
        u = bits_to_d12(w.x, w.y) - 1.0;
    }
    const double v = __dmul_rn(u, total);
    int lo = 0, hi = P.nb;                                  // last CTA index with pfx[b] <= v  (pfx[0] = 0)
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pfx[mid] <= v) lo = mid; else hi = mid; }
    const int base = lo * P.ppb, len = min(P.ppb, N - base);
    const double sc = scale[lo], pf = pfx[lo];
    const double *L = P.L[par ^ 1] + (size_t)f * N + base;
    int a = 0, e = len;                                      // first i with pf + sc*L[i] > v
    while (a < e) { const int mid = (a + e) >> 1; if (__fma_rn(sc, __ldcg(&L[mid]), pf) <= v) a = mid + 1; else e = mid; }
    return base + min(a, len - 1);
}

// End of one PMCMC iteration (pmcmc.py:371 particle_path_sampler + the three small results the MH loop reads), by one
// thread per filter after everything else of the filter is globally visible.  Same draw and indexing as
// path_sample_kernel.
template <int C>
__device__ void iteration_epilogue(const PfDev &P, const int f) {
    const int T = P.T, N = P.N;
    double *out = P.iter_out + (size_t)f * (SEM_ITER_HEADER + (size_t)T * C);
    const int status = *(volatile int32_t *)&P.status[f];
    out[0] = __ldcg(&P.log_zetas[(size_t)f * T + T - 1]);
    out[1] = (double)status;
    out[2] = P.n_events ? (double)__ldcg(&P.n_events[f]) : 0.0;
    if (status != 0) { out[3] = -1.0; return; }
    const int32_t *X = P.X_hist + (size_t)f * P.hist_rows * C * N, *A = P.ancestry + (size_t)f * P.hist_rows * N;
    const uint4 w = philox4x32_10(0u, 0u, 0u, stream_word(DOM_PATH, P.filter_id0 + f), P.key);
    int chosen = min((int)((bits_to_d12(w.x, w.y) - 1.0) * (double)N), N - 1);          // np.random.randint(0, N) (pmcmc.py:241)
    out[3] = (double)chosen;
    double *traj = out + SEM_ITER_HEADER;
#pragma unroll
    for (int c = 0; c < C; c++) traj[(size_t)(T - 1) * C + c] = (double)__ldcg(&X[((size_t)(T - 1) * C + c) * N + chosen]);
    for (int p = T - 2; p >= 0; p--) {
        chosen = __ldcg(&A[(size_t)(P.path_exact ? p + 1 : p) * N + chosen]);           // reference indexes row p
#pragma unroll
        for (int c = 0; c < C; c++) traj[(size_t)p * C + c] = (double)__ldcg(&X[((size_t)p * C + c) * N + chosen]);
    }
}

template <int C>
__global__ void iteration_epilogue_kernel(const __grid_constant__ PfDev P) {
    if (threadIdx.x == 0) iteration_epilogue<C>(P, blockIdx.x);
}

// Whole filter in ONE cooperative launch (one CTA per SM, all co-resident): the resampling barrier of every step is
// a grid.sync(); after it every CTA combines the nb CTA partials itself (nb <= 1024 values, redundantly) instead of
// waiting for a "last CTA" and a new launch.  Same arithmetic as pf_init + pf_step, bit-identical results.
template <class Model, int ARITH>
__global__ void __launch_bounds__(kMaxThreads) pf_persistent(const __grid_constant__ PfDev P) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double s_dyn[];          // pfx[nb], scale[nb] of the previous step (+ the sorted layout's exchange area)
    double *s_pfx = s_dyn, *s_scale = s_dyn + P.nb;
    __shared__ double sm[32];
    __shared__ double2 s_tab[kLogTabSize];
    __shared__ unsigned long long s_pairs;
    const int f = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
    // Thread -> particle.  Plain: thread t owns particle t of the CTA.  Balanced (P.split_main = 128 W > 0): the CTA holds
    // 128 W + e particles, e <= 64, i.e. W full warps per scheduler plus up to two more warps' worth -- which would make
    // two of the four schedulers run W + 1 full rounds while the others idle.  Instead the extra particles are split in
    // TIME between two helper warps on different schedulers: warp 4W + g runs group g (32 particles) until t >= dt / 2,
    // hands the continuation (state, time, stream counter) over through shared memory, and warp 4W + 2 + g finishes the
    // interval and owns the particle in the weights / scan.  Every scheduler then carries W + 1/2 rounds.  The legs
    // reproduce the single run bit for bit (ssa_run_spec_leg).
    constexpr bool kLegs = LegLoop<Model, ARITH>::available;
    constexpr bool kUnif = ARITH == SEM_ARITH_UNIFORMIZED32;  // its own two-leg loop (ssa_unif32_leg)
    const int N = P.N, warp = tid >> 5, lane = tid & 31;
    const int main_n = (kLegs && P.split_main > 0) ? P.split_main : (int)blockDim.x;
    const int helper = tid < main_n ? -1 : warp - (main_n >> 5);          // -1 main; 0,1 first leg of group 0,1; 2,3 second leg
    const int pidx = helper < 0 ? tid : main_n + 32 * (helper & 1) + lane;
    const int j = b * P.ppb + pidx;
    const bool has = pidx < P.ppb && j < N;
    const bool starts = has && helper < 2;                   // resamples, gathers and starts the interval
    const bool active = has && (helper < 0 || helper >= 2);  // owns the particle at the observation time (store, weigh, scan)
    __shared__ double s_cx[(kLegs || kUnif) ? 2 : 1][32][Model::C], s_ct[(kLegs || kUnif) ? 2 : 1][32];
    __shared__ uint32_t s_ck[(kLegs || kUnif) ? 2 : 1][32];
    __shared__ int s_cfin[(kLegs || kUnif) ? 2 : 1][32];
    __shared__ double s_cB[kUnif ? 2 : 1][32], s_ch[kUnif ? 2 : 1][32];   // rest of the uniformized loop's continuation
    __shared__ uint32_t s_cu[kUnif ? 2 : 1][32][3];
    const uint32_t fid = P.filter_id0 + f;
    int32_t *Xf = P.X_hist + (size_t)f * P.hist_rows * Model::C * N;
    int32_t *Af = P.ancestry + (size_t)f * P.hist_rows * N;
    load_logtab(s_tab);
    if (tid == 0) s_pairs = 0ull;
    if (b == 0 && tid == 0) {                                // the launch needs no memsets: the first grid.sync orders these
        P.status[f] = 0;                                     // before any other CTA's write
        P.log_zetas[(size_t)f * P.T] = 0.0;                  // zetas[0] = 1 (pmcmc.py:154)
        if (P.n_events) P.n_events[f] = 0ull;
    }
    __syncthreads();
    if (P.wtab) {                                            // tabulate the observation weights (all CTAs, once per launch)
        const size_t nthr = (size_t)gridDim.x * gridDim.y * blockDim.x;
        weight_table_fill(P, ((size_t)f * gridDim.x + b) * blockDim.x + tid, nthr, s_tab);
        grid.sync();
    }
    double x[Model::C];
    // ------------------------------------------------------------------------ step 0: X_0 (pmcmc.py:156-170)
    if (active) {
        if (!P.init_poisson) {
#pragma unroll
            for (int c = 0; c < Model::C; c++) x[c] = (double)P.X0[(size_t)c * N + j];
        } else {
#pragma unroll
            for (int c = 0; c < Model::C; c++) x[c] = 0.0;
#pragma unroll
            for (int g = 0; g < Model::G; g++) {
                PairSource<false> src; src.init(P.key, (uint32_t)(P.j0 + j), (uint32_t)g, stream_word(DOM_INIT, fid));
                const double i0 = poisson_draw(src, P.mu[g]);
                constexpr bool seir = (Model::C == 4);
                x[seir ? 2 : 3 * g + 1] = i0;
                x[seir ? 0 : 3 * g] = P.npop[g] - i0;
            }
        }
        Af[j] = 0;
#pragma unroll
        for (int c = 0; c < Model::C; c++) Xf[(size_t)c * N + j] = (int32_t)x[c];
    }
    if (P.T > 1) weigh_local<Model>(P, 0, f, b, tid, active, j, x, sm, s_tab);
    bool dead = false;
    double lz = 0.0;
    unsigned long long my_pairs = 0;
    for (int p = 1; p < P.T; p++) {
        PHASE(0);
        grid.sync();                                         // the resampling barrier (grid-wide fence + barrier)
        PHASE(1);
        if (dead) continue;
        const int par = p & 1;
        const int row = p % P.hist_rows, prow = (p + P.hist_rows - 1) % P.hist_rows;
        double M, total;
        combine_partials(P, f, par ^ 1, tid, sm, s_pfx, s_scale, M, total);
        const bool ok = (M > -CUDART_INF && M < CUDART_INF) && (total > 0.0);
        if (!ok) {
            if (b == 0 && tid == 0) {
                double *lzp = P.log_zetas + (size_t)f * P.T;
                P.status[f] = p;                             // np.random.choice raises at step p (pmcmc.py:191-192)
                for (int q = p; q < P.T; q++) lzp[q] = -CUDART_INF;
            }
            dead = true;
            continue;
        }
        if constexpr (kUnif) {                               // sorted layout: clear the bins and the range of K
            if (P.split_main < 0) {
                uint32_t *bins = (uint32_t *)(s_dyn + 2 * P.nb + 2 * blockDim.x) + (2 * Model::C + 3) * blockDim.x;
                for (int i = tid; i < 130; i += blockDim.x) bins[i] = i == 129 ? 0xffffffffu : 0u;
            }
        }
        __syncthreads();                                     // s_pfx / s_scale complete
        // zetas[p] = zetas[p-1] * mean(w) (pmcmc.py:183), off the CTA's critical path: by the last thread, whose warp is
        // a second-leg helper waiting for its hand-over in the balanced layout
        if (b == 0 && tid == (int)blockDim.x - 1) {
            lz = lz + M + log(total) - log((double)N);
            P.log_zetas[(size_t)f * P.T + p] = lz;
        }
        PHASE(2);
        long long pairs = 0;
        int32_t *Xr = Xf + (size_t)row * Model::C * N;
        Model m;
        PairSource<false> src;
        if (starts) {
            const int a = select_ancestor<false>(P, p, f, j, fid, s_pfx, s_scale, total);
            PHASE(6);
            Af[(size_t)row * N + j] = a;
            const int32_t *Xp = Xf + (size_t)prow * Model::C * N;
#pragma unroll
            for (int c = 0; c < Model::C; c++) x[c] = (double)__ldcg(&Xp[(size_t)c * N + a]);   // written by other CTAs: L2, not L1
            m.setup(P.theta + (size_t)f * P.ntheta, x);
            src.init(P.key, (uint32_t)(P.j0 + j), (uint32_t)p, stream_word(DOM_SSA, fid));
            PHASE(7);
        }
        if constexpr (kUnif) {
            // The uniformized loop knows its amount of work BEFORE it runs: the candidate count K of the (first) batch is
            // drawn in the setup.  Sorted layout (P.split_main < 0): every thread sets up its own particle, the CTA sorts
            // the particles by K (counting sort over 128 bins in shared memory) and thread t runs the particle at sorted
            // position slot(t), so the 32 lanes of a warp carry (nearly) equal work, and the sorted chunks of 32 are dealt
            // to the warps in snake order over the four schedulers (warp w issues on scheduler w & 3), which balances
            // the schedulers' sums.  The end state returns to the particle's home thread through shared memory.  Streams
            // are keyed by the particle, so WHO runs it changes nothing: results stay bit-identical to the plain layout.
            const bool sorted = P.split_main < 0;
            bool run = false;
            Unif32State ust;
            unif32_begin(ust, P.dt);
            PairSource<false> aux;
            int home = has ? pidx : -1, leg = 0, hg = 0;     // leg: 0 whole interval, 1 / 2 first / second leg of helper group hg
            if (starts) {
                aux.init(P.key, (uint32_t)(P.j0 + j), (uint32_t)p, stream_word(DOM_AUX, fid));
                double r0[Model::R], a00;
                run = unif32_batch_setup(m, x, ust, aux, r0, a00, s_tab);
                ust.aux_k = aux.k;
            }
            PHASE(12);
            if (sorted) {
                const int NT = blockDim.x;
                double *x_h = s_dyn + 2 * P.nb, *x_B = x_h + NT;
                int32_t *x_x = (int32_t *)(x_B + NT), *x_ret = x_x + Model::C * NT;
                uint32_t *x_K = (uint32_t *)(x_ret + Model::C * NT), *x_aux = x_K + NT;
                int32_t *x_home = (int32_t *)(x_aux + NT);
                uint32_t *x_hist = (uint32_t *)(x_home + NT);         // [128] bins + [2] range; zeroed before the barrier above
                const uint32_t K = run ? ust.last : 0u;
                const uint32_t wmax = __reduce_max_sync(0xffffffffu, K), wmin = __reduce_min_sync(0xffffffffu, K ? K : 0xffffffffu);
                if (lane == 0) { atomicMax(&x_hist[128], wmax); atomicMin(&x_hist[129], wmin); }
                __syncthreads();
                PHASE(13);
                const uint32_t kmax = x_hist[128], kmin = min(x_hist[129], kmax);
                const float inv = 126.0f / (float)(kmax - kmin + 1u);
                // descending in K; then the absorbed particles (nothing to run); threads without a particle come last
                const int bin = !has ? 127 : K ? min(125, (int)((float)(kmax - K) * inv)) : 126;
                const uint32_t rank = atomicAdd(&x_hist[bin], 1u);
                __syncthreads();
                PHASE(14);
                const uint4 hh = reinterpret_cast<const uint4 *>(x_hist)[lane];           // every warp scans the 128 bins itself
                const uint32_t s4 = hh.x + hh.y + hh.z + hh.w;
                uint32_t inc = s4;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
                const uint32_t e0 = inc - s4, e1 = e0 + hh.x, e2 = e1 + hh.y, e3 = e2 + hh.z;
                const int sl = bin >> 2, sk = bin & 3;
                const uint32_t t0 = __shfl_sync(0xffffffffu, e0, sl), t1 = __shfl_sync(0xffffffffu, e1, sl),
                               t2 = __shfl_sync(0xffffffffu, e2, sl), t3 = __shfl_sync(0xffffffffu, e3, sl);
                const int pos = (int)((sk == 0 ? t0 : sk == 1 ? t1 : sk == 2 ? t2 : t3) + rank);
                x_K[pos] = K; x_home[pos] = home < 0 ? -1 : (home | (run ? 0x40000000 : 0)); x_h[pos] = ust.h;   // (K = 0 may still have to run: a batch that covers only part of the interval)
                x_B[pos] = ust.B; x_aux[pos] = ust.aux_k;
                if (has) {
#pragma unroll
                    for (int c = 0; c < Model::C; c++) x_x[c * NT + pos] = (int32_t)x[c];
                }
                __syncthreads();
                PHASE(15);
                // Sorted chunk -> warp.  The main warps take the chunks in snake order over the schedulers.  When the
                // chunks are 4 W + 1 or 4 W + 2 (P.split_main == -2) the last one or two are shared in TIME by two helper
                // warps each, on different schedulers: warp 4W + g serves the first half of the batch's candidates
                // and hands the continuation over, warp 4W + 2 + g finishes the interval -- W + 1/2 rounds per scheduler
                // instead of W + 1 on two of them.
                const int nw = NT >> 5, main_w = P.split_main == -2 ? nw - 4 : nw;
                int chunk;
                if (warp < main_w) {
                    const int rnd = warp >> 2, r_last = (main_w - 1) >> 2;
                    chunk = 4 * rnd + ((((r_last - rnd) & 1) == 0) ? (warp & 3) : 3 - (warp & 3));
                } else {
                    hg = (warp - main_w) & 1;
                    leg = (warp - main_w) < 2 ? 1 : 2;
                    chunk = main_w + hg;
                }
                const int slot = 32 * chunk + lane;
                const int hv = x_home[slot];
                const bool run_rec = hv >= 0 && (hv & 0x40000000) != 0;
                home = hv < 0 ? -1 : (hv & 0x3fffffff);
                run = false;
                if (leg == 2) {                              // second leg: wait for the continuation
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + hg) : "memory");
                    if (home >= 0) {
#pragma unroll
                        for (int c = 0; c < Model::C; c++) x[c] = s_cx[hg][lane][c];
                        if (!s_cfin[hg][lane]) {
                            run = true;
                            ust.t_rem = s_ct[hg][lane]; ust.B = s_cB[hg][lane]; ust.h = s_ch[hg][lane];
                            ust.cand = s_ck[hg][lane]; ust.first = s_cu[hg][lane][0]; ust.last = s_cu[hg][lane][1]; ust.aux_k = s_cu[hg][lane][2];
                        }
                    }
                } else {
                    run = run_rec;
                    if (home >= 0) {
#pragma unroll
                        for (int c = 0; c < Model::C; c++) x[c] = (double)x_x[c * NT + slot];
                    }
                    if (run) { ust.h = x_h[slot]; ust.B = x_B[slot]; ust.last = x_K[slot]; ust.aux_k = x_aux[slot]; }
                }
            }
            if (sorted && run) m.setup(P.theta + (size_t)f * P.ntheta, x);
            PHASE(11);
            bool fin = true;
            if (run) {                                       // ONE call site of the loop
                const int jr = sorted ? b * P.ppb + home : j;
                long long fired = 0;
                ust.in_batch = 1;
                src.init(P.key, (uint32_t)(P.j0 + jr), (uint32_t)p, stream_word(DOM_SSA, fid));
                aux.init(P.key, (uint32_t)(P.j0 + jr), (uint32_t)p, stream_word(DOM_AUX, fid));
                fin = ssa_unif32_leg<Model, false>(m, x, ust, fired, leg == 1, src, aux, s_tab);
                pairs = fired;
            }
            if (leg == 1) {                                  // hand over
#pragma unroll
                for (int c = 0; c < Model::C; c++) s_cx[hg][lane][c] = x[c];
                s_ct[hg][lane] = ust.t_rem; s_cB[hg][lane] = ust.B; s_ch[hg][lane] = ust.h;
                s_ck[hg][lane] = ust.cand; s_cu[hg][lane][0] = ust.first; s_cu[hg][lane][1] = ust.last; s_cu[hg][lane][2] = ust.aux_k;
                s_cfin[hg][lane] = fin ? 1 : 0;
                __threadfence_block();
                asm volatile("bar.sync %0, 64;" ::"r"(1 + hg) : "memory");
            }
            WARP_END(run ? ust.last : 0u);
            if (sorted && home >= 0 && leg != 1) {           // back to the home thread (read after the barrier below)
                const int NT = blockDim.x;
                int32_t *x_ret = (int32_t *)(s_dyn + 2 * P.nb + 2 * NT) + Model::C * NT;
#pragma unroll
                for (int c = 0; c < Model::C; c++) x_ret[c * NT + home] = (int32_t)x[c];
            }
        } else if constexpr (kLegs) {
            // ONE call site of the event loop for every role: warps that ran different copies of the loop side by side
            // on a scheduler cost 14 % (instruction cache), measured
            const int g = helper & 1;
            bool run = starts, fin = true;
            double t = 0.0;
            const double handoff = (helper == 0 || helper == 1) ? 0.5 * P.dt : CUDART_INF;
            if (helper >= 2) {                               // second leg: wait for the continuation
                asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory");
                run = false;
                if (has) {
#pragma unroll
                    for (int c = 0; c < Model::C; c++) x[c] = s_cx[g][lane][c];
                    if (!s_cfin[g][lane]) {
                        run = true;
                        t = s_ct[g][lane];
                        m.setup(P.theta + (size_t)f * P.ntheta, x);
                        src.init(P.key, (uint32_t)(P.j0 + j), (uint32_t)p, stream_word(DOM_SSA, fid));
                        src.k = s_ck[g][lane];
                    }
                }
            }
            if (run) pairs = ssa_run_spec_leg<Model, LegLoop<Model, ARITH>::U, LegLoop<Model, ARITH>::bits32>(m, x, t, handoff, P.dt, src, s_tab, fin);
            if (helper == 0 || helper == 1) {                // first leg: hand over
#pragma unroll
                for (int c = 0; c < Model::C; c++) s_cx[g][lane][c] = x[c];
                s_ct[g][lane] = t; s_ck[g][lane] = src.k; s_cfin[g][lane] = fin ? 1 : 0;
                __threadfence_block();
                asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory");
            }
        } else {
            if (starts) pairs = ssa_run<Model, ARITH, false, false>(m, x, P.dt, src, s_tab, NoRec());
        }
        const bool via_smem = kUnif && P.split_main < 0;     // sorted layout: the state comes home after the barrier
        if (active && !via_smem) {
#pragma unroll
            for (int c = 0; c < Model::C; c++) Xr[(size_t)c * N + j] = (int32_t)x[c];
        }
        my_pairs += (unsigned long long)pairs;
        PHASE(3);
#ifndef SEM_NO_SSA_BARRIER
        __syncthreads();                                     // keep the CTA in the SSA loop until its last warp is done: letting early
#endif
        PHASE(4);                                            // warps run ahead into the weights code costs 27% (measured; profiles/)
        if constexpr (kUnif) {
            if (active && via_smem) {
                const int32_t *x_ret = (const int32_t *)(s_dyn + 2 * P.nb + 2 * blockDim.x) + Model::C * blockDim.x;
#pragma unroll
                for (int c = 0; c < Model::C; c++) {
                    const int32_t v = x_ret[c * blockDim.x + pidx];
                    x[c] = (double)v;
                    Xr[(size_t)c * N + j] = v;
                }
            }
        }
        if (p < P.T - 1) weigh_local<Model>(P, p, f, b, tid, active, j, x, sm, s_tab);
        PHASE(5);
    }
    if (P.n_events) {                                        // one global atomic per CTA for the whole filter
#pragma unroll
        for (int d = 16; d; d >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, d);
        if ((tid & 31) == 0 && my_pairs) atomicAdd(&s_pairs, my_pairs);
        __syncthreads();
        if (tid == 0 && s_pairs) atomicAdd(&P.n_events[f], s_pairs);
    }
    if (P.iter_out) {                                        // path sample + packed result of the MH iteration
        grid.sync();
        if (b == 0 && tid == 0) iteration_epilogue<Model::C>(P, f);
    }
}

// Step p >= 1: resample, gather, propagate, store, weigh (see the file header).
template <class Model, int ARITH, bool REPLAY>
__global__ void __launch_bounds__(ARITH == SEM_ARITH_UNIFORMIZED ? kMaxThreadsUnif : kMaxThreads)
__maxnreg__(ARITH == SEM_ARITH_UNIFORMIZED ? 88 : 80) pf_step(const __grid_constant__ PfDev P, const int p) {
    extern __shared__ double s_pfx[];                        // previous step's CTA prefixes (when they fit)
    __shared__ double sm[32];
    __shared__ double2 s_tab[kLogTabSize];
    __shared__ unsigned long long s_pairs;
    __shared__ bool is_last;
    const int f = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
    if (P.status[f] != 0) return;                            // collapsed (or replay exhausted) earlier
    const int N = P.N, j = b * P.ppb + tid;
    const bool active = tid < P.ppb && j < N;
    const int par = p & 1;
    const int row = p % P.hist_rows, prow = (p + P.hist_rows - 1) % P.hist_rows;
    const double *pfx_g = P.pfx[par ^ 1] + (size_t)f * P.nb;
    load_logtab(s_tab);
    if (P.pfx_in_smem && !P.sharded) for (int i = tid; i < P.nb; i += blockDim.x) s_pfx[i] = pfx_g[i];
    if (tid == 0) s_pairs = 0ull;
    __syncthreads();
    const double *pfx = P.pfx_in_smem ? s_pfx : pfx_g;
    int32_t *Xf = P.X_hist + (size_t)f * P.hist_rows * Model::C * N;
    int32_t *Af = P.ancestry + (size_t)f * P.hist_rows * N;
    const uint32_t fid = P.filter_id0 + f;
    double x[Model::C];
    long long pairs = 0;
    bool replay_dry = false;

    if (active && P.sharded) {
        // -------------------------------------------------------------------- children records delivered by the exchange
        const int32_t *rec = P.X_in + (size_t)j * (Model::C + 1);
#pragma unroll
        for (int c = 0; c < Model::C; c++) x[c] = (double)rec[c];
        Af[(size_t)row * N + j] = rec[Model::C];             // global ancestor index
    }
    if (active && !P.sharded) {
        // -------------------------------------------------------------------- resample (pmcmc.py:187-193)
        const int a = select_ancestor<REPLAY>(P, p, f, j, fid, pfx, P.scale[par ^ 1] + (size_t)f * P.nb, P.total[par ^ 1][f]);
        Af[(size_t)row * N + j] = a;
        // -------------------------------------------------------------------- gather parent (pmcmc.py:195-199)
        const int32_t *Xp = Xf + (size_t)prow * Model::C * N;
#pragma unroll
        for (int c = 0; c < Model::C; c++) x[c] = (double)Xp[(size_t)c * N + a];
    }
    if (active) {
        // -------------------------------------------------------------------- propagate (pmcmc.py:200-220)
        Model m;
        m.setup(P.theta + (size_t)f * P.ntheta, x);
        PairSource<REPLAY> src;
        if constexpr (REPLAY) {
            const size_t q = (size_t)(p - 1) * N + j;
            src.init(P.ssa_u, P.ssa_off[q], P.ssa_off[q + 1]);
        } else {
            src.init(P.key, (uint32_t)(P.j0 + j), (uint32_t)p, stream_word(DOM_SSA, fid));
        }
        pairs = ssa_run<Model, ARITH, REPLAY, false>(m, x, P.dt, src, s_tab, NoRec());
        if (pairs < 0) { replay_dry = true; pairs = 0; }
        // -------------------------------------------------------------------- store X[p] (SoA, coalesced)
        int32_t *Xr = Xf + (size_t)row * Model::C * N;
#pragma unroll
        for (int c = 0; c < Model::C; c++) Xr[(size_t)c * N + j] = (int32_t)x[c];
    }
    if (replay_dry) atomicExch(&P.status[f], SEM_ERR_REPLAY);
    if (P.n_events) {                                        // one global atomic per CTA, not per particle
        unsigned long long wp = (unsigned long long)pairs;
#pragma unroll
        for (int d = 16; d; d >>= 1) wp += __shfl_xor_sync(0xffffffffu, wp, d);
        if ((tid & 31) == 0 && wp) atomicAdd(&s_pairs, wp);
        __syncthreads();
        if (tid == 0 && s_pairs) atomicAdd(&P.n_events[f], s_pairs);
    }
    if (p >= P.T - 1) return;                                // the final state is never weighed
    weigh_scan_finalize<Model>(P, p, f, b, tid, active, j, x, sm, s_tab, &is_last);
}

// ---------------------------------------------------------------------------------------------- sharded filter
// Global systematic resampling across shards.  Slot j of the global next generation draws
// v_j = ((j + u0)/N) * Total and takes the first particle i whose global cdf exceeds v_j.  In offspring form:
// particle i owns the slots [J(lower_i), J(upper_i)) with J(c) = smallest j with v_j >= c, and every boundary is
// shared bit-for-bit by its two neighbours (CTA prefixes inside a shard, G_r / G_next between shards), so the
// slots are covered without gaps.  Each shard writes one record (state, global ancestor index) per child, ordered
// by slot; the host all-to-all-v's the records to the shards that own the slots.
struct SlotMap { double u0, Nd, total; long long N; };

__device__ __forceinline__ double slot_v(const SlotMap &sm, long long j) {
    return __dmul_rn(__ddiv_rn(__dadd_rn((double)j, sm.u0), sm.Nd), sm.total);
}
__device__ __forceinline__ long long first_slot_ge(const SlotMap &sm, double c) {
    if (c >= sm.total) return sm.N;                          // the global total closes the last particle's range
    const double g = ceil(__dsub_rn(__dmul_rn(__ddiv_rn(c, sm.total), sm.Nd), sm.u0));
    long long j = g < 0.0 ? 0 : (g > sm.Nd ? sm.N : (long long)g);
    while (j > 0 && slot_v(sm, j - 1) >= c) j--;
    while (j < sm.N && slot_v(sm, j) < c) j++;
    return j;
}

struct OffDev {
    int N, C, nb, ppb, j0;
    const int32_t *X;            // [C][N] states of the generation being resampled
    const double *L, *pfx, *scale;
    double total_local, G, G_next, s;   // this shard's cdf = G + s * local cdf; G_next = next shard's G (or Total)
    SlotMap sm;
    long long slot0;             // J(G): first slot owned by this shard's particles
    int32_t *send;               // [n_children][C+1]
};

__global__ void pf_offspring(const OffDev P) {
    const int b = blockIdx.x, t = threadIdx.x, i = b * P.ppb + t;
    if (t >= P.ppb || i >= P.N) return;
    const int len = min(P.ppb, P.N - b * P.ppb);
    const double pf = P.pfx[b], sc = P.scale[b];
    const double lower = (t == 0) ? pf : __fma_rn(sc, P.L[i - 1], pf);
    const bool last_in_cta = (t == len - 1), last_cta = (b == P.nb - 1);
    const double upper = last_in_cta ? (last_cta ? P.total_local : P.pfx[b + 1]) : __fma_rn(sc, P.L[i], pf);
    const double glo = (i == 0) ? P.G : __fma_rn(P.s, lower, P.G);
    const double ghi = (last_in_cta && last_cta) ? P.G_next : __fma_rn(P.s, upper, P.G);
    const long long c_lo = first_slot_ge(P.sm, glo), c_hi = first_slot_ge(P.sm, ghi);
    if (c_hi <= c_lo) return;
    int32_t st[SEM_MAX_GROUPS * 3];
    for (int c = 0; c < P.C; c++) st[c] = P.X[(size_t)c * P.N + i];
    for (long long ch = c_lo; ch < c_hi; ch++) {
        int32_t *rec = P.send + (size_t)(ch - P.slot0) * (P.C + 1);
        for (int c = 0; c < P.C; c++) rec[c] = st[c];
        rec[P.C] = P.j0 + i;
    }
}

// (T,C,N) int32 -> (T,N,C) float64, the layout pmcmc.py:151 returns
__global__ void hist_to_f64_kernel(const int32_t *X, int T, int N, int C, double *out) {
    const size_t n = (size_t)T * N * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = i / ((size_t)N * C), r = i - t * (size_t)N * C;
        const int jj = (int)(r / C), c = (int)(r - (size_t)jj * C);
        out[i] = (double)X[(t * C + c) * N + jj];
    }
}

__global__ void i32_to_f64_kernel(const int32_t *in, size_t n, double *out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

// particle_path_sampler (pmcmc.py:236-248): one thread chases the genealogy backwards.
__global__ void path_sample_kernel(const int32_t *X, const int32_t *A, int T, int N, int C, int chosen, int exact,
                                   const __grid_constant__ PhiloxKey key, uint32_t fid, int32_t *traj) {
    if (blockIdx.x || threadIdx.x) return;
    if (chosen < 0) {                                          // np.random.randint(0, N) (pmcmc.py:241)
        const uint4 w = philox4x32_10(0u, 0u, 0u, stream_word(DOM_PATH, fid), key);
        chosen = min((int)((bits_to_d12(w.x, w.y) - 1.0) * (double)N), N - 1);
    }
    for (int c = 0; c < C; c++) traj[(size_t)(T - 1) * C + c] = X[((size_t)(T - 1) * C + c) * N + chosen];
    for (int p = T - 2; p >= 0; p--) {
        chosen = A[(size_t)(exact ? p + 1 : p) * N + chosen];  // reference indexes row p
        for (int c = 0; c < C; c++) traj[(size_t)p * C + c] = X[((size_t)p * C + c) * N + chosen];
    }
}

// ---------------------------------------------------------------------------------------------- host side
// particles per CTA: one CTA per SM when the whole population is co-resident, else 256-wide CTAs
static int choose_ppb(const sem_pf_config *c) {
    if (c->block_particles > 0) {
        const int lim = c->arith == SEM_ARITH_UNIFORMIZED ? kMaxThreadsUnif : kMaxThreads;
        return c->block_particles > lim ? lim : c->block_particles;
    }
    const long long all = (long long)c->n_particles * c->n_filters;
    const long long per_sm = (all + sm_count() - 1) / sm_count();
    // co-resident population: ceil(N*F/SMs) particles per SM split over as few CTAs as the thread cap allows (threads =
    // particles rounded up to a warp), so the grid is a whole number of CTAs per SM; larger populations use 256-wide
    // CTAs scheduled in waves
    const int cap = c->arith == SEM_ARITH_UNIFORMIZED ? kMaxThreadsUnif : kMaxThreads;
    const long long per_cta = (per_sm + (per_sm + cap - 1) / cap - 1) / ((per_sm + cap - 1) / cap > 0 ? (per_sm + cap - 1) / cap : 1);
    long long ppb = per_sm <= 1536 ? per_cta : 256;
    if (ppb < 32) ppb = 32;
    if (ppb > c->n_particles) ppb = c->n_particles;
    return (int)ppb;
}

static int validate(const sem_pf_config *c) {
    if (!c) { set_error("null config"); return SEM_ERR_INVALID; }
    if (c->model < 0 || c->model > 3) { set_error("bad model"); return SEM_ERR_INVALID; }
    const int G = c->model >= SEM_MODEL_SIR_SUBGROUPS ? c->n_groups : 1;
    if (G < 1 || G > SEM_MAX_GROUPS) { set_error("n_groups must be 1..4"); return SEM_ERR_INVALID; }
    if (c->n_particles < 1 || c->n_obs < 1 || c->n_filters < 1) { set_error("bad sizes"); return SEM_ERR_INVALID; }
    const int C = model_cols(c->model, G);
    const int want = c->model == SEM_MODEL_SIR_SUBGROUPS2 ? 3 : C;
    if (c->n_obs_cols != want) { set_error("n_obs_cols does not match the model"); return SEM_ERR_INVALID; }
    return SEM_OK;
}

struct WsLayout { size_t L[2], pfx[2], scale[2], total[2], part, counter, wtab, bytes; int nb, ppb, wt_n; };

// total population = the largest count a compartment (or a group sum) can hold; 0 = no table (unknown, or > 1 GiB)
static int weight_table_n(const sem_pf_config *c) {
    const int G = c->model >= SEM_MODEL_SIR_SUBGROUPS ? c->n_groups : 1;
    double tot = 0;
    for (int g = 0; g < G; g++) tot += c->n_population[g];
    static int env_off = -1;
    if (env_off < 0) { const char *e = getenv("SEM_NO_WEIGHT_TABLE"); env_off = (e && e[0] == '1') ? 1 : 0; }
    if (env_off || c->obs_kind != SEM_OBS_BINOMIAL || !(tot >= 1) || tot != (double)(long long)tot || c->n_obs < 2) return 0;   // (the normal pdf is cheap)
    const double bytes = (double)(c->n_obs - 1) * c->n_obs_cols * (tot + 1) * sizeof(double);
    return bytes <= 1073741824.0 ? (int)tot : 0;
}
static WsLayout ws_layout(const sem_pf_config *c) {
    WsLayout w;
    w.ppb = choose_ppb(c);
    w.nb = (c->n_particles + w.ppb - 1) / w.ppb;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t F = c->n_filters;
    for (int i = 0; i < 2; i++) w.L[i] = take(F * c->n_particles * sizeof(double));
    for (int i = 0; i < 2; i++) w.pfx[i] = take(F * w.nb * sizeof(double));
    for (int i = 0; i < 2; i++) w.scale[i] = take(F * w.nb * sizeof(double));
    for (int i = 0; i < 2; i++) w.total[i] = take(F * sizeof(double));
    w.part = take(2 * F * w.nb * sizeof(double2));
    w.counter = take(F * sizeof(unsigned int));
    w.wt_n = weight_table_n(c);
    w.wtab = take(w.wt_n ? (size_t)(c->n_obs - 1) * c->n_obs_cols * ((size_t)w.wt_n + 1) * sizeof(double) : 0);
    w.bytes = off;
    return w;
}

template <class Model>
static void launch_model(const PfDev &P, int p, int arith, bool replay, dim3 grid, int threads, cudaStream_t s) {
    const size_t smem = P.pfx_in_smem ? (size_t)P.nb * sizeof(double) : 0;
    if (p == 0) {
        if (P.wtab) weight_table_kernel<<<sm_count() * 8, 256, 0, s>>>(P);
        pf_init<Model><<<grid, threads, 0, s>>>(P);
    }
    else if (replay) pf_step<Model, SEM_ARITH_REFERENCE, true><<<grid, threads, smem, s>>>(P, p);
    else if (arith == SEM_ARITH_REFERENCE) pf_step<Model, SEM_ARITH_REFERENCE, false><<<grid, threads, smem, s>>>(P, p);
    else if (arith == SEM_ARITH_UNIFORMIZED) pf_step<Model, SEM_ARITH_UNIFORMIZED, false><<<grid, threads, smem, s>>>(P, p);
    else if (arith == SEM_ARITH_FAST32) pf_step<Model, SEM_ARITH_FAST32, false><<<grid, threads, smem, s>>>(P, p);
    else if (arith == SEM_ARITH_UNIFORMIZED32) pf_step<Model, SEM_ARITH_UNIFORMIZED32, false><<<grid, threads, smem, s>>>(P, p);
    else pf_step<Model, SEM_ARITH_FAST, false><<<grid, threads, smem, s>>>(P, p);
}

}  // namespace sem

using namespace sem;

extern "C" {

const char *sem_last_error(void) { return err_buf(); }
int sem_abi_version(void) { return SEM_ABI_VERSION; }

int sem_device_info(int *sms, int *major, int *minor) {
    int dev = 0;
    SEM_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp pr;
    SEM_CUDA(cudaGetDeviceProperties(&pr, dev));
    if (sms) *sms = pr.multiProcessorCount;
    if (major) *major = pr.major;
    if (minor) *minor = pr.minor;
    return SEM_OK;
}

size_t sem_pf_workspace_bytes(const sem_pf_config *cfg) { return validate(cfg) ? 0 : ws_layout(cfg).bytes; }

static int hist_rows(const sem_pf_config *c) { return c->store_history ? c->n_obs : (c->n_obs < 2 ? c->n_obs : 2); }

size_t sem_pf_hist_elems(const sem_pf_config *c) {
    if (validate(c)) return 0;
    const int G = c->model >= SEM_MODEL_SIR_SUBGROUPS ? c->n_groups : 1;
    return (size_t)c->n_filters * hist_rows(c) * model_cols(c->model, G) * c->n_particles;
}
size_t sem_pf_ancestry_elems(const sem_pf_config *c) { return validate(c) ? 0 : (size_t)c->n_filters * hist_rows(c) * c->n_particles; }
int sem_pf_launch_count(const sem_pf_config *c);

static int fill_dev(const sem_pf_config *cfg, const sem_pf_buffers *buf, PfDev &P, WsLayout &w, bool &replay) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!buf || !buf->Y || !buf->theta || !buf->X_hist || !buf->ancestry || !buf->log_zetas || !buf->status || !buf->workspace) {
        set_error("null buffer"); return SEM_ERR_INVALID;
    }
    replay = buf->replay_ssa_u != nullptr;
    if (replay && (!buf->replay_resample_u || !buf->replay_ssa_off || !buf->X0)) { set_error("replay needs resample_u, ssa_off and X0"); return SEM_ERR_INVALID; }
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
    w = ws_layout(cfg);
    char *ws = (char *)buf->workspace;
    P.N = cfg->n_particles; P.T = cfg->n_obs; P.Cobs = cfg->n_obs_cols; P.obs_kind = cfg->obs_kind; P.resampler = cfg->resampler;
    P.nb = w.nb; P.ppb = w.ppb; P.hist_rows = hist_rows(cfg); P.model = cfg->model; P.n_filters = cfg->n_filters;
    P.ntheta = model_ntheta(cfg->model, G); P.init_poisson = buf->X0 == nullptr;
    P.pfx_in_smem = w.nb <= 4096;                           // 32 KB of dynamic shared memory at most
    P.probs = cfg->probs; P.dt = cfg->dt;
    P.key = make_philox_key(cfg->seed); P.filter_id0 = cfg->filter_id0;
    for (int g = 0; g < SEM_MAX_GROUPS; g++) { P.mu[g] = cfg->mu[g]; P.npop[g] = cfg->n_population[g]; }
    P.Y = buf->Y; P.theta = buf->theta; P.X0 = buf->X0;
    P.res_u = buf->replay_resample_u; P.ssa_u = buf->replay_ssa_u; P.ssa_off = (const long long *)buf->replay_ssa_off;
    P.X_hist = buf->X_hist; P.ancestry = buf->ancestry; P.status = buf->status; P.log_zetas = buf->log_zetas;
    P.n_events = (unsigned long long *)buf->n_events;
    for (int i = 0; i < 2; i++) {
        P.L[i] = (double *)(ws + w.L[i]); P.pfx[i] = (double *)(ws + w.pfx[i]);
        P.scale[i] = (double *)(ws + w.scale[i]); P.total[i] = (double *)(ws + w.total[i]);
    }
    P.part = (double2 *)(ws + w.part); P.counter = (unsigned int *)(ws + w.counter);
    // the table covers counts up to the configured population: only valid when X_0 is drawn from it (pmcmc.py:156-169)
    P.wt_n = (P.init_poisson && !replay) ? w.wt_n : 0;
    P.wtab = P.wt_n ? (double *)(ws + w.wtab) : nullptr;
    P.j0 = 0; P.sharded = 0; P.X_in = nullptr; P.summary = nullptr; P.split_main = 0;
    P.iter_out = buf->iteration_result; P.path_exact = (int)cfg->path_exact;
    if (P.iter_out && !cfg->store_history) { set_error("iteration_result needs store_history = 1"); return SEM_ERR_INVALID; }
    return SEM_OK;
}

static void launch_step(const sem_pf_config *cfg, const PfDev &P, const WsLayout &w, int p, bool replay, cudaStream_t s) {
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
    const int threads = (w.ppb + 31) / 32 * 32;
    const dim3 grid(w.nb, cfg->n_filters);
    switch (cfg->model) {
        case SEM_MODEL_SIR: launch_model<SirModel>(P, p, cfg->arith, replay, grid, threads, s); break;
#ifndef SEM_ONLY_SIR                                         /* (reduced build of the probe / variant tools) */
        case SEM_MODEL_SEIR: launch_model<SeirModel>(P, p, cfg->arith, replay, grid, threads, s); break;
        default:
            switch (G) {
                case 1: launch_model<SubModel<1>>(P, p, cfg->arith, replay, grid, threads, s); break;
                case 2: launch_model<SubModel<2>>(P, p, cfg->arith, replay, grid, threads, s); break;
                case 3: launch_model<SubModel<3>>(P, p, cfg->arith, replay, grid, threads, s); break;
                default: launch_model<SubModel<4>>(P, p, cfg->arith, replay, grid, threads, s); break;
            }
#else
        default: break;
#endif
    }
}

}  // extern "C"

template <class Model>
static const void *persistent_fn(int arith) {
    return arith == SEM_ARITH_REFERENCE ? (const void *)pf_persistent<Model, SEM_ARITH_REFERENCE>
           : arith == SEM_ARITH_FAST32  ? (const void *)pf_persistent<Model, SEM_ARITH_FAST32>
           : arith == SEM_ARITH_UNIFORMIZED32 ? (const void *)pf_persistent<Model, SEM_ARITH_UNIFORMIZED32>
                                        : (const void *)pf_persistent<Model, SEM_ARITH_FAST>;
}
static const void *persistent_kernel(const sem_pf_config *cfg) {
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
    switch (cfg->model) {
        case SEM_MODEL_SIR: return persistent_fn<SirModel>(cfg->arith);
#ifndef SEM_ONLY_SIR
        case SEM_MODEL_SEIR: return persistent_fn<SeirModel>(cfg->arith);
        default:
            switch (G) {
                case 1: return persistent_fn<SubModel<1>>(cfg->arith);
                case 2: return persistent_fn<SubModel<2>>(cfg->arith);
                case 3: return persistent_fn<SubModel<3>>(cfg->arith);
                default: return persistent_fn<SubModel<4>>(cfg->arith);
            }
#else
        default: return persistent_fn<SirModel>(cfg->arith);
#endif
    }
}

// Threads per CTA of the whole-filter kernel and its layout (*split_main): 0 = plain; > 0 = balanced -- when a CTA's
// ppb particles are W full warps per scheduler plus at most two more warps' worth, those extra particles are time-split
// between four helper warps (direct-method loops; see pf_persistent); < 0 = sorted by the candidate count (uniformized32,
// SIR / SEIR).  SEM_NO_SPLIT=1 keeps the plain layout.
static int persistent_threads(const sem_pf_config *cfg, const WsLayout &w, int *split_main) {
    *split_main = 0;
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1, C = model_cols(cfg->model, G);
    const bool legs = cfg->arith == SEM_ARITH_FAST32 || (cfg->arith == SEM_ARITH_FAST && C <= 6);
    const int e = w.ppb % 128, main_n = w.ppb - e;
    static int env_off = -1;
    if (env_off < 0) { const char *s = getenv("SEM_NO_SPLIT"); env_off = (s && s[0] == '1') ? 1 : 0; }
    if (cfg->arith == SEM_ARITH_UNIFORMIZED32 && !env_off && C <= 4 && w.ppb > 32) {
        const int nchunks = (w.ppb + 31) / 32, extra = nchunks % 4;          // sorted layout; helper legs for 4 W + 1 or + 2 chunks
        static int env_nh = -1;
        if (env_nh < 0) { const char *s = getenv("SEM_NO_HELPERS"); env_nh = (s && s[0] == '1') ? 1 : 0; }
        if (!env_nh && nchunks >= 4 && (extra == 1 || extra == 2) && (nchunks - extra + 4) * 32 <= kMaxThreads) {
            *split_main = -2;
            return (nchunks - extra + 4) * 32;
        }
        *split_main = -1;
    }
    if (legs && !env_off && main_n >= 128 && e > 0 && e <= 64 && main_n + 128 <= kMaxThreads) { *split_main = main_n; return main_n + 128; }
    return (w.ppb + 31) / 32 * 32;
}

// dynamic shared memory of the whole-filter kernel: pfx / scale of the CTAs, plus the sorted layout's exchange area
// (h, B, state out, state back, K, aux counter, home index per thread, 128 bins + range)
static size_t persistent_smem(const sem_pf_config *cfg, const WsLayout &w, int threads, int split_main) {
    size_t b = 2 * (size_t)w.nb * sizeof(double);
    if (split_main < 0) {
        const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1, C = model_cols(cfg->model, G);
        b += (size_t)threads * (2 * sizeof(double) + (2 * C + 3) * sizeof(int32_t)) + 132 * sizeof(uint32_t);
    }
    return b;
}

// opt in to more than 48 KB of shared memory per CTA where the exchange area needs it (once per kernel)
static int persistent_prepare(const void *fn, size_t smem) {
    static const void *done_fn[32];
    static size_t done_sz[32];
    static int n_done = 0;
    if (smem <= 24 * 1024) return SEM_OK;
    for (int i = 0; i < n_done; i++) if (done_fn[i] == fn && done_sz[i] >= smem) return SEM_OK;
    SEM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_done < 32) { done_fn[n_done] = fn; done_sz[n_done] = smem; n_done++; }
    return SEM_OK;
}

// One cooperative launch for the whole filter when every CTA can be co-resident (SEM_NO_PERSISTENT=1 or
// cfg->reserved = 1 forces the launch-per-step path; both give bit-identical results).
static bool use_persistent(const sem_pf_config *cfg, const WsLayout &w, bool replay) {
    if (replay || cfg->reserved == 1 || cfg->arith == SEM_ARITH_UNIFORMIZED || w.nb > 1024 || cfg->n_obs < 2) return false;
    static int env_off = -1;
    if (env_off < 0) { const char *e = getenv("SEM_NO_PERSISTENT"); env_off = (e && e[0] == '1') ? 1 : 0; }
    if (env_off) return false;
    int dev = 0, coop = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return false;
    int split_main;
    const int threads = persistent_threads(cfg, w, &split_main);
    const size_t smem = persistent_smem(cfg, w, threads, split_main);
    if (persistent_prepare(persistent_kernel(cfg), smem) != SEM_OK) { cudaGetLastError(); return false; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_kernel(cfg), threads, smem) != cudaSuccess) { cudaGetLastError(); return false; }
    return (long long)per_sm * sm_count() >= (long long)w.nb * cfg->n_filters;
}

extern "C" {

int sem_pf_launch_count(const sem_pf_config *c) {
    if (validate(c)) return 0;
    const WsLayout w = ws_layout(c);
    return use_persistent(c, w, false) ? 1 : c->n_obs + (w.wt_n ? 1 : 0);      // (+ the weight-table kernel; assumes X_0 drawn on the device)
}

int sem_pf_run(const sem_pf_config *cfg, const sem_pf_buffers *buf, void *stream) {
    PfDev P; WsLayout w; bool replay;
    int rc = fill_dev(cfg, buf, P, w, replay);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (use_persistent(cfg, w, replay)) {                    // one launch, nothing else (the kernel initialises its outputs)
        void *args[] = {(void *)&P};
        const dim3 grid(w.nb, cfg->n_filters), block(persistent_threads(cfg, w, &P.split_main));
        SEM_CUDA(cudaLaunchCooperativeKernel(persistent_kernel(cfg), grid, block, args, persistent_smem(cfg, w, block.x, P.split_main), s));
        return SEM_OK;
    }
    SEM_CUDA(cudaMemsetAsync(P.counter, 0, cfg->n_filters * sizeof(unsigned int), s));
    SEM_CUDA(cudaMemsetAsync(P.status, 0, cfg->n_filters * sizeof(int32_t), s));
    SEM_CUDA(cudaMemsetAsync(P.log_zetas, 0, (size_t)cfg->n_filters * cfg->n_obs * sizeof(double), s));   // zetas[0] = 1 (pmcmc.py:154)
    if (P.n_events) SEM_CUDA(cudaMemsetAsync(P.n_events, 0, cfg->n_filters * sizeof(unsigned long long), s));
    for (int p = 0; p < cfg->n_obs; p++) launch_step(cfg, P, w, p, replay, s);
    if (P.iter_out) {
        const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
        switch (model_cols(cfg->model, G)) {
            case 3: iteration_epilogue_kernel<3><<<cfg->n_filters, 32, 0, s>>>(P); break;
            case 4: iteration_epilogue_kernel<4><<<cfg->n_filters, 32, 0, s>>>(P); break;
            case 6: iteration_epilogue_kernel<6><<<cfg->n_filters, 32, 0, s>>>(P); break;
            case 9: iteration_epilogue_kernel<9><<<cfg->n_filters, 32, 0, s>>>(P); break;
            default: iteration_epilogue_kernel<12><<<cfg->n_filters, 32, 0, s>>>(P); break;
        }
    }
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

// ---- one shard of a particle-sharded filter; the host drives the steps and the exchanges
static int shard_dev(const sem_pf_config *cfg, const sem_pf_buffers *buf, int32_t particle_offset, double *summary,
                     PfDev &P, WsLayout &w) {
    bool replay;
    int rc = fill_dev(cfg, buf, P, w, replay);
    if (rc) return rc;
    if (replay || cfg->n_filters != 1 || !summary) { set_error("shard calls need n_filters = 1, Philox mode and a summary buffer"); return SEM_ERR_INVALID; }
    P.j0 = particle_offset; P.sharded = 1; P.summary = summary;
    return SEM_OK;
}

int sem_shard_init(const sem_pf_config *cfg, const sem_pf_buffers *buf, int32_t particle_offset, double *summary, void *stream) {
    PfDev P; WsLayout w;
    int rc = shard_dev(cfg, buf, particle_offset, summary, P, w);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    SEM_CUDA(cudaMemsetAsync(P.counter, 0, sizeof(unsigned int), s));
    SEM_CUDA(cudaMemsetAsync(P.status, 0, sizeof(int32_t), s));
    if (P.n_events) SEM_CUDA(cudaMemsetAsync(P.n_events, 0, sizeof(unsigned long long), s));
    launch_step(cfg, P, w, 0, false, s);
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

int sem_shard_offspring(const sem_pf_config *cfg, const sem_pf_buffers *buf, const sem_shard_step *st, int32_t *send_records,
                        void *stream) {
    PfDev P; WsLayout w; bool replay;
    int rc = fill_dev(cfg, buf, P, w, replay);
    if (rc) return rc;
    if (!st || !send_records || st->step < 1 || st->step >= cfg->n_obs) { set_error("bad shard step"); return SEM_ERR_INVALID; }
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1, C = model_cols(cfg->model, G);
    const int par = (st->step - 1) & 1, prow = (st->step - 1) % P.hist_rows;
    OffDev O;
    O.N = P.N; O.C = C; O.nb = w.nb; O.ppb = w.ppb; O.j0 = st->particle_offset;
    O.X = P.X_hist + (size_t)prow * C * P.N;
    O.L = P.L[par]; O.pfx = P.pfx[par]; O.scale = P.scale[par];
    O.total_local = st->total_local; O.G = st->G; O.G_next = st->G_next; O.s = st->s;
    O.sm.u0 = st->u0; O.sm.Nd = (double)st->n_global; O.sm.total = st->total; O.sm.N = st->n_global;
    O.slot0 = st->slot0; O.send = send_records;
    pf_offspring<<<w.nb, (w.ppb + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(O);
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

int sem_shard_propagate(const sem_pf_config *cfg, const sem_pf_buffers *buf, const sem_shard_step *st,
                        const int32_t *recv_records, double *summary, void *stream) {
    PfDev P; WsLayout w;
    int rc = shard_dev(cfg, buf, st ? st->particle_offset : 0, summary, P, w);
    if (rc) return rc;
    if (!st || !recv_records || st->step < 1 || st->step >= cfg->n_obs) { set_error("bad shard step"); return SEM_ERR_INVALID; }
    P.X_in = recv_records;
    launch_step(cfg, P, w, st->step, false, (cudaStream_t)stream);
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

#ifdef SEM_PHASES
int sem_debug_phases(unsigned long long *host_out) {
    SEM_CUDA(cudaMemcpyFromSymbol(host_out, g_phase, sizeof(unsigned long long) * 16 * 256));
    return SEM_OK;
}
int sem_debug_warps(unsigned long long *end_out, unsigned int *work_out) {
    SEM_CUDA(cudaMemcpyFromSymbol(end_out, g_warp_end, sizeof(unsigned long long) * 32 * 256));
    SEM_CUDA(cudaMemcpyFromSymbol(work_out, g_warp_work, sizeof(unsigned int) * 32 * 256));
    return SEM_OK;
}
#endif

int sem_path_sample(const int32_t *X_hist, const int32_t *ancestry, int32_t T, int32_t N, int32_t C, int32_t chosen,
                    int32_t exact, uint64_t seed, uint32_t filter_id, int32_t *traj, void *stream) {
    if (!X_hist || !ancestry || !traj || T < 1 || N < 1 || chosen >= N) { set_error("bad path_sample args"); return SEM_ERR_INVALID; }
    const PhiloxKey key = make_philox_key(seed);
    path_sample_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(X_hist, ancestry, T, N, C, chosen, exact, key, filter_id, traj);
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

int sem_hist_to_f64(const int32_t *X_hist, int32_t T, int32_t N, int32_t C, double *out, void *stream) {
    if (!X_hist || !out) { set_error("null"); return SEM_ERR_INVALID; }
    const size_t n = (size_t)T * N * C;
    const int blocks = (int)((n + 255) / 256 < (size_t)sm_count() * 16 ? (n + 255) / 256 : (size_t)sm_count() * 16);
    hist_to_f64_kernel<<<blocks ? blocks : 1, 256, 0, (cudaStream_t)stream>>>(X_hist, T, N, C, out);
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

int sem_pf_run_host(const sem_pf_config *cfg, const double *Y, const double *theta, const int32_t *X0,
                    double *log_zetas_out, double *zetas_out, double *hidden_out, double *ancestry_out, uint64_t *n_events_out) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (cfg->n_filters != 1) { set_error("sem_pf_run_host handles one filter"); return SEM_ERR_INVALID; }
    sem_pf_config c = *cfg;
    c.store_history = 1;
    const int G = c.model >= SEM_MODEL_SIR_SUBGROUPS ? c.n_groups : 1, C = model_cols(c.model, G), T = c.n_obs, N = c.n_particles;
    const size_t nh = sem_pf_hist_elems(&c), na = sem_pf_ancestry_elems(&c), wsb = sem_pf_workspace_bytes(&c);
    const size_t nY = (size_t)T * c.n_obs_cols, nth = model_ntheta(c.model, G);
    char *d = nullptr;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    size_t oY = 0, oth = oY + al(nY * 8), oX0 = oth + al(nth * 8), oH = oX0 + al(X0 ? (size_t)C * N * 4 : 0), oA = oH + al(nh * 4),
           oZ = oA + al(na * 4), oS = oZ + al((size_t)T * 8), oE = oS + 256, oW = oE + 256, oF = oW + al(wsb),
           tot = oF + al(((hidden_out ? nh : 0) > (ancestry_out ? na : 0) ? (hidden_out ? nh : 0) : (ancestry_out ? na : 0)) * 8);
    SEM_CUDA(cudaMalloc(&d, tot));
    cudaStream_t s = 0;
    rc = SEM_OK;
    auto fail = [&](cudaError_t e, const char *what) { if (e != cudaSuccess && rc == SEM_OK) { set_error("%s: %s", what, cudaGetErrorString(e)); rc = SEM_ERR_CUDA; } };
    fail(cudaMemcpyAsync(d + oY, Y, nY * 8, cudaMemcpyHostToDevice, s), "H2D Y");
    fail(cudaMemcpyAsync(d + oth, theta, nth * 8, cudaMemcpyHostToDevice, s), "H2D theta");
    if (X0) fail(cudaMemcpyAsync(d + oX0, X0, (size_t)C * N * 4, cudaMemcpyHostToDevice, s), "H2D X0");
    sem_pf_buffers b{};
    b.Y = (const double *)(d + oY); b.theta = (const double *)(d + oth); b.X0 = X0 ? (const int32_t *)(d + oX0) : nullptr;
    b.X_hist = (int32_t *)(d + oH); b.ancestry = (int32_t *)(d + oA); b.log_zetas = (double *)(d + oZ);
    b.status = (int32_t *)(d + oS); b.n_events = (uint64_t *)(d + oE); b.workspace = d + oW;
    if (rc == SEM_OK) rc = sem_pf_run(&c, &b, s);
    int32_t status = 0;
    if (rc == SEM_OK) {
        fail(cudaMemcpyAsync(&status, b.status, 4, cudaMemcpyDeviceToHost, s), "D2H status");
        if (log_zetas_out || zetas_out) {
            double *tmp = log_zetas_out ? log_zetas_out : zetas_out;
            fail(cudaMemcpyAsync(tmp, b.log_zetas, (size_t)T * 8, cudaMemcpyDeviceToHost, s), "D2H log_zetas");
        }
        if (n_events_out) fail(cudaMemcpyAsync(n_events_out, b.n_events, 8, cudaMemcpyDeviceToHost, s), "D2H n_events");
        if (hidden_out) {
            sem_hist_to_f64(b.X_hist, T, N, C, (double *)(d + oF), s);
            fail(cudaMemcpyAsync(hidden_out, d + oF, nh * 8, cudaMemcpyDeviceToHost, s), "D2H hidden");
        }
        if (ancestry_out) {
            i32_to_f64_kernel<<<sm_count() * 8, 256, 0, s>>>(b.ancestry, na, (double *)(d + oF));
            fail(cudaMemcpyAsync(ancestry_out, d + oF, na * 8, cudaMemcpyDeviceToHost, s), "D2H ancestry");
        }
        fail(cudaStreamSynchronize(s), "sync");
        if (rc == SEM_OK && zetas_out) {
            const double *src = log_zetas_out ? log_zetas_out : zetas_out;
            for (int i = T - 1; i >= 0; i--) zetas_out[i] = exp(src[i]);
        }
    }
    cudaFree(d);
    if (rc != SEM_OK) return rc;
    return status;
}

}  // extern "C"
