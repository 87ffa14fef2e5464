// sem_pf_dev.cuh -- device code of the bootstrap particle filter shared by the translation units of libsem_b200.so:
// sem_pf.cu (one GPU: whole-filter kernel pf_persistent, launch-per-step kernels) and sem_pf_xchg.cu (one filter
// sharded over the GPUs of a node: pf_persistent_x, whose resampling barrier and particle migration run over peer memory).
#pragma once
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
    const double *probs_f;     // [F] observation parameter per filter, or null = probs for every filter
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
__device__ __forceinline__ double column_logw(const PfDev &P, double y, double xc, const double2 *tab, const double probs) {
    return (P.obs_kind == SEM_OBS_BINOMIAL) ? binom_logpmf_obs(binom_obs(y, tab), xc, probs, tab) : norm_logpdf(y, xc, probs, tab);
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
        P.wtab[i] = (y != y) ? 0.0 : column_logw(P, y, (double)(i - pc * per_col), tab, P.probs);   // (no table with per-filter parameters)
    }
}

template <class Model>
__device__ __forceinline__ double particle_logw(const PfDev &P, const double *x, const double *Yrow, const double *wrow, const double2 *tab,
                                                const double probs) {
    double lw = CUDART_INF;
    // a particle holding a negative count (S0 = n_population - Poisson(mu) < 0, pmcmc.py:156-169) has
    // no weight -- scipy returns nan for binom.pmf(k, n < 0, p) and for a negative scale of norm.pdf, np.random.choice
    // refuses the weights and the reference's filter returns (None, None, None) (pmcmc.py:187-192): NaN here, which
    // collapses the filter at this step in the same way (and keeps negative propensities out of the interval simulation)
    bool neg = false;
#pragma unroll
    for (int c = 0; c < Model::C; c++) neg = neg || x[c] < 0.0;
    if (neg) return CUDART_NAN;
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
            else l = column_logw(P, y, xc, tab, probs);
            lw = (l < lw || l != l) ? l : lw;                // min over columns; NaN sticks
        }
    }
    return lw == CUDART_INF ? 0.0 : lw;                      // nothing observed at this time: weight 1
}

#ifdef SEM_PHASES
static __device__ unsigned long long g_phase[24 * 256];
#define PHASE(k) do { if (tid == 0 && b == 0 && p < 256) { g_phase[p * 24 + (k)] = (unsigned long long)clock64(); } } while (0)
static __device__ unsigned long long g_warp_end[256 * 32];          // CTA 0: when each warp left the SSA loop, and its work
static __device__ unsigned int g_warp_work[256 * 32];
#define WARP_END(work) do { const unsigned int wk_ = __reduce_max_sync(0xffffffffu, (unsigned int)(work)); \
    if ((tid & 31) == 0 && b == 0 && p < 256) { g_warp_end[p * 32 + (tid >> 5)] = (unsigned long long)clock64(); g_warp_work[p * 32 + (tid >> 5)] = wk_; } } while (0)
static __device__ unsigned long long g_cta_t[2 * 128 * 160];          // per step and CTA: globaltimer (ns) after the SSA phase / at the publication
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CTA_T(k) do { if (tid == 0 && p < 128 && b < 160) g_cta_t[((k) * 128 + p) * 160 + b] = gtimer(); } while (0)
#else
#define PHASE(k)
#define WARP_END(work)
#define CTA_T(k)
#endif

// Weigh the CTA's particles against Y[p] and CTA-local scan: writes L[par] and the CTA partial (m_b, s_b) -- or, for the
// peer-memory exchange (KEEP), hands this thread's inclusive scan value and the CTA partial back in registers.
struct LocalScan { double incl, mb, sb; };
template <class Model, bool KEEP = false>
__device__ __forceinline__ LocalScan weigh_local(const PfDev &P, const int p, const int f, const int b, const int tid,
                                                 const bool active, const int j, const double *x, double *sm,
                                                 const double2 *tab) {
    const int N = P.N, par = p & 1;
    double lw = -CUDART_INF;
    if (active) {
        lw = particle_logw<Model>(P, x, P.Y + (size_t)p * P.Cobs, P.wtab ? P.wtab + (size_t)p * P.Cobs * (P.wt_n + 1) : nullptr, tab,
                                  P.probs_f ? __ldg(&P.probs_f[f]) : P.probs);
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
    if constexpr (!KEEP) {
        if (active) P.L[par][(size_t)f * N + j] = incl;
        if (tid == 0) P.part[((size_t)par * P.n_filters + f) * P.nb + b] = make_double2(mb, sb);
    }
    return LocalScan{incl, mb, sb};
}

// Combine nb CTA partials (m_i, s_i) into the global max M, per-CTA (exclusive prefix, scale = exp(m_i - M)) and the total.
// ONE summation order whatever the CTA size: entries are scanned in groups of 32 (warp scan), the 32 group totals of a
// block of 1024 entries by another warp scan, blocks of 1024 chained by a running carry -- so the whole-filter kernels,
// the launch-per-step path and the sharded filter produce bit-identical cdfs for the same partials.  load(i, m, s) reads
// entry i (global or shared memory); pfx_out / scale_out may be shared or global memory.  Called by every thread of a CTA.
template <class Load>
__device__ __forceinline__ void combine_canonical(const int nb, const int tid, double *sm, Load load, double *pfx_out,
                                                  double *scale_out, double &M_out, double &total_out) {
    const int nwarps = (blockDim.x + 31) >> 5, warp = tid >> 5, lane = tid & 31;
    double M = -CUDART_INF;
    for (int i = tid; i < nb; i += blockDim.x) { double m, s; load(i, m, s); M = fmax(M, m); }
    M = block_max(M, sm, tid, nwarps);
    const bool finiteM = (M > -CUDART_INF && M < CUDART_INF);
    double carry = 0.0;
    for (int base = 0; base < nb; base += 1024) {
        __syncthreads();                                     // sm[] is reused (block_max above, the previous block's totals)
        for (int gi = warp; gi < 32; gi += nwarps) {         // group gi of this block: one warp, one entry per lane
            const int i = base + gi * 32 + lane;
            double sc = 0.0, val = 0.0;
            if (i < nb && finiteM) {
                double m, sb;
                load(i, m, sb);
                sc = (m > -CUDART_INF) ? exp(m - M) : 0.0;
                val = sc * sb;
            }
            const double v = warp_incl_scan_d(val, lane);
            if (i < nb) { pfx_out[i] = v; scale_out[i] = sc; }   // (inclusive within the group, completed below)
            if (lane == 31) sm[gi] = v;
        }
        __syncthreads();
        const double ws = warp_incl_scan_d(sm[lane], lane);  // inclusive scan of the 32 group totals
        for (int gi = warp; gi < 32; gi += nwarps) {
            const int i = base + gi * 32 + lane;
            const double off = __shfl_sync(0xffffffffu, ws, gi > 0 ? gi - 1 : 0);
            if (i < nb) {
                double m, sb;
                load(i, m, sb);
                const double sc = scale_out[i], val = finiteM ? sc * sb : 0.0, v = pfx_out[i];
                pfx_out[i] = carry + ((gi > 0 ? v + off : v) - val);
            }
        }
        carry = carry + __shfl_sync(0xffffffffu, ws, 31);
    }
    M_out = M; total_out = carry;
}

// Combine the CTA partials of parity `par` (global memory) into (M, total) and per-CTA (prefix, scale) written to
// pfx_out/scale_out (shared or global memory).  Called by every thread of a CTA.
__device__ __forceinline__ void combine_partials(const PfDev &P, const int f, const int par, const int tid, double *sm,
                                                 double *pfx_out, double *scale_out, double &M_out, double &total_out) {
    const int nwarps = (blockDim.x + 31) >> 5;
    const double2 *part = P.part + ((size_t)par * P.n_filters + f) * P.nb;
    if (P.nb <= (int)blockDim.x) {                           // one partial per thread: a single L2 round trip (same arithmetic
        double2 ps = make_double2(-CUDART_INF, 0.0);         // as combine_canonical: warp scans + a scan of the warp totals)
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
    combine_canonical(P.nb, tid, sm, [&](int i, double &m, double &s) { const double2 v = __ldcg(&part[i]); m = v.x; s = v.y; },
                      pfx_out, scale_out, M_out, total_out);
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

// End of one PMCMC iteration (pmcmc.py:371 particle_path_sampler + the three small results the MH loop reads), by ONE
// WARP per filter after everything else of the filter is globally visible.  Same draw and indexing as
// path_sample_kernel.  The ancestry chase is a chain of T dependent loads (L2 or DRAM latency each: the history of the
// headline filter is 160 MB); lane 0 walks it alone, parking the indices in the first column of the output, and the
// 3 T state loads -- which depended on the chain when one thread did everything, 0.18 ms of a 4.5 ms pass -- are then
// issued by all lanes at once.  Call with the 32 lanes of a warp converged.
template <int C>
__device__ void iteration_epilogue(const PfDev &P, const int f) {
    const int T = P.T, N = P.N, lane = threadIdx.x & 31;
    double *out = P.iter_out + (size_t)f * (SEM_ITER_HEADER + (size_t)T * C);
    const int status = *(volatile int32_t *)&P.status[f];
    const int32_t *X = P.X_hist + (size_t)f * P.hist_rows * C * N, *A = P.ancestry + (size_t)f * P.hist_rows * N;
    const uint4 w = philox4x32_10(0u, 0u, 0u, stream_word(DOM_PATH, P.filter_id0 + f), P.key);
    int chosen = min((int)((bits_to_d12(w.x, w.y) - 1.0) * (double)N), N - 1);          // np.random.randint(0, N) (pmcmc.py:241)
    double *traj = out + SEM_ITER_HEADER;
    if (lane == 0) {
        out[0] = __ldcg(&P.log_zetas[(size_t)f * T + T - 1]);
        out[1] = (double)status;
        out[2] = P.n_events ? (double)__ldcg(&P.n_events[f]) : 0.0;
        out[3] = status != 0 ? -1.0 : (double)chosen;
        if (status == 0) {
            traj[(size_t)(T - 1) * C] = (double)chosen;
            for (int p = T - 2; p >= 0; p--) {
                chosen = __ldcg(&A[(size_t)(P.path_exact ? p + 1 : p) * N + chosen]);   // reference indexes row p
                traj[(size_t)p * C] = (double)chosen;
            }
        }
    }
    __syncwarp();
    if (status != 0) return;
    for (int p = lane; p < T; p += 32) {
        const int idx = (int)*(volatile double *)&traj[(size_t)p * C];
        double v[C];
#pragma unroll
        for (int c = 0; c < C; c++) v[c] = (double)__ldcg(&X[((size_t)p * C + c) * N + idx]);
#pragma unroll
        for (int c = 0; c < C; c++) traj[(size_t)p * C + c] = v[c];
    }
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

// J(c) without divisions in the common case: the real-valued slot position x = c N / total - u0 is evaluated with a
// Newton reciprocal; unless x lies within 1e-5 of an integer (error of x: a few ulp of N <= 2^31, i.e. < 1e-6) its
// ceiling IS the exact answer of first_slot_ge, which is only called for the rare near-ties.  n_over_total = N * rcp(total).
__device__ __forceinline__ long long first_slot_ge_quick(const SlotMap &sm, const double c, const double n_over_total) {
    if (c >= sm.total) return sm.N;
    const double xr = __fma_rn(c, n_over_total, -sm.u0), up = ceil(xr);
    if (fabs(xr - rint(xr)) < 1e-5) return first_slot_ge(sm, c);
    return up <= 0.0 ? 0 : (long long)up;
}

// ---------------------------------------------------------------------------------------------- peer-memory exchange
// One filter sharded over the W GPUs of a node, every rank running pf_persistent_x (sem_pf_xchg.cu).  Rank r owns the
// particles [r N, (r+1) N) and the global CTAs [r nb, (r+1) nb); resampling is global and systematic.  Two exchanges per
// observation step, both "data is the flag" (no separate signal, no fence on the critical path of the data):
//   partials  every CTA stores its (m_b, s_b) into EVERY rank's table part[gen % 3][W nb] (one 16-byte store per
//             rank, over NVLink for the peers); every CTA of every rank polls its own rank's table until all W nb
//             entries differ from the all-ones sentinel -- this IS the resampling barrier (it replaces grid.sync()) --
//             and combines them itself, so every CTA on every GPU holds bit-identical prefixes / total.
//   records   each particle computes the slots [J(lower), J(upper)) of its children and stores one record
//             (state, global parent index) per child straight into the receive buffer rec[gen & 1][N] of the rank that owns
//             the slot; the child's thread polls its own record until no word is the all-ones sentinel, then resets it.
//             Words travel with the sign bit flipped (rec_enc): a state count may be NEGATIVE -- the reference draws
//             I0 ~ Poisson(mu) and sets S0 = n_population - I0 without a clamp (pmcmc.py:156-169) -- so the sign cannot mark
//             an empty word; all-ones would be the count 2^31 - 1, which the int32 history cannot hold anyway.
// Re-use is safe without further synchronisation: a slot written in generation g is next written in generation g + 2
// (records) / g + 3 (partials), and nobody can get there before its reader has published generation g + 1, which it
// does after a system-scope fence that follows its reset (see the order of operations in pf_persistent_body).
// Generations keep counting across launches (gen0), so a launch needs no memset and no host barrier.
#define SEM_MAX_RANKS 8
struct XchgDev {
    int W, rank, NB, kper;          // ranks, this rank, global CTA count W nb, partials combined per thread
    long long Ng;                   // global particle count W N
    unsigned int gen0, tag;         // generation of this launch's step 0; launch tag of the path-sampler tokens (1..4095)
    long long timeout;              // spin limit in clock64 ticks: a lost peer ends the launch with status SEM_ERR_PEER
    double2 *part[SEM_MAX_RANKS];   // [3][NB]   per rank
    int32_t *rec[SEM_MAX_RANKS];    // [2][N][RW]
    unsigned long long *mail[SEM_MAX_RANKS];   // path-sampler token, one slot per rank
    double *iter[SEM_MAX_RANKS];    // [SEM_ITER_HEADER + T*C] packed result of the iteration, one copy per rank (or null)
    int *err;                       // this rank's error flag (1 = a peer did not answer)
    size_t filter_stride;           // one rank, several filters (grid y): filter f uses the arena at f * filter_stride bytes
};

__device__ __forceinline__ void st_vol(double2 *p, double a, double b) {
    asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void ld_vol(const double2 *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_vol(int32_t *p, int a, int b, int c, int d) {
    asm volatile("st.volatile.global.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int4 ld_vol(const int32_t *p) {
    int4 v;
    asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
constexpr unsigned long long kPartSentinel = ~0ull;
template <int C> struct RecWords { static constexpr int value = (C + 1 + 3) & ~3; };
__device__ __forceinline__ int rec_enc(const int v) { return v ^ (int)0x80000000u; }   // (its own inverse)

template <class T> __device__ __forceinline__ T *xsh(T *p, const size_t fsh) { return (T *)((char *)p + fsh); }   // pointer into filter f's arena
__device__ __forceinline__ void xchg_fail(const XchgDev &X, const size_t fsh) { *(volatile int *)xsh(X.err, fsh) = 1; }
__device__ __forceinline__ bool xchg_failed(const XchgDev &X, const size_t fsh) { return *(volatile int *)xsh(X.err, fsh) == 1; }   // (an arena marked empty holds all-ones)

// Publish this CTA's partial of generation `gen` to every rank: the first W lanes of the CTA's LAST warp, one destination
// each.  The same lanes ran xchg_early_fence after this step's resets (records and partial table; made visible to them
// by a CTA barrier): fence.sys ... store by one thread is the release that orders the resets before the publication.
// The fence costs ~3 us on B200, so it is issued right after the resets -- the last warp is a second-leg helper that
// waits for its hand-over (balanced / helper layouts) or carries the CTA's lightest chunk (sorted layout) -- and not
// between the weights and the publication, where every CTA of every rank would wait for it.
__device__ __forceinline__ bool xchg_is_publisher(const XchgDev &X, const int tid) {
    return (tid >> 5) == (((int)blockDim.x - 1) >> 5) && (tid & 31) < X.W;
}
__device__ __forceinline__ void xchg_early_fence(const XchgDev &X, const int tid) {
    if (xchg_is_publisher(X, tid)) {
        if (X.W == 1) __threadfence();                       // one rank: every access comes from this GPU
        else __threadfence_system();
    }
}
__device__ __forceinline__ void xchg_publish(const XchgDev &X, const size_t fsh, const unsigned gen, const int gb, const int tid, const double mb, const double sb) {
    if (xchg_is_publisher(X, tid)) st_vol(xsh(X.part[tid & 31], fsh) + (size_t)(gen % 3u) * X.NB + gb, mb, sb);
}

// Wait for all NB partials of generation `gen` (the resampling barrier) and combine them: global max M, per-CTA
// (exclusive prefix, scale) into shared memory, total.  Thread t polls the entries t, t + blockDim, ... and stages them
// in shared memory (stage[0..NB) = m, stage[NB..2NB) = s); the combine is combine_partials' arithmetic.  Also resets this
// CTA's share of the table of generation gen + 2.  Called by every thread of the CTA.
__device__ __forceinline__ void xchg_wait_combine(const XchgDev &X, const size_t fsh, const unsigned gen, const int b, const int nb, const int tid, double *sm,
                                                  double *s_pfx, double *s_scale, double *stage, double &M_out, double &total_out, const int p) {
    const int nwarps = (blockDim.x + 31) >> 5, NB = X.NB;
    const double2 *tab = xsh(X.part[X.rank], fsh) + (size_t)(gen % 3u) * NB;
    const long long t0 = clock64();
    unsigned spins = 0;
    double m0 = -CUDART_INF, s0 = 0.0;
    for (int i = tid; i < NB; i += blockDim.x) {
        double m = -CUDART_INF, sv = 0.0;
        for (;;) {
            unsigned long long a, c;
            ld_vol(tab + i, a, c);
            if (a != kPartSentinel && c != kPartSentinel) { m = __longlong_as_double((long long)a); sv = __longlong_as_double((long long)c); break; }
            if ((++spins & 1023u) == 0u && clock64() - t0 > X.timeout) { xchg_fail(X, fsh); break; }
        }
        stage[i] = m; stage[NB + i] = sv;
        if (i == tid) { m0 = m; s0 = sv; }
    }
    PHASE(16);
    if (NB <= (int)blockDim.x) {                             // one partial per thread, straight from the registers
        const double M = block_max(m0, sm, tid, nwarps);
        const bool finiteM = (M > -CUDART_INF && M < CUDART_INF);
        const double sc = (tid < NB && finiteM && m0 > -CUDART_INF) ? exp(m0 - M) : 0.0, val = sc * s0;
        double tot;
        const double incl = block_incl_scan(val, sm, tid, nwarps, &tot);
        if (tid < NB) { s_pfx[tid] = 0.0 + (incl - val); s_scale[tid] = sc; }
        M_out = M; total_out = 0.0 + tot;
    } else {
        combine_canonical(NB, tid, sm, [&](int i, double &m, double &sv) { m = stage[i]; sv = stage[NB + i]; }, s_pfx, s_scale, M_out, total_out);
    }
    PHASE(17);
    // every thread of this CTA has left its poll (CTA barriers above): all CTAs of all ranks have published generation gen,
    // hence finished reading generation gen - 1, whose table is the one generation gen + 2 will use
    if (tid < X.W) st_vol(xsh(X.part[X.rank], fsh) + (size_t)((gen + 2u) % 3u) * NB + tid * nb + b, __longlong_as_double(-1ll), __longlong_as_double(-1ll));
}

// Children of the CTA's particles (global systematic resampling in offspring form): particle i of global CTA gb owns the
// slots [J(lower_i), J(upper_i)), upper_i = min(pf + sc L_i, prefix of the next CTA) and lower_i = upper_{i-1} -- exactly
// the particle the search of select_ancestor finds for those slots.  One record per child goes to the owner of the slot.
template <class Model>
__device__ __forceinline__ void xchg_offspring(const PfDev &P, const XchgDev &X, const size_t fsh, const unsigned gen, const int p_next, const uint32_t fid,
                                               const int b, const int gb, const int tid, const int pidx, const bool active,
                                               const int j, const double *x, const double incl, const double total,
                                               const double *s_pfx, const double *s_scale, int *s_J) {
    constexpr int C = Model::C, RW = RecWords<C>::value;
    const int N = P.N, len = min(P.ppb, N - b * P.ppb), lane = tid & 31;
    SlotMap smap;
    smap.u0 = systematic_u0(P, p_next, fid); smap.Nd = (double)X.Ng; smap.total = total; smap.N = X.Ng;
    const double pf = s_pfx[gb], sc = s_scale[gb];
    const double cta_hi = (gb == X.NB - 1) ? total : s_pfx[gb + 1];
    const double n_over_total = __dmul_rn(smap.Nd, rcp_nr(total));
    long long hi = 0;
    if (active) {
        const double up = (pidx == len - 1) ? cta_hi : fmin(__fma_rn(sc, incl, pf), cta_hi);
        hi = first_slot_ge_quick(smap, up, n_over_total);
        s_J[pidx] = (int)hi;
    }
    __shared__ int s_J0;
    if (tid == (int)blockDim.x - 1) s_J0 = (int)first_slot_ge_quick(smap, pf, n_over_total);
    __syncthreads();
    long long lo = 0;
    if (active) lo = pidx == 0 ? s_J0 : s_J[pidx - 1];
    if (!active || hi < lo) hi = lo;
    int w[RW];
#pragma unroll
    for (int c = 0; c < RW; c++) w[c] = rec_enc(c < C ? (active ? (int)x[c] : 0) : 0);
    w[C] = rec_enc(P.j0 + j);
    const size_t half = (size_t)((gen + 1u) & 1u) * N * RW;
    auto put = [&](const long long ch, const int (&rec)[RW]) {
        const int r = (int)(ch / N);
        int32_t *dst = xsh(X.rec[r], fsh) + half + (size_t)(ch - (long long)r * N) * RW;
#pragma unroll
        for (int q = 0; q < RW; q += 4) st_vol(dst + q, rec[q], rec[q + 1], rec[q + 2], rec[q + 3]);
    };
    // the first two children by the particle's own thread; a particle with more (weight degeneracy) is served by its warp
    if (hi > lo) put(lo, w);
    if (hi > lo + 1) put(lo + 1, w);
    unsigned many = __ballot_sync(0xffffffffu, hi > lo + 2);
    while (many) {
        const int src = __ffs(many) - 1;
        many &= many - 1;
        int rec[RW];
#pragma unroll
        for (int c = 0; c < RW; c++) rec[c] = __shfl_sync(0xffffffffu, w[c], src);
        const long long l2 = __shfl_sync(0xffffffffu, lo, src) + 2, h2 = __shfl_sync(0xffffffffu, hi, src);
        for (long long ch = l2 + lane; ch < h2; ch += 32) put(ch, rec);
    }
}

// The child's side: wait for the record of local slot j in generation `gen`, reset it, return state and parent index.
template <class Model>
__device__ __forceinline__ int xchg_take_record(const PfDev &P, const XchgDev &X, const size_t fsh, const unsigned gen, const int j, double *x) {
    constexpr int C = Model::C, RW = RecWords<C>::value;
    int32_t *src = xsh(X.rec[X.rank], fsh) + ((size_t)(gen & 1u) * P.N + j) * RW;
    int w[RW];
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int q = 0; q < RW; q += 4) {
            const int4 v = ld_vol(src + q);
            w[q] = v.x; w[q + 1] = v.y; w[q + 2] = v.z; w[q + 3] = v.w;
            ok = ok && (max(max((unsigned)v.x, (unsigned)v.y), max((unsigned)v.z, (unsigned)v.w)) != 0xffffffffu);
        }
        if (ok) break;
        if ((++spins & 1023u) == 0u && clock64() - t0 > X.timeout) {
            xchg_fail(X, fsh);
#pragma unroll
            for (int c = 0; c < RW; c++) w[c] = rec_enc(0);      // an extinct particle: the launch ends quickly, status SEM_ERR_PEER
            break;
        }
    }
#pragma unroll
    for (int q = 0; q < RW; q += 4) st_vol(src + q, -1, -1, -1, -1);
#pragma unroll
    for (int c = 0; c < C; c++) x[c] = (double)rec_enc(w[c]);
    return rec_enc(w[C]);
}

// particle_path_sampler (pmcmc.py:236-248) over the shards, by warp 0 of CTA 0 of every rank after its grid barrier.
// The lineage is chased by whichever rank owns the current particle (local loads only); when the parent lives on
// another rank a token (time, global index) goes to that rank's mailbox.  The holder walks its SEGMENT of the lineage
// with lane 0 alone (a chain of dependent ancestry loads, indices parked in its own copy of the result), then all lanes
// fetch the segment's rows at once and store them into EVERY rank's packed iteration result, so all ranks end with the
// same trajectory and can run the MH accept step redundantly (no host collective per iteration).  Every lane fences its
// own remote stores (system scope) before lane 0 passes the token on.  Call with the 32 lanes of the warp converged.
template <int C>
__device__ void xchg_iteration_epilogue(const PfDev &P, const XchgDev &X, const int f, const size_t fsh) {
    const int T = P.T, N = P.N, W = X.W, me = X.rank, lane = threadIdx.x & 31;
    const size_t ish = (size_t)f * (SEM_ITER_HEADER + (size_t)T * C);    // this filter's block of the packed results (one rank)
    double *out = P.iter_out + ish;
    const int status = *(volatile int32_t *)&P.status[f];
    const uint4 wd = philox4x32_10(0u, 0u, 0u, stream_word(DOM_PATH, P.filter_id0 + f), P.key);
    long long idx = min((long long)((bits_to_d12(wd.x, wd.y) - 1.0) * (double)X.Ng), X.Ng - 1);   // np.random.randint(0, N) (pmcmc.py:241)
    if (lane == 0) {
        out[0] = __ldcg(&P.log_zetas[(size_t)f * T + T - 1]);
        out[1] = (double)status;
        out[2] = P.n_events ? (double)__ldcg(&P.n_events[f]) : 0.0;
        out[3] = status != 0 ? -1.0 : (double)idx;
    }
    if (status != 0) return;
    const unsigned long long tag = (unsigned long long)X.tag << 52, kDone = 0xFFFFFull;
    volatile unsigned long long *mail = xsh(X.mail[me], fsh);
    volatile double *park = X.iter[me] + ish + SEM_ITER_HEADER;           // row p, column 0: local index of the lineage at p
    int cur = T - 1;
    bool hold = (int)(idx / N) == me;
    const long long t0 = clock64();
    for (;;) {
        int seg_hi = 0, seg_lo = 0, next_owner = -2;                      // -2: nothing to do (failed / done), -1: lineage complete
        long long next_idx = 0;
        if (lane == 0) {
            bool go = true;
            if (!hold) {
                unsigned long long tok = 0;
                unsigned spins = 0;
                for (;;) {
                    tok = *mail;
                    if ((tok >> 52) == X.tag) break;
                    if ((++spins & 1023u) == 0u && clock64() - t0 > X.timeout) { xchg_fail(X, fsh); go = false; break; }
                }
                if (go) {
                    *mail = 0ull;
                    const unsigned long long pp = (tok >> 32) & kDone;
                    if (pp == kDone) go = false;
                    else { cur = (int)pp - 1; idx = (long long)(tok & 0xffffffffull); }
                }
            }
            if (go) {                                                     // walk this rank's segment of the lineage
                seg_hi = cur;
                long long id = idx;
                int p = cur;
                for (;;) {
                    const int loc = (int)(id - (long long)me * N);
                    park[(size_t)p * C] = (double)loc;
                    if (p == 0) { next_owner = -1; seg_lo = 0; break; }
                    const long long nid = __ldcg(&P.ancestry[((size_t)f * P.hist_rows + (P.path_exact ? p : p - 1)) * N + loc]);   // reference indexes row p
                    p -= 1;
                    if ((int)(nid / N) != me) { next_owner = (int)(nid / N); next_idx = nid; seg_lo = p + 1; break; }
                    id = nid;
                }
            }
        }
        next_owner = __shfl_sync(0xffffffffu, next_owner, 0);
        if (next_owner == -2) return;
        seg_hi = __shfl_sync(0xffffffffu, seg_hi, 0); seg_lo = __shfl_sync(0xffffffffu, seg_lo, 0);
        __syncwarp();                                                     // lane 0's parked indices are visible to the warp
        for (int q = seg_lo + lane; q <= seg_hi; q += 32) {               // the segment's rows, all lanes at once
            const int loc = (int)park[(size_t)q * C];
            double v[C];
#pragma unroll
            for (int c = 0; c < C; c++) v[c] = (double)__ldcg(&P.X_hist[(((size_t)f * P.hist_rows + q) * C + c) * N + loc]);
            for (int r = 0; r < W; r++) {
#pragma unroll
                for (int c = 0; c < C; c++) *(volatile double *)&X.iter[r][ish + SEM_ITER_HEADER + (size_t)q * C + c] = v[c];
            }
        }
        __threadfence_system();                                           // (keeps the rows stored so far ahead of the token, transitively ahead of DONE)
        __syncwarp();
        if (next_owner == -1) {
            if (lane == 0) for (int r = 0; r < W; r++) if (r != me) *(volatile unsigned long long *)X.mail[r] = tag | (kDone << 32);
            return;
        }
        if (lane == 0) {
            cur = seg_lo - 1;
            *(volatile unsigned long long *)X.mail[next_owner] = tag | ((unsigned long long)(cur + 1) << 32) | (unsigned long long)next_idx;
            hold = false;
        }
    }
}

// Whole filter in ONE cooperative launch (one CTA per SM, all co-resident): the resampling barrier of every step is
// a grid.sync(); after it every CTA combines the nb CTA partials itself (nb <= 1024 values, redundantly) instead of
// waiting for a "last CTA" and a new launch.  Same arithmetic as pf_init + pf_step, bit-identical results.
//
// PUSH (pf_persistent_x, one filter sharded over the GPUs of a node; Xp = the peer-memory exchange): the barrier is the
// arrival of all W nb partials in this rank's table, and instead of searching for its ancestor a thread waits for the
// record its parent stored for it (xchg_* above).  Per generation a CTA does, in this order:
//   take + reset its records | SSA, store, weigh, CTA scan | [CTA barriers] fence, publish its partial | wait for all
//   partials, reset its share of the table two generations ahead, combine | store its children's records.
template <class Model, int ARITH, bool PUSH>
__device__ __forceinline__ void pf_persistent_body(const PfDev &P, const XchgDev *Xp) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    // One rank may run several filters side by side (grid y; a batch of Metropolis-Hastings proposals): filter f has its
    // own arena, fsh bytes further on, so the filters never wait for each other -- the only grid-wide barriers are the
    // two outside the step loop.  (P and *Xp stay kernel parameters: constant-bank operands, no registers.)
    const size_t fsh = (PUSH && gridDim.y > 1) ? (size_t)blockIdx.y * Xp->filter_stride : 0;
    extern __shared__ __align__(16) double s_dyn[];          // pfx[nb], scale[nb] of the previous step (+ the sorted layout's exchange area)
    const int nbt = PUSH ? Xp->NB : P.nb;                    // CTAs of the whole filter
    const int soff = PUSH ? 4 * nbt : 2 * nbt;               // doubles before the sorted layout's area (PUSH: + the staged partials)
    double *s_pfx = s_dyn, *s_scale = s_dyn + nbt;
    __shared__ int s_J[PUSH ? kMaxThreads : 1];
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
    } else if (PUSH) grid.sync();                            // (orders CTA 0's output initialisation; the loop has no grid barrier)
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
    LocalScan ls{0.0, 0.0, 0.0};
    const int gb = PUSH ? Xp->rank * P.nb + b : b;           // this CTA in the whole filter
    if (P.T > 1) {
        ls = weigh_local<Model, PUSH>(P, 0, f, b, tid, active, j, x, sm, s_tab);
        if constexpr (PUSH) xchg_publish(*Xp, fsh, Xp->gen0, gb, tid, ls.mb, ls.sb);
    }
    bool dead = false;
    double lz = 0.0;
    unsigned long long my_pairs = 0;
    for (int p = 1; p < P.T; p++) {
        PHASE(0);
        if constexpr (!PUSH) grid.sync();                    // the resampling barrier (grid-wide fence + barrier)
        PHASE(1);
        if (dead) { if constexpr (PUSH) break; else continue; }
        const int par = p & 1;
        const int row = p % P.hist_rows, prow = (p + P.hist_rows - 1) % P.hist_rows;
        double M, total;
        if constexpr (PUSH) xchg_wait_combine(*Xp, fsh, Xp->gen0 + (unsigned)(p - 1), b, P.nb, tid, sm, s_pfx, s_scale, s_dyn + 2 * nbt, M, total, p);
        else combine_partials(P, f, par ^ 1, tid, sm, s_pfx, s_scale, M, total);
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
                uint32_t *bins = (uint32_t *)(s_dyn + soff + 2 * blockDim.x) + (2 * Model::C + 3) * blockDim.x;
                for (int i = tid; i < 130; i += blockDim.x) bins[i] = i == 129 ? 0xffffffffu : 0u;
            }
        }
        __syncthreads();                                     // s_pfx / s_scale complete
        // zetas[p] = zetas[p-1] * mean(w) (pmcmc.py:183), off the CTA's critical path: by the last thread, whose warp is
        // a second-leg helper waiting for its hand-over in the balanced layout
        if (b == 0 && tid == (int)blockDim.x - 1) {
            lz = lz + M + log(total) - log(PUSH ? (double)Xp->Ng : (double)N);
            P.log_zetas[(size_t)f * P.T + p] = lz;
        }
        PHASE(2);
        if constexpr (PUSH)                                  // resample, offspring form: one record per child, to the slot's owner
            xchg_offspring<Model>(P, *Xp, fsh, Xp->gen0 + (unsigned)(p - 1), p, fid, b, gb, tid, pidx, active, j, x, ls.incl, total, s_pfx, s_scale, s_J);
        PHASE(18);
        long long pairs = 0;
        int32_t *Xr = Xf + (size_t)row * Model::C * N;
        Model m;
        PairSource<false> src;
        if (starts) {
            if constexpr (PUSH) {
                const int a = xchg_take_record<Model>(P, *Xp, fsh, Xp->gen0 + (unsigned)p, j, x);   // (state, global parent) stored by the parent
                PHASE(6);
                Af[(size_t)row * N + j] = a;
            } else {
                const int a = select_ancestor<false>(P, p, f, j, fid, s_pfx, s_scale, total);
                PHASE(6);
                Af[(size_t)row * N + j] = a;
                const int32_t *Xq = Xf + (size_t)prow * Model::C * N;
#pragma unroll
                for (int c = 0; c < Model::C; c++) x[c] = (double)__ldcg(&Xq[(size_t)c * N + a]);   // written by other CTAs: L2, not L1
            }
            m.setup(P.theta + (size_t)f * P.ntheta, x);
            src.init(P.key, (uint32_t)(P.j0 + j), (uint32_t)p, stream_word(DOM_SSA, fid));
            PHASE(7);
        }
        if constexpr (PUSH) {                                // all resets of this step are done: fence them, off the critical path
            if (!(kUnif && P.split_main < 0)) {              // (the sorted layout fences after the last barrier of its sort)
                __syncthreads();
                xchg_early_fence(*Xp, tid);
            }
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
                double *x_h = s_dyn + soff, *x_B = x_h + NT;
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
                x_B[pos] = ust.s; x_aux[pos] = ust.aux_k;
                if (has) {
#pragma unroll
                    for (int c = 0; c < Model::C; c++) x_x[c * NT + pos] = (int32_t)x[c];
                }
                __syncthreads();
                PHASE(15);
                if constexpr (PUSH) xchg_early_fence(*Xp, tid);   // after the LAST CTA-wide barrier before the loop (nobody waits for it)
                // Sorted chunk -> warp.  The main warps take the chunks in snake order over the schedulers.  When the
                // chunks are 4 W + 1 or 4 W + 2 (P.split_main == -2) the last one or two are shared in TIME by two helper
                // warps each, on different schedulers: warp 4W + g serves the first half of the batch's candidates
                // and hands the continuation over, warp 4W + 2 + g finishes the interval -- W + 1/2 rounds per scheduler
                // instead of W + 1 on two of them.
                // Heavier chunks go to HIGHER warp ids: the scheduler's arbiter favours the higher warp id among eligible
                // warps, so the warps with the longest dependent chains (the CTA's critical path) get more than an equal
                // share and finish with the lighter ones instead of alone, latency-bound, after them (measured: the 32
                // heaviest particles of a CTA end 3 us after the median warp instead of 8; +1.2 % on the filter;
                // profiles/r02b_probe_nat_vs_rev.txt).  Without helper warps the last warp keeps the lightest chunk: it
                // issues the exchange's fence.
                const int nw = NT >> 5, main_w = P.split_main == -2 ? nw - 4 : nw;
                const int nrev = (P.split_main == -2 || !PUSH) ? main_w : main_w - 1;
                int chunk;
                if (warp < nrev) {
                    const int wr = nrev - 1 - warp, rnd = wr >> 2, r_last = (nrev - 1) >> 2;
                    chunk = 4 * rnd + ((((r_last - rnd) & 1) == 0) ? (wr & 3) : 3 - (wr & 3));   // (the last round is in natural order)
                } else if (warp < main_w) {
                    chunk = main_w - 1;
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
                            ust.t_rem = s_ct[hg][lane]; ust.s = s_cB[hg][lane]; ust.h = s_ch[hg][lane];
                            ust.cand = s_ck[hg][lane]; ust.first = s_cu[hg][lane][0]; ust.last = s_cu[hg][lane][1]; ust.aux_k = s_cu[hg][lane][2];
                        }
                    }
                } else {
                    run = run_rec;
                    if (home >= 0) {
#pragma unroll
                        for (int c = 0; c < Model::C; c++) x[c] = (double)x_x[c * NT + slot];
                    }
                    if (run) { ust.h = x_h[slot]; ust.s = x_B[slot]; ust.last = x_K[slot]; ust.aux_k = x_aux[slot]; }
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
                s_ct[hg][lane] = ust.t_rem; s_cB[hg][lane] = ust.s; s_ch[hg][lane] = ust.h;
                s_ck[hg][lane] = ust.cand; s_cu[hg][lane][0] = ust.first; s_cu[hg][lane][1] = ust.last; s_cu[hg][lane][2] = ust.aux_k;
                s_cfin[hg][lane] = fin ? 1 : 0;
                __threadfence_block();
                asm volatile("bar.sync %0, 64;" ::"r"(1 + hg) : "memory");
            }
            WARP_END(run ? ust.last : 0u);
            if (sorted && home >= 0 && leg != 1) {           // back to the home thread (read after the barrier below)
                const int NT = blockDim.x;
                int32_t *x_ret = (int32_t *)(s_dyn + soff + 2 * NT) + Model::C * NT;
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
        CTA_T(0);
        if constexpr (kUnif) {
            if (active && via_smem) {
                const int32_t *x_ret = (const int32_t *)(s_dyn + soff + 2 * blockDim.x) + Model::C * blockDim.x;
#pragma unroll
                for (int c = 0; c < Model::C; c++) {
                    const int32_t v = x_ret[c * blockDim.x + pidx];
                    x[c] = (double)v;
                    Xr[(size_t)c * N + j] = v;
                }
            }
        }
        if (p < P.T - 1) {
            ls = weigh_local<Model, PUSH>(P, p, f, b, tid, active, j, x, sm, s_tab);
            PHASE(19);
            CTA_T(1);
            if constexpr (PUSH) xchg_publish(*Xp, fsh, Xp->gen0 + (unsigned)p, gb, tid, ls.mb, ls.sb);
        }
        PHASE(5);
    }
    if (P.n_events) {                                        // one global atomic per CTA for the whole filter
#pragma unroll
        for (int d = 16; d; d >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, d);
        if ((tid & 31) == 0 && my_pairs) atomicAdd(&s_pairs, my_pairs);
        __syncthreads();
        if (tid == 0 && s_pairs) atomicAdd(&P.n_events[f], s_pairs);
    }
    if constexpr (PUSH) {
        grid.sync();                                         // this rank's history, events and status are complete
        if (b == 0 && tid == 0 && xchg_failed(*Xp, fsh)) atomicExch(&P.status[f], SEM_ERR_PEER);
        if (b == 0 && tid < 32 && P.iter_out) {
            __syncwarp();
            if (Xp->W == 1) iteration_epilogue<Model::C>(P, f);          // one rank: the lineage never leaves this GPU
            else xchg_iteration_epilogue<Model::C>(P, *Xp, f, fsh);
        }
    } else if (P.iter_out) {                                 // path sample + packed result of the MH iteration
        grid.sync();
        if (b == 0 && tid < 32) iteration_epilogue<Model::C>(P, f);
    }
}

template <class Model, int ARITH>
__global__ void __launch_bounds__(kMaxThreads) pf_persistent(const __grid_constant__ PfDev P) {
    pf_persistent_body<Model, ARITH, false>(P, nullptr);
}

}  // namespace sem
