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
#include "sem_pf_host.h"

namespace sem {

__global__ void __launch_bounds__(256) weight_table_kernel(const __grid_constant__ PfDev P) {
    __shared__ double2 s_tab[kLogTabSize];
    load_logtab(s_tab);
    __syncthreads();
    weight_table_fill(P, blockIdx.x * (size_t)blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x, s_tab);
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

template <int C>
__global__ void iteration_epilogue_kernel(const __grid_constant__ PfDev P) {
    iteration_epilogue<C>(P, blockIdx.x);                   // (one warp per filter)
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


size_t sem_pf_hist_elems(const sem_pf_config *c) {
    if (validate(c)) return 0;
    const int G = c->model >= SEM_MODEL_SIR_SUBGROUPS ? c->n_groups : 1;
    return (size_t)c->n_filters * hist_rows(c) * model_cols(c->model, G) * c->n_particles;
}
size_t sem_pf_ancestry_elems(const sem_pf_config *c) { return validate(c) ? 0 : (size_t)c->n_filters * hist_rows(c) * c->n_particles; }
int sem_pf_launch_count(const sem_pf_config *c);

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
    const size_t smem = persistent_smem(cfg, w.nb, threads, split_main);
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
    if (!replay && use_persistent(cfg, w, replay)) {         // systematic resampling, one filter: the exchange kernel with one rank
        bool launched = false;                               // (resampling in offspring form, no grid barrier, no ancestor search)
        rc = xchg_run_single(cfg, P, w, (char *)buf->workspace + w.xarena, s, &launched);
        if (rc) return rc;
        if (launched) return SEM_OK;
    }
    if (use_persistent(cfg, w, replay)) {                    // one launch, nothing else (the kernel initialises its outputs)
        void *args[] = {(void *)&P};
        const dim3 grid(w.nb, cfg->n_filters), block(persistent_threads(cfg, w, &P.split_main));
        SEM_CUDA(cudaLaunchCooperativeKernel(persistent_kernel(cfg), grid, block, args, persistent_smem(cfg, w.nb, block.x, P.split_main), s));
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

// One PMCMC iteration's device work enqueued by ONE call (pmcmc.py:354-371 for a batch of n_filters proposals): H2D of
// the proposals' theta (and observation parameters) from pinned host memory, sem_pf_run with iteration_result, D2H of
// the packed results.  Does not synchronise.
int sem_pf_iteration(const sem_pf_config *cfg, const sem_pf_buffers *buf, const double *theta_host, const double *probs_host,
                     double *result_host, void *stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!buf || !buf->theta || !buf->iteration_result || !theta_host || !result_host) { set_error("sem_pf_iteration: null buffer"); return SEM_ERR_INVALID; }
    if (probs_host && !buf->probs_per_filter) { set_error("sem_pf_iteration: probs_host needs probs_per_filter"); return SEM_ERR_INVALID; }
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1, C = model_cols(cfg->model, G);
    const size_t F = (size_t)cfg->n_filters;
    cudaStream_t s = (cudaStream_t)stream;
    SEM_CUDA(cudaMemcpyAsync((void *)buf->theta, theta_host, F * model_ntheta(cfg->model, G) * sizeof(double), cudaMemcpyHostToDevice, s));
    if (probs_host) SEM_CUDA(cudaMemcpyAsync((void *)buf->probs_per_filter, probs_host, F * sizeof(double), cudaMemcpyHostToDevice, s));
    rc = sem_pf_run(cfg, buf, stream);
    if (rc) return rc;
    SEM_CUDA(cudaMemcpyAsync(result_host, buf->iteration_result, F * (SEM_ITER_HEADER + (size_t)cfg->n_obs * C) * sizeof(double),
                             cudaMemcpyDeviceToHost, s));
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
    SEM_CUDA(cudaMemcpyFromSymbol(host_out, g_phase, sizeof(unsigned long long) * 24 * 256));
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

static void *g_host_ws[64] = {nullptr};                     // sem_pf_run_host's workspace, per device
static size_t g_host_ws_bytes[64] = {0};
static std::mutex g_host_ws_mu[64];

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
    // One grow-only device workspace per device, reused across calls (a PMCMC loop calls this once per iteration; the
    // cudaMalloc / cudaFree pair of a 160 MB block costs more than the headline filter itself).  Calls on the same device
    // are serialised by the lock; sem_host_workspace_release() returns the memory.
    int dev = 0;
    SEM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index out of range"); return SEM_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(g_host_ws_mu[dev]);
    if (g_host_ws_bytes[dev] < tot) {
        if (g_host_ws[dev]) { cudaFree(g_host_ws[dev]); g_host_ws[dev] = nullptr; g_host_ws_bytes[dev] = 0; }
        SEM_CUDA(cudaMalloc(&g_host_ws[dev], tot));
        g_host_ws_bytes[dev] = tot;
    }
    d = (char *)g_host_ws[dev];
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
    if (rc != SEM_OK) return rc;
    return status;
}

int sem_host_workspace_release(void) {
    for (int dev = 0; dev < 64; dev++) {
        std::lock_guard<std::mutex> lock(g_host_ws_mu[dev]);
        if (g_host_ws[dev]) {
            int cur = 0;
            SEM_CUDA(cudaGetDevice(&cur));
            SEM_CUDA(cudaSetDevice(dev));
            cudaFree(g_host_ws[dev]);
            SEM_CUDA(cudaSetDevice(cur));
            g_host_ws[dev] = nullptr; g_host_ws_bytes[dev] = 0;
        }
    }
    return SEM_OK;
}

}  // extern "C"
