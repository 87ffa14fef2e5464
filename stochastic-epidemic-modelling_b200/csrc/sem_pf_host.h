// sem_pf_host.h -- host-side sizing / argument marshalling of the particle filter, shared by sem_pf.cu and sem_pf_xchg.cu
#pragma once
#include <mutex>

#include "sem_pf_dev.cuh"

namespace sem {

// ---------------------------------------------------------------------------------------------- host side
// particles per CTA: one CTA per SM when the whole population is co-resident, else 256-wide CTAs
static int choose_ppb(const sem_pf_config *c) {
    if (c->block_particles > 0) {
        const int lim = c->arith == SEM_ARITH_UNIFORMIZED ? kMaxThreadsUnif : kMaxThreads;
        return c->block_particles > lim ? lim : c->block_particles;
    }
    if (c->n_filters > 1) {
        // several filters side by side: give each an equal share of the SMs so that all CTAs of all filters are
        // co-resident (one launch for the whole batch); the occupancy query of the caller has the last word
        const int cap = c->arith == SEM_ARITH_UNIFORMIZED ? kMaxThreadsUnif : kMaxThreads;
        const int share = sm_count() / c->n_filters > 0 ? sm_count() / c->n_filters : 1;
        long long ppb = (c->n_particles + share - 1) / share;
        if (ppb < 32) ppb = 32;
        if (ppb > cap) ppb = 256;                            // not co-resident anyway: 256-wide CTAs in waves
        if (ppb > c->n_particles) ppb = c->n_particles;
        return (int)ppb;
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

struct WsLayout { size_t L[2], pfx[2], scale[2], total[2], part, counter, wtab, xarena, xarena_bytes, bytes; int nb, ppb, wt_n; };

// total population = the largest count a compartment (or a group sum) can hold; 0 = no table (unknown, or > 1 GiB)
static int weight_table_n(const sem_pf_config *c) {
    const int G = c->model >= SEM_MODEL_SIR_SUBGROUPS ? c->n_groups : 1;
    double tot = 0;
    for (int g = 0; g < G; g++) tot += c->n_population[g];
    static int env_off = -1;
    if (env_off < 0) { const char *e = getenv("SEM_NO_WEIGHT_TABLE"); env_off = (e && e[0] == '1') ? 1 : 0; }
    if (env_off || c->obs_kind != SEM_OBS_BINOMIAL || !(tot >= 1) || tot != (double)(long long)tot || c->n_obs < 2) return 0;   // (the normal pdf is cheap)
    // The table costs (T-1) Cobs (pop+1) evaluations per launch against N F (T-1) Cobs direct ones: it pays only when the
    // particles comfortably outnumber the counts a compartment can take (headline: N = 1e5, pop = 1e4), and stays small.
    if ((double)c->n_particles * c->n_filters < 2.0 * (tot + 1)) return 0;
    const double bytes = (double)(c->n_obs - 1) * c->n_obs_cols * (tot + 1) * sizeof(double);
    return bytes <= 268435456.0 ? (int)tot : 0;
}
static WsLayout ws_layout(const sem_pf_config *c);

// One rank's arena: control block (error flag, path-sampler mailbox), partial tables [3][NB], record buffers [2][N][RW],
// packed iteration result.  Identical on every rank (same cfg, same device type).
struct ArenaLayout { size_t err, mail, part, rec, iter, bytes; int NB, nb, RW, C; };

static ArenaLayout arena_layout_nb(const sem_pf_config *cfg, int world, int nb_rank) {
    ArenaLayout a;
    struct { int nb; } w{nb_rank};
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
    a.C = model_cols(cfg->model, G);
    a.RW = (a.C + 1 + 3) & ~3;
    a.nb = w.nb; a.NB = w.nb * world;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    a.err = take(64); a.mail = take(64);
    a.part = take(3 * (size_t)a.NB * sizeof(double2));
    a.rec = take(2 * (size_t)cfg->n_particles * a.RW * sizeof(int32_t));
    a.iter = take((SEM_ITER_HEADER + (size_t)cfg->n_obs * a.C) * sizeof(double));
    a.bytes = off;
    return a;
}


// push-form resampling (the sharded filter's kernel with one rank, sem_pf_xchg.cu) is also the single-GPU whole-filter
// kernel of sem_pf_run when these hold; its arena then lives in the caller's workspace and is re-marked per launch
static bool push_eligible(const sem_pf_config *c) {
    static int env_off = -1;
    if (env_off < 0) { const char *e = getenv("SEM_NO_PUSH"); env_off = (e && e[0] == '1') ? 1 : 0; }
    return !env_off && c->n_filters >= 1 && c->resampler == SEM_RESAMPLE_SYSTEMATIC && c->reserved == 0 && c->n_obs >= 2 &&
           (c->arith == SEM_ARITH_FAST32 || c->arith == SEM_ARITH_UNIFORMIZED32);
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
    w.xarena_bytes = push_eligible(c) ? arena_layout_nb(c, 1, w.nb).bytes * (size_t)c->n_filters : 0;   // one arena per filter
    w.xarena = take(w.xarena_bytes);
    w.bytes = off;
    return w;
}
static ArenaLayout arena_layout(const sem_pf_config *cfg, int world) { return arena_layout_nb(cfg, world, ws_layout(cfg).nb); }

// sem_pf_xchg.cu: launch the one-rank exchange kernel on an arena inside the workspace (launched = false: not available
// for this configuration / device, the caller takes the grid-barrier kernel)
int xchg_run_single(const sem_pf_config *cfg, PfDev &P, const WsLayout &w, void *arena, cudaStream_t s, bool *launched);
bool xchg_single_available(const sem_pf_config *cfg);

static int hist_rows(const sem_pf_config *c) { return c->store_history ? c->n_obs : (c->n_obs < 2 ? c->n_obs : 2); }

static int fill_dev(const sem_pf_config *cfg, const sem_pf_buffers *buf, PfDev &P, WsLayout &w, bool &replay) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!buf || !buf->Y || !buf->theta || !buf->X_hist || !buf->ancestry || !buf->log_zetas || !buf->status || !buf->workspace) {
        set_error("null buffer"); return SEM_ERR_INVALID;
    }
    replay = buf->replay_ssa_u != nullptr;
    if (cfg->arith == SEM_ARITH_UNIFORMIZED32 && (rc = ktab_bind()) != SEM_OK) return rc;   // candidate-count tables of this device
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
    P.Y = buf->Y; P.theta = buf->theta; P.X0 = buf->X0; P.probs_f = buf->probs_per_filter;
    P.res_u = buf->replay_resample_u; P.ssa_u = buf->replay_ssa_u; P.ssa_off = (const long long *)buf->replay_ssa_off;
    P.X_hist = buf->X_hist; P.ancestry = buf->ancestry; P.status = buf->status; P.log_zetas = buf->log_zetas;
    P.n_events = (unsigned long long *)buf->n_events;
    for (int i = 0; i < 2; i++) {
        P.L[i] = (double *)(ws + w.L[i]); P.pfx[i] = (double *)(ws + w.pfx[i]);
        P.scale[i] = (double *)(ws + w.scale[i]); P.total[i] = (double *)(ws + w.total[i]);
    }
    P.part = (double2 *)(ws + w.part); P.counter = (unsigned int *)(ws + w.counter);
    // the table covers counts up to the configured population: only valid when X_0 is drawn from it (pmcmc.py:156-169)
    P.wt_n = (P.init_poisson && !replay && !P.probs_f) ? w.wt_n : 0;
    P.wtab = P.wt_n ? (double *)(ws + w.wtab) : nullptr;
    P.j0 = 0; P.sharded = 0; P.X_in = nullptr; P.summary = nullptr; P.split_main = 0;
    P.iter_out = buf->iteration_result; P.path_exact = (int)cfg->path_exact;
    if (P.iter_out && !cfg->store_history) { set_error("iteration_result needs store_history = 1"); return SEM_ERR_INVALID; }
    return SEM_OK;
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
static size_t persistent_smem(const sem_pf_config *cfg, int nb_filter, int threads, int split_main, bool push = false) {
    size_t b = (push ? 4 : 2) * (size_t)nb_filter * sizeof(double);   // nb_filter: CTAs of the whole filter (all ranks of a sharded one)
    if (split_main < 0) {
        const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1, C = model_cols(cfg->model, G);
        b += (size_t)threads * (2 * sizeof(double) + (2 * C + 3) * sizeof(int32_t)) + 132 * sizeof(uint32_t);
    }
    return b;
}

// opt in to more than 48 KB of shared memory per CTA where the exchange area needs it (once per kernel AND device: the
// attribute is per device; one process may drive several GPUs from several threads)
static int persistent_prepare(const void *fn, size_t smem) {
    struct Done { const void *fn; size_t sz; int dev; };
    static Done done[128];
    static int n_done = 0;
    static std::mutex mu;
    if (smem <= 24 * 1024) return SEM_OK;
    int dev = 0;
    SEM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_done; i++) if (done[i].fn == fn && done[i].dev == dev && done[i].sz >= smem) return SEM_OK;
    SEM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_done < 128) done[n_done++] = Done{fn, smem, dev};
    return SEM_OK;
}

}  // namespace sem
