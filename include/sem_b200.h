/*
 * sem_b200.h -- C ABI of the B200 (sm_100a) particle-filter / SSA / ABC engine.
 *
 * This is the drop-in boundary for the hot path of GeorgeEfstathiadis/Stochastic-Epidemic-Modelling.
 * The reference is pure Python and has no FFI of its own; each entry point below replaces the body of
 * one reference function (file:line given), and INTEGRATION.md shows the ctypes stub a maintainer
 * would add to the reference to bind it.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (stream is an opaque void* = cudaStream_t).
 *   - every pointer documented "device" is a CUDA device pointer owned by the caller; the library
 *     allocates nothing persistent.  `workspace` is caller-allocated scratch of sem_pf_workspace_bytes().
 *   - every call enqueues on `stream` and returns without synchronising, except the *_host variants,
 *     which take host buffers, do their own staging and synchronise before returning.
 *   - return value: 0 = ok, negative = error (sem_last_error() gives the text).  A collapsed filter is
 *     NOT an error: it is reported through `status` (reference: pmcmc.py:191-192 returns (None,None,None)).
 *   - thread-safe per stream; no global mutable state besides the thread-local error string.
 *
 * State layout in HBM (structure of arrays, int32):
 *   X_hist   [n_filters][T][C][N]   compartment counts of every particle at every observation time
 *   ancestry [n_filters][T][N]      parent index at the previous time (row 0 = 0), as pmcmc.py:193
 *   log_zetas[n_filters][T]         log of the reference's running likelihood zetas[p] (pmcmc.py:183)
 */
#ifndef SEM_B200_H
#define SEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEM_ABI_VERSION 2
#define SEM_MAX_GROUPS 4

/* pmcmc.py:116-120 ModelType */
enum { SEM_MODEL_SIR = 0, SEM_MODEL_SEIR = 1, SEM_MODEL_SIR_SUBGROUPS = 2, SEM_MODEL_SIR_SUBGROUPS2 = 3 };
/* pmcmc.py:178-181: observations=False -> binomial pmf, observations=True -> normal pdf */
enum { SEM_OBS_BINOMIAL = 0, SEM_OBS_NORMAL = 1 };
/* pmcmc.py:188-190 is multinomial; systematic is the production resampler (one uniform per step) */
enum { SEM_RESAMPLE_MULTINOMIAL = 0, SEM_RESAMPLE_SYSTEMATIC = 1 };
/* How one observation interval is simulated.  REFERENCE / FAST are the Gillespie direct method in the reference's
 * own fp64 operation order (gillespie_algo.py:38-39,62-63) or an algebraically equal form with a single division.
 * UNIFORMIZED is exact too (same law of the state at the end of the interval) but draws no waiting times: the jump
 * chain is thinned from a rate-B Poisson stream of candidates whose NUMBER in the interval is drawn once; if a
 * fired event lifts the total propensity above B, that candidate's time is drawn from its order-statistic (Beta)
 * law and the rest of the interval restarts with a new bound.
 * FAST32 is FAST with 32-bit uniforms: event k of a particle-step takes two words of Philox call k/2 (u1 = words 0/2,
 * u2 = words 1/3), so one Philox4x32-10 call serves two events; all arithmetic stays fp64.
 * UNIFORMIZED32 is the production form of UNIFORMIZED: 32-bit candidate uniforms (candidate c takes word (c & 3) of Philox
 * call c >> 2: four candidates per call), a bound that anticipates growth (B = max(a0(x), a0(x + drift h)) (1 + 2/sqrt(a0 h + 1)),
 * batches shortened so that the anticipated growth stays below 1.25), and no direct-method tail.  Same law of the state
 * at the end of the interval as the direct method.  Because a batch's candidate count is drawn BEFORE the loop runs, the
 * whole-filter kernel sorts each CTA's particles by it so that the lanes of a warp carry equal work (a scheduling change
 * only: streams are keyed by the particle).  The Python layer's default "auto" picks UNIFORMIZED32 for SIR / SEIR
 * filters and FAST32 otherwise. */
enum { SEM_ARITH_REFERENCE = 0, SEM_ARITH_FAST = 1, SEM_ARITH_UNIFORMIZED = 2, SEM_ARITH_FAST32 = 3, SEM_ARITH_UNIFORMIZED32 = 4 };

enum {
    SEM_OK = 0,
    SEM_ERR_INVALID = -1,   /* bad argument */
    SEM_ERR_CUDA = -2,      /* CUDA runtime error */
    SEM_ERR_REPLAY = -3,    /* replay buffer exhausted */
    SEM_ERR_NO_DEVICE = -4, /* no sm_100 device */
    SEM_ERR_PEER = -5       /* sharded filter: a peer rank did not answer within the exchange's time limit */
};

int sem_abi_version(void);
const char *sem_last_error(void);
/* number of SMs / compute capability of the current device, <0 on error */
int sem_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Particle filter: replaces the body of particle_filter (pmcmc.py:123-233): X_0 initialisation,
 * and for p = 1..T-1: weight X[p-1] against Y[p-1] (min over columns), accumulate the likelihood,
 * resample, propagate every particle one observation interval with exact Gillespie SSA
 * (gillespie_algo.py:10-233).  The whole time loop runs on the device: ONE cooperative launch for the whole filter when
 * all CTAs are co-resident (grid-wide barrier per step), else one launch per step.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sem_pf_config {
    int32_t model;          /* SEM_MODEL_* */
    int32_t obs_kind;       /* SEM_OBS_* */
    int32_t resampler;      /* SEM_RESAMPLE_* */
    int32_t arith;          /* SEM_ARITH_* */
    int32_t n_particles;    /* N */
    int32_t n_obs;          /* T = rows of Y */
    int32_t n_groups;       /* G (1 for SIR/SEIR) */
    int32_t n_obs_cols;     /* columns of Y: 3, 4, 3G or (SUBGROUPS2) 3 */
    int32_t n_filters;      /* independent filters (chains / thetas) run side by side, >= 1 */
    int32_t block_particles;/* particles per CTA, 0 = choose from N and the SM count */
    int32_t store_history;  /* 1: write all T rows of X_hist/ancestry; 0: keep only two rows (ping-pong) */
    int32_t reserved;       /* launch mode.  0 (default): one cooperative launch for the whole filter when all CTAs are
                               co-resident -- with systematic resampling and one filter the kernel resamples in offspring
                               form through record buffers in the workspace (pf_persistent_x with one rank), otherwise
                               with a grid barrier + ancestor search per step (pf_persistent); 1: force one launch per
                               step; 2: force the grid-barrier kernel.  Results are bit-identical in all modes. */
    double probs;           /* p_obs (binomial) or noise ratio (normal), pmcmc.py:128 */
    double dt;              /* observation interval, the reference uses 1 (pmcmc.py:205) */
    uint64_t seed;          /* Philox key */
    uint32_t filter_id0;    /* stream id of filter 0; filter f uses filter_id0 + f */
    uint32_t path_exact;    /* iteration_result only: 0 = the reference's off-by-one ancestry indexing (pmcmc.py:244-246),
                               1 = true genealogy */
    double mu[SEM_MAX_GROUPS];           /* Poisson mean of I_0 per group (pmcmc.py:157,161,167) */
    double n_population[SEM_MAX_GROUPS]; /* population per group (pmcmc.py:158,162,168) */
} sem_pf_config;

typedef struct sem_pf_buffers {
    const double *Y;            /* device [T][n_obs_cols] */
    const double *theta;        /* device [n_filters][P]; P = 2 (SIR: beta,gamma), 3 (SEIR: beta,alpha,gamma),
                                   G*G+1 (subgroups: betas row-major [infector][susceptible], gamma).
                                   PRECONDITION: every rate constant >= 0 -- the reference rejects negative proposals before
                                   it calls the filter (pmcmc.py:333-337) and its simulators raise on a negative propensity
                                   (gillespie_algo.py:63); the event loops here do not end under one.  The Python layer checks
                                   host inputs (engine.check_rates); a C caller holding theta on the device must do the same. */
    const int32_t *X0;          /* device [C][N] initial state shared by all filters, or NULL = draw
                                   I_0 ~ Poisson(mu) on the device */
    /* replay mode (all three NULL = Philox).  Uniforms in the reference's consumption order:        */
    const double *replay_resample_u; /* device [T-1][N]: the N doubles np.random.choice draws at step p */
    const double *replay_ssa_u;      /* device: per (step,particle) the (u1,u2) pairs of its SSA events */
    const int64_t *replay_ssa_off;   /* device [(T-1)*N+1]: CSR offsets (in doubles) into replay_ssa_u */
    int32_t *X_hist;            /* device, see layout above (T rows, or 2 if !store_history) */
    int32_t *ancestry;          /* device */
    double *log_zetas;          /* device [n_filters][T] */
    int32_t *status;            /* device [n_filters]: 0 ok, p>0 collapsed at step p, -3 replay exhausted */
    uint64_t *n_events;         /* device [n_filters] total SSA events drawn (incl. discarded overshoots) or NULL */
    void *workspace;            /* device, sem_pf_workspace_bytes() */
    /* One PMCMC iteration in one call (pmcmc.py:354-371: particle_filter + particle_path_sampler): when not NULL
     * (needs store_history = 1) the run ends by sampling a path (final particle ~ Philox(seed, domain 4, filter id),
     * ancestry chased backwards) and packs, per filter, device double[SEM_ITER_HEADER + T*C] =
     * { log_zetas[T-1], status, n_events, chosen final particle, trajectory[T][C] } for a single D2H copy. */
    double *iteration_result;
    /* Observation parameter per filter (device [n_filters]; binomial p / normal sd factor), or NULL = cfg->probs for
     * every filter.  With theta per filter this lets one launch evaluate a BATCH of Metropolis-Hastings proposals that
     * also differ in the estimated p_obs (pmcmc.py:283-296 / 339-352). */
    const double *probs_per_filter;
} sem_pf_buffers;
#define SEM_ITER_HEADER 4

size_t sem_pf_workspace_bytes(const sem_pf_config *cfg);
/* number of int32 in X_hist / ancestry and doubles in log_zetas the caller must provide */
size_t sem_pf_hist_elems(const sem_pf_config *cfg);
size_t sem_pf_ancestry_elems(const sem_pf_config *cfg);
/* launches made by one sem_pf_run (for launch accounting) */
int sem_pf_launch_count(const sem_pf_config *cfg);
int sem_pf_run(const sem_pf_config *cfg, const sem_pf_buffers *buf, void *stream);
/* One Metropolis-Hastings iteration's device work in one call (pmcmc.py:354-371, for a batch of n_filters proposals):
 * copies theta_host [n_filters][P] (and probs_host [n_filters], or NULL) from pinned host memory into buf->theta
 * (buf->probs_per_filter), runs sem_pf_run with buf->iteration_result, and copies the packed results to result_host
 * [n_filters][SEM_ITER_HEADER + T*C].  Everything is enqueued on `stream`; the caller synchronises. */
int sem_pf_iteration(const sem_pf_config *cfg, const sem_pf_buffers *buf, const double *theta_host, const double *probs_host,
                     double *result_host, void *stream);

/* Host-buffer variant (end-to-end call): Y, theta, X0 (may be NULL) and all outputs are HOST pointers;
 * outputs may be NULL when not wanted.  X_hist_out is (T,N,C) float64 and ancestry_out (T,N) float64 exactly
 * as pmcmc.py:151-152,233 returns them; zetas_out = exp(log_zetas).  Returns 0, >0 = collapsed at that step. */
int sem_pf_run_host(const sem_pf_config *cfg, const double *Y, const double *theta, const int32_t *X0,
                    double *log_zetas_out, double *zetas_out, double *hidden_process_out, double *ancestry_out,
                    uint64_t *n_events_out);
/* sem_pf_run_host keeps one grow-only device workspace per device between calls; this returns the memory. */
int sem_host_workspace_release(void);

/* ------------------------------------------------------------------------------------------------
 * Particle-sharded filter (one filter too large for one GPU): every rank owns a contiguous slice
 * of the N_global particles and runs these three calls per step; the host (torch.distributed / NCCL) all-gathers the
 * per-shard weight summaries (max logw M_r, sum exp(logw - M_r)) and all-to-all-v's the children records.
 * cfg->n_particles is the LOCAL particle count, n_filters = 1.  Resampling is global systematic: slot j draws
 * v_j = ((j+u0)/N_global)*total and takes the first particle whose global cdf G_r + s_r*cdf_local exceeds it.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sem_shard_step {
    int32_t step;               /* p >= 1 */
    int32_t particle_offset;    /* global index of this shard's particle 0 (keys the Philox streams) */
    int64_t n_global;           /* N_global */
    double u0;                  /* the step's single systematic uniform */
    double total;               /* global weight total, in units of exp(-M_global) */
    double total_local;         /* this shard's total in its own units (summary[1] of the previous step) */
    double G, G_next;           /* global cdf at this shard's first particle / at the next shard's first particle */
    double s;                   /* exp(M_r - M_global) */
    int64_t slot0;              /* first global slot whose ancestor lives on this shard */
} sem_shard_step;

/* X_0 of the local slice + weights against Y[0]; summary: device double[2] <- (M_r, total_r) */
int sem_shard_init(const sem_pf_config *cfg, const sem_pf_buffers *buf, int32_t particle_offset, double *summary,
                   void *stream);
/* writes one record int32[C+1] = (state, global ancestor index) per child of this shard's particles, ordered by slot,
 * into send_records (device, capacity = number of slots in [slot0, slot0 + children)) */
int sem_shard_offspring(const sem_pf_config *cfg, const sem_pf_buffers *buf, const sem_shard_step *st,
                        int32_t *send_records, void *stream);
/* recv_records: device int32 [n_particles][C+1] for the local slots; propagates them to time p, stores X[p] /
 * ancestry[p] (global indices), weighs against Y[p] and refreshes summary */
int sem_shard_propagate(const sem_pf_config *cfg, const sem_pf_buffers *buf, const sem_shard_step *st,
                        const int32_t *recv_records, double *summary, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Particle-sharded filter with the exchange on the device (one cooperative launch per rank for the WHOLE filter,
 * no host work and no collective call per step).  Every rank runs sem_pf_run_sharded with the same cfg
 * (cfg->n_particles = particles per rank, equal on all ranks; n_filters = 1; resampler = systematic); rank r owns the
 * global particles [r n, (r+1) n).  The resampling barrier and the particle migration go through peer memory (NVLink):
 * each rank owns an "arena" that every other rank has mapped (CUDA IPC between processes, peer access inside one
 * process).  Per step every CTA stores its 16-byte weight partial into every rank's arena and every parent particle
 * stores one record (state, global parent index) per child into the arena of the rank that owns the child's slot;
 * readers poll for the data itself.  Replaces the globalised form of pmcmc.py:183-199.
 * Results equal the single-GPU filter of W n particles with the same seed and CTA size.
 * ---------------------------------------------------------------------------------------------- */
#define SEM_MAX_RANKS 8
typedef struct sem_xchg_desc {
    int32_t world, rank;
    uint32_t generation;    /* running counter of exchange generations: 0 after sem_xchg_reset; advanced by every run */
    uint32_t launch_tag;    /* advanced by every run (path-sampler tokens) */
    double timeout_s;       /* a rank that waits longer than this for a peer gives up: status SEM_ERR_PEER (<= 0: 20 s) */
    void *arena[SEM_MAX_RANKS];  /* every rank's arena as mapped in THIS process; arena[rank] is this rank's own */
} sem_xchg_desc;

/* bytes of one rank's arena for this configuration (identical on every rank) */
size_t sem_xchg_bytes(const sem_pf_config *cfg, int32_t world);
/* cudaMalloc an arena on the current device and export its CUDA IPC handle (64 bytes, may be NULL) */
int sem_xchg_alloc(size_t bytes, void **arena, unsigned char *ipc_handle);
/* map a peer process's arena from its IPC handle / unmap it / free an own arena */
int sem_xchg_open(const unsigned char *ipc_handle, void **peer_arena);
int sem_xchg_close(void *peer_arena);
int sem_xchg_free(void *arena);
/* same-process multi-GPU: let `device` access `peer_device`'s memory (idempotent) */
int sem_peer_enable(int32_t device, int32_t peer_device);
/* fill an own arena with the "empty" marks.  Needed once before the first run and after a run that ended with a
 * non-zero status (collapse or SEM_ERR_PEER); all ranks must have reset (host barrier) before any rank launches again,
 * and the descriptor's generation restarts at 0. */
int sem_xchg_reset(const sem_pf_config *cfg, int32_t world, void *arena, void *stream);
/* this rank's packed iteration result inside its arena (device double[SEM_ITER_HEADER + T*C]); pass it as
 * buf->iteration_result to have the run end with the path sample over all shards: every rank receives the same
 * trajectory; header = { log_zetas[T-1], status, this rank's n_events, chosen GLOBAL final particle } */
double *sem_xchg_iteration_result(const sem_pf_config *cfg, int32_t world, void *arena);
/* 1 when this configuration can run with the device-side exchange (all CTAs of a rank co-resident, n_filters = 1,
 * systematic resampling, world <= SEM_MAX_RANKS, Philox mode), else 0 with the reason in sem_last_error() */
int sem_pf_sharded_supported(const sem_pf_config *cfg, int32_t world);
/* one filter pass of this rank's shard.  ancestry holds GLOBAL parent indices; log_zetas / status are identical on all
 * ranks.  X0 (if given) is this rank's slice [C][n].  Advances x->generation and x->launch_tag. */
int sem_pf_run_sharded(const sem_pf_config *cfg, const sem_pf_buffers *buf, sem_xchg_desc *x, void *stream);
/* sem_pf_iteration for the sharded filter: theta_host [P] (pinned) -> buf->theta, this rank's launch with
 * buf->iteration_result = sem_xchg_iteration_result(arena), packed result -> result_host (pinned).  Enqueue only. */
int sem_pf_iteration_sharded(const sem_pf_config *cfg, const sem_pf_buffers *buf, sem_xchg_desc *x, const double *theta_host,
                             double *result_host, void *stream);

/* particle_path_sampler (pmcmc.py:236-248).  chosen < 0: pick uniformly with Philox(seed); exact = 0 keeps
 * the reference's off-by-one ancestry indexing, 1 follows the true genealogy.  traj: device [T][C] int32. */
int sem_path_sample(const int32_t *X_hist, const int32_t *ancestry, int32_t T, int32_t N, int32_t C,
                    int32_t chosen, int32_t exact, uint64_t seed, uint32_t filter_id, int32_t *traj, void *stream);

/* (T,C,N) int32 SoA history -> (T,N,C) float64 as the reference returns it (pmcmc.py:151); device pointers */
int sem_hist_to_f64(const int32_t *X_hist, int32_t T, int32_t N, int32_t C, double *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Independent SSA runs: replaces sir_simulate / seir_simulate / sir_subgroups_simulate
 * (gillespie_algo.py:10-75, 78-146, 148-233) for a batch of n_sims simulations.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sem_sim_config {
    int32_t model, n_groups, arith, n_sims;
    int32_t shared_theta;   /* 1: theta is [P] for all sims, 0: [n_sims][P] */
    int32_t shared_x0;      /* 1: x0 is [C], 0: [n_sims][C] */
    int64_t record_capacity;/* rows per sim in times/states (0 = last values only) */
    double max_time;
    uint64_t seed;
    uint32_t sim_index0;    /* Philox item id of sim 0 */
    int32_t daily_rows;     /* H > 0: states receives [n_sims][H][C] = state at the integer times 1..H (max_time >= H) */
} sem_sim_config;

/* x0 device int32; theta device; replay_u/replay_off device or NULL (Philox); x_out device [n_sims][C] int32;
 * n_rows device [n_sims] int64 (events+1); times device [n_sims][cap] double, states [n_sims][cap][C] int32. */
int sem_ssa_simulate(const sem_sim_config *cfg, const int32_t *x0, const double *theta, const double *replay_u,
                     const int64_t *replay_off, int32_t *x_out, int64_t *n_rows, double *times, int32_t *states,
                     void *stream);

/* ------------------------------------------------------------------------------------------------
 * Deterministic ODE data synthesiser, batched (replaces differential_sir / differential_seir /
 * differential_sir_subroups + odeint of pmcmc.py:16-52 and the daily sub-sampling of *_simulate_discrete,
 * pmcmc.py:54-113).  One thread per parameter set integrates the mean-field ODE over the caller's time
 * grid with classical RK4 (`substeps` equal steps between consecutive grid points, fp64) and stores the grid points
 * selected by row_of_grid (row index in the output, or -1): the reference keeps, per integer day d, the LAST grid
 * point with ceil(t) == d.  theta: SIR (beta, gamma); SEIR (beta, alpha, gamma); subgroups: beta row-major [i][j]
 * (the rate applied to susceptibles of group i by infectives of group j, as pmcmc.py:47 uses it), then gamma. */
typedef struct sem_ode_config {
    int32_t model, n_groups;    /* SEM_MODEL_*; subgroup models: G */
    int32_t n_sets;             /* parameter sets (threads) */
    int32_t n_grid;             /* points of the time grid */
    int32_t n_rows;             /* rows kept per set */
    int32_t substeps;           /* RK4 steps per grid interval (>= 1) */
    int32_t shared_y0;          /* 1: y0 is [C] for all sets, 0: [n_sets][C] */
    int32_t shared_theta;       /* 1: theta is [P], 0: [n_sets][P] */
} sem_ode_config;
/* all pointers device: y0 double, theta double, t_grid [n_grid] double, row_of_grid [n_grid] int32, out [n_sets][n_rows][C] double */
int sem_ode_daily(const sem_ode_config *cfg, const double *y0, const double *theta, const double *t_grid,
                  const int32_t *row_of_grid, double *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * ABC rejection trials: replaces the trial loop body of abc_algo (abc_algo.py:33-99) for SIR:
 * prior draw, Poisson-perturbed start, SSA, daily discretisation, L1 distance (abc_algo.py:10-13).
 * ---------------------------------------------------------------------------------------------- */
typedef struct sem_abc_config {
    int32_t n_days;         /* T = rows of observed_data */
    int32_t arith;
    int32_t early_reject;   /* stop a trial once its partial distance already exceeds threshold */
    int32_t reserved;
    int64_t n_trials;
    uint64_t trial0;        /* id of the first trial (shards: rank r starts at r*n_trials) */
    double threshold;
    double prior[4];        /* beta lo,hi ; gamma lo,hi (abc_algo.py:36-37) */
    uint64_t seed;
} sem_abc_config;

/* obs device [T][3] (S,I,R).  trial_ids device [n_trials] uint64 or NULL (= trial0 + i): lets accepted trials be
 * re-simulated to emit their trajectories.  Replay (tests): theta_in [n][2], n_start_in [n][3] int64,
 * replay_u / replay_off CSR; all NULL = Philox.  Outputs: theta_out device [n][2], distance device [n] (inf when
 * rejected early), traj device [n][T][3] int32 or NULL, n_events device [1] or NULL, work_counter device [1] uint64
 * scratch (zeroed by the call). */
int sem_abc_run(const sem_abc_config *cfg, const double *obs, const uint64_t *trial_ids, const double *theta_in,
                const int64_t *n_start_in, const double *replay_u, const int64_t *replay_off, double *theta_out,
                double *distance, int32_t *traj, uint64_t *n_events, uint64_t *work_counter, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Small exported pieces used by the parity tests (device-side evaluation of host arrays of length n).
 * ---------------------------------------------------------------------------------------------- */
int sem_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int sem_test_binom_logpmf(const double *k, const double *n, const double *p, double *out, int64_t count);
int sem_test_norm_logpdf(const double *y, const double *x, const double *probs, double *out, int64_t count);
int sem_test_poisson(double mu, uint64_t seed, uint32_t domain, uint32_t c2, double *out, int64_t count);
/* the FAST arithmetic's -log(x), x in (0,1], and 1/a (host arrays in, host arrays out) */
int sem_test_fast_math(const double *x, const double *a, double *neglog_out, double *rcp_out, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* SEM_B200_H */
