// sem_sim_abc.cu -- independent SSA runs (gillespie_algo.py:10-233) and ABC rejection trials (abc_algo.py:17-109).
//
// sem_ssa_simulate : one thread per simulation; optional event log (last_values_only=False, gillespie_algo.py:68-75).
// sem_abc_run      : persistent lanes pull trial ids from a global work counter, so a warp never idles on the
//                    short epidemics of its neighbours (trial length varies from ~10 to ~10^4 events with theta).
//                    The daily discretisation + L1 distance of abc_algo.py:58-99 is accumulated on the fly at day
//                    boundaries; no event trajectory is materialised.
#include "sem_common.cuh"
#include "sem_host.h"

namespace sem {

// ------------------------------------------------------------------------------------------ simulate
struct SimDev {
    int n_sims, shared_theta, shared_x0, ntheta, daily;
    long long cap;
    double max_time;
    PhiloxKey key;
    uint32_t sim0;
    const int32_t *x0;
    const double *theta, *replay_u;
    const long long *replay_off;
    int32_t *x_out, *states;
    long long *n_rows;
    double *times;
};

template <int C>
struct EventLog {
    double *times; int32_t *states; long long cap; long long n;
    __device__ __forceinline__ void operator()(double t, const double *x) {
        if (n < cap) {
            times[n] = t;
#pragma unroll
            for (int c = 0; c < C; c++) states[n * C + c] = (int32_t)x[c];
        }
        n++;
    }
};
// state at the integer times 1..H (what the reference's forward-prediction script extracts per day,
// tests/pred_tmps.py:55-64): days strictly before an event's time keep the pre-event state
template <int C>
struct DailyLog {
    int32_t *rows; int H, day; double prev[C];
    __device__ __forceinline__ void flush_to(double t) {
        while (day <= H && (double)day < t) {
#pragma unroll
            for (int c = 0; c < C; c++) rows[(size_t)(day - 1) * C + c] = (int32_t)prev[c];
            day++;
        }
    }
    __device__ __forceinline__ void operator()(double t, const double *x) {
        flush_to(t);
#pragma unroll
        for (int c = 0; c < C; c++) prev[c] = x[c];
    }
};
template <int C>
struct DailyLogRef { DailyLog<C> *l; __device__ __forceinline__ void operator()(double t, const double *x) const { (*l)(t, x); } };

template <int C>
struct EventLogRef { EventLog<C> *l; __device__ __forceinline__ void operator()(double t, const double *x) const { (*l)(t, x); } };

template <class Model, int ARITH, bool REPLAY>
__global__ void __launch_bounds__(128) sim_kernel(const __grid_constant__ SimDev P) {
    __shared__ double2 s_tab[kLogTabSize];
    if (ARITH != SEM_ARITH_REFERENCE && !REPLAY) load_logtab(s_tab);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_sims) return;
    double x[Model::C];
    const int32_t *x0 = P.x0 + (P.shared_x0 ? 0 : (size_t)i * Model::C);
#pragma unroll
    for (int c = 0; c < Model::C; c++) x[c] = (double)x0[c];
    Model m;
    m.setup(P.theta + (P.shared_theta ? 0 : (size_t)i * P.ntheta), x);
    PairSource<REPLAY> src;
    if constexpr (REPLAY) src.init(P.replay_u, P.replay_off[i], P.replay_off[i + 1]);
    else src.init(P.key, P.sim0 + (uint32_t)i, 0u, stream_word(DOM_SIM, 0));
    long long rows = 1;
    // event times / daily states only exist in the direct method: the uniformized orders log with their direct counterpart
    constexpr int kLogArith = ARITH == SEM_ARITH_UNIFORMIZED ? SEM_ARITH_FAST : ARITH == SEM_ARITH_UNIFORMIZED32 ? SEM_ARITH_FAST32 : ARITH;
    if (P.daily > 0) {
        DailyLog<Model::C> log;
        log.rows = P.states + (size_t)i * P.daily * Model::C; log.H = P.daily; log.day = 1;
#pragma unroll
        for (int c = 0; c < Model::C; c++) log.prev[c] = x[c];
        const long long pr = ssa_run<Model, kLogArith, REPLAY, true>(m, x, P.max_time, src, s_tab, DailyLogRef<Model::C>{&log});
        log.flush_to(1e300);                                              // forward fill to the horizon
        rows = pr < 0 ? -1 : P.daily;
    } else if (P.cap > 0) {
        EventLog<Model::C> log{P.times + (size_t)i * P.cap, P.states + (size_t)i * P.cap * Model::C, P.cap, 0};
        log(0.0, x);                                                      // row 0 = initial state at time 0 (gillespie_algo.py:28-33)
        const long long pr = ssa_run<Model, kLogArith, REPLAY, true>(m, x, P.max_time, src, s_tab, EventLogRef<Model::C>{&log});
        rows = pr < 0 ? -1 : log.n;
    } else {
        const long long pr = ssa_run<Model, ARITH, REPLAY, false>(m, x, P.max_time, src, s_tab, NoRec());
        rows = pr < 0 ? -1 : 0;
    }
#pragma unroll
    for (int c = 0; c < Model::C; c++) P.x_out[(size_t)i * Model::C + c] = (int32_t)x[c];
    if (P.n_rows) P.n_rows[i] = rows;
}

template <class Model>
static void launch_sim(const SimDev &P, int arith, bool replay, cudaStream_t s) {
    const int threads = 128, blocks = (P.n_sims + threads - 1) / threads;
    if (replay) sim_kernel<Model, SEM_ARITH_REFERENCE, true><<<blocks, threads, 0, s>>>(P);
    else if (arith == SEM_ARITH_REFERENCE) sim_kernel<Model, SEM_ARITH_REFERENCE, false><<<blocks, threads, 0, s>>>(P);
    else if (arith == SEM_ARITH_UNIFORMIZED && P.cap == 0 && P.daily == 0)      // no event times exist to log
        sim_kernel<Model, SEM_ARITH_UNIFORMIZED, false><<<blocks, threads, 0, s>>>(P);
    else if (arith == SEM_ARITH_UNIFORMIZED32 && P.cap == 0 && P.daily == 0)
        sim_kernel<Model, SEM_ARITH_UNIFORMIZED32, false><<<blocks, threads, 0, s>>>(P);
    else if (arith == SEM_ARITH_FAST32 || arith == SEM_ARITH_UNIFORMIZED32) sim_kernel<Model, SEM_ARITH_FAST32, false><<<blocks, threads, 0, s>>>(P);
    else sim_kernel<Model, SEM_ARITH_FAST, false><<<blocks, threads, 0, s>>>(P);
}

// ------------------------------------------------------------------------------------------ ODE synthesiser
// Mean-field ODEs of pmcmc.py:16-52 (N = sum of the state, recomputed per evaluation like the reference), classical RK4,
// one thread per parameter set; the state lives in registers (C <= 12 doubles).
struct OdeDev {
    int n_sets, n_grid, n_rows, substeps, shared_y0, shared_theta, G;
    const double *y0, *theta, *t;
    const int32_t *row_of_grid;
    double *out;
};

template <int MODEL, int G> struct OdeDims { static constexpr int C = MODEL == SEM_MODEL_SIR ? 3 : MODEL == SEM_MODEL_SEIR ? 4 : 3 * G,
                                              P = MODEL == SEM_MODEL_SIR ? 2 : MODEL == SEM_MODEL_SEIR ? 3 : G * G + 1; };

template <int MODEL, int G>
__device__ __forceinline__ void ode_rhs(const double *th, const double *y, double *d) {
    constexpr int C = OdeDims<MODEL, G>::C;
    double N = 0.0;
#pragma unroll
    for (int c = 0; c < C; c++) N += y[c];
    if constexpr (MODEL == SEM_MODEL_SIR) {                      // pmcmc.py:16-24
        d[0] = -th[0] * y[0] * y[1] / N;
        d[1] = ((th[0] * y[0] / N) - th[1]) * y[1];
        d[2] = th[1] * y[1];
    } else if constexpr (MODEL == SEM_MODEL_SEIR) {              // :27-35
        d[0] = -th[0] * y[0] * y[2] / N;
        d[1] = th[0] * y[0] * y[2] / N - th[1] * y[1];
        d[2] = th[1] * y[1] - th[2] * y[2];
        d[3] = th[2] * y[2];
    } else {                                                     // :38-52: group i is infected by sum_j beta[i][j] I_j
        const double gamma = th[G * G];
#pragma unroll
        for (int i = 0; i < G; i++) {
            double force = 0.0;
#pragma unroll
            for (int j = 0; j < G; j++) force += th[i * G + j] * y[3 * j + 1];
            const double inf = y[3 * i] * force / N;
            d[3 * i] = -inf; d[3 * i + 1] = inf - gamma * y[3 * i + 1]; d[3 * i + 2] = gamma * y[3 * i + 1];
        }
    }
}

template <int MODEL, int G>
__global__ void __launch_bounds__(128) ode_kernel(const OdeDev P) {
    constexpr int C = OdeDims<MODEL, G>::C, NP = OdeDims<MODEL, G>::P;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.n_sets) return;
    double th[NP], y[C], k1[C], k2[C], k3[C], k4[C], w[C];
#pragma unroll
    for (int i = 0; i < NP; i++) th[i] = P.theta[(P.shared_theta ? 0 : (size_t)b * NP) + i];
#pragma unroll
    for (int c = 0; c < C; c++) y[c] = P.y0[(P.shared_y0 ? 0 : (size_t)b * C) + c];
    double *out = P.out + (size_t)b * P.n_rows * C;
    for (int k = 0; k < P.n_grid; k++) {
        if (k > 0) {
            const double h = (P.t[k] - P.t[k - 1]) / (double)P.substeps;
            for (int q = 0; q < P.substeps; q++) {
                ode_rhs<MODEL, G>(th, y, k1);
#pragma unroll
                for (int c = 0; c < C; c++) w[c] = y[c] + 0.5 * h * k1[c];
                ode_rhs<MODEL, G>(th, w, k2);
#pragma unroll
                for (int c = 0; c < C; c++) w[c] = y[c] + 0.5 * h * k2[c];
                ode_rhs<MODEL, G>(th, w, k3);
#pragma unroll
                for (int c = 0; c < C; c++) w[c] = y[c] + h * k3[c];
                ode_rhs<MODEL, G>(th, w, k4);
#pragma unroll
                for (int c = 0; c < C; c++) y[c] += h / 6.0 * (k1[c] + 2.0 * k2[c] + 2.0 * k3[c] + k4[c]);
            }
        }
        const int row = P.row_of_grid[k];
        if (row >= 0 && row < P.n_rows) {
#pragma unroll
            for (int c = 0; c < C; c++) out[(size_t)row * C + c] = y[c];
        }
    }
}

template <int MODEL, int G>
static void launch_ode(const OdeDev &P, cudaStream_t s) {
    ode_kernel<MODEL, G><<<(P.n_sets + 127) / 128, 128, 0, s>>>(P);
}

// ------------------------------------------------------------------------------------------ ABC
struct AbcDev {
    int T, early_reject;
    long long n_trials;
    unsigned long long trial0;
    double threshold, prior[4];
    PhiloxKey key;
    const double *obs;                  // [T][3]
    const unsigned long long *trial_ids;
    const double *theta_in, *replay_u;
    const long long *n_start_in, *replay_off;
    double *theta_out, *distance;
    int32_t *traj;
    unsigned long long *n_events, *work;
};

constexpr int kAbcSmemDays = 2048;        // observed series up to this length are staged in shared memory (32 KB); longer ones are read from L2

template <int ARITH, bool REPLAY>
__global__ void __launch_bounds__(128) abc_kernel(const __grid_constant__ AbcDev P) {
    extern __shared__ double s_obs_dyn[];                                // (I_obs, R_obs) per day, when the series fits
    __shared__ double2 s_tab[kLogTabSize];
    constexpr bool FAST = (ARITH == SEM_ARITH_FAST || ARITH == SEM_ARITH_FAST32) && !REPLAY;
    constexpr bool BITS32 = FAST && ARITH == SEM_ARITH_FAST32;           // two events per Philox call
    if (FAST) load_logtab(s_tab);
    const bool obs_smem = P.T <= kAbcSmemDays;
    if (obs_smem) for (int i = threadIdx.x; i < P.T; i += blockDim.x) { s_obs_dyn[2 * i] = P.obs[3 * i + 1]; s_obs_dyn[2 * i + 1] = P.obs[3 * i + 2]; }
    // (the reference has no limit on the length of the observed series, abc_algo.py:58-99)
    const double *obs_ir = obs_smem ? s_obs_dyn : P.obs + 1;             // element (day, c) at obs_ir[stride * day + c]
    const int obs_stride = obs_smem ? 2 : 3;
    __syncthreads();
    const int T = P.T;
    const double t_stop = (double)(T - 1);                               // rows 0..T-1 are the states at integer times (abc_algo.py:58-93)
    const double reject_at = P.threshold * 2.0 * T;
    unsigned long long my_events = 0;

    // per-lane trial state
    bool have = false, exhausted = false;
    long long slot = 0;
    double x[3] = {0, 0, 0}, t = 0, sI = 0, sR = 0;
    int day = 0;
    SirModel m;
    PairSource<REPLAY> src;

#define record_day()                                                                                   \
    do {                                                                                               \
        if (P.traj) {                                                                                  \
            int32_t *tr = P.traj + ((size_t)slot * T + day) * 3;                                       \
            tr[0] = (int32_t)x[0]; tr[1] = (int32_t)x[1]; tr[2] = (int32_t)x[2];                       \
        }                                                                                              \
        sI += fabs(x[1] - obs_ir[obs_stride * day]); sR += fabs(x[2] - obs_ir[obs_stride * day + 1]);  \
        day++;                                                                                         \
    } while (0)

    while (true) {
        if (!have && !exhausted) {
            // ------------------------------------------------------------ fetch + initialise a trial (abc_algo.py:34-40)
            slot = (long long)atomicAdd(P.work, 1ull);
            if (slot >= P.n_trials) exhausted = true;
            else {
                const unsigned long long id = P.trial_ids ? P.trial_ids[slot] : P.trial0 + (unsigned long long)slot;
                double beta, gamma;
                if constexpr (REPLAY) {
                    beta = P.theta_in[2 * slot]; gamma = P.theta_in[2 * slot + 1];
#pragma unroll
                    for (int c = 0; c < 3; c++) x[c] = (double)P.n_start_in[3 * slot + c];
                    src.init(P.replay_u, P.replay_off[slot], P.replay_off[slot + 1]);
                } else {
                    PairSource<false> ps; ps.init(P.key, (uint32_t)id, (uint32_t)(id >> 32), stream_word(DOM_ABC_PRIOR, 0));
                    double u1, u2; ps.next(u1, u2);
                    beta = P.prior[0] + (P.prior[1] - P.prior[0]) * u1;  // np.random.uniform(lo,hi) = lo + (hi-lo)*u (:36-37)
                    gamma = P.prior[2] + (P.prior[3] - P.prior[2]) * u2;
#pragma unroll
                    for (int c = 0; c < 3; c++) x[c] = poisson_draw(ps, (double)(long long)P.obs[c]);   // :39-40
                    src.init(P.key, (uint32_t)id, (uint32_t)(id >> 32), stream_word(DOM_ABC_SSA, 0));
                }
                P.theta_out[2 * slot] = beta; P.theta_out[2 * slot + 1] = gamma;
                const double th[2] = {beta, gamma};
                m.setup(th, x);
                t = 0; sI = 0; sR = 0; day = 0;
                record_day();                                             // row 0 = the perturbed start
                have = true;
            }
        }
        if (__all_sync(0xffffffffu, exhausted && !have)) break;
        if constexpr (FAST) {
            // ------------------------------------------------------------ one speculative block of events (ssa_block):
            // the events that fit before t_stop fire in order; a day boundary strictly before an event's time keeps the
            // pre-event state (abc_algo.py:58-84)
            if (have) {
                constexpr int U = BITS32 ? 4 : 2;
                bool done = !(m.alive(x) && day < T);
                bool rejected = false;
                if (!done) {
                    double xs[U + 1][3], tn[U], a0[U];
#pragma unroll
                    for (int c = 0; c < 3; c++) xs[0][c] = x[c];
                    ssa_block<SirModel, U, BITS32, true>(m, xs, t, src, s_tab, tn, a0);
                    int nf = 0;
#pragma unroll
                    for (int i = 0; i < U; i++) {
                        if (tn[i] <= t_stop) {                            // times are non-decreasing, NaN once a0 <= 0
                            nf++;
                            t = tn[i];
                            while (day < T && (double)day < t) {
#pragma unroll
                                for (int c = 0; c < 3; c++) x[c] = xs[i][c];
                                record_day();
                            }
                        }
                    }
                    double a_stop = a0[0];
#pragma unroll
                    for (int i = 1; i < U; i++) a_stop = (nf == i) ? a0[i] : a_stop;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        double v = xs[0][c];
#pragma unroll
                        for (int i = 1; i <= U; i++) v = (nf == i) ? xs[i][c] : v;
                        x[c] = v;
                    }
                    my_events += (unsigned long long)(nf + ((nf < U && a_stop > 0) ? 1 : 0));
                    if (nf < U) done = true;                              // overshoot of t_stop, extinction or a0 = 0
                    else if (P.early_reject && sI + sR > reject_at) { done = true; rejected = true; }
                }
                if (done) {
                    if (!rejected) while (day < T) record_day();          // forward fill (abc_algo.py:85-91)
                    P.distance[slot] = rejected ? CUDART_INF : (sI / T + sR / T) / 2;   // distance_function (abc_algo.py:10-13)
                    have = false;
                }
            }
        } else if (have) {
            // ------------------------------------------------------------ one SSA event (gillespie_algo.py:48-70)
            bool done = !(m.alive(x) && day < T);
            bool rejected = false, dry = false;
            if (!done) {
                double r[2], u1 = 0.5, u2 = 0.5, tau; int j;
                const double a0 = ssa_total<SirModel, SEM_ARITH_REFERENCE>(m, x, r);
                bool drew = a0 > 0;
                if (drew) { drew = src.next(u1, u2); dry = !drew; }
                if (!drew) done = true;
                else {
                    my_events++;
                    ssa_pick_ref<SirModel>(r, a0, u1, u2, tau, j);
                    const double tn = __dadd_rn(t, tau);
                    if (tn > t_stop) done = true;
                    else {
                        t = tn;
                        while (day < T && (double)day < t) record_day();  // days strictly before this event keep the old state
                        m.apply(x, j);
                        if (P.early_reject && sI + sR > reject_at) { done = true; rejected = true; }
                    }
                }
            }
            if (done) {
                if (!rejected) while (day < T) record_day();              // forward fill (abc_algo.py:85-91)
                // distance_function (abc_algo.py:10-13)
                P.distance[slot] = dry ? CUDART_NAN : (rejected ? CUDART_INF : (sI / T + sR / T) / 2);
                have = false;
            }
        }
    }
#undef record_day
    if (P.n_events && my_events) atomicAdd(P.n_events, my_events);
}

}  // namespace sem

using namespace sem;

extern "C" {

int sem_ssa_simulate(const sem_sim_config *cfg, const int32_t *x0, const double *theta, const double *replay_u,
                     const int64_t *replay_off, int32_t *x_out, int64_t *n_rows, double *times, int32_t *states, void *stream) {
    if (!cfg || !x0 || !theta || !x_out || cfg->n_sims < 1) { set_error("bad simulate args"); return SEM_ERR_INVALID; }
    if (cfg->model < 0 || cfg->model > 3) { set_error("bad model"); return SEM_ERR_INVALID; }
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
    if (G < 1 || G > SEM_MAX_GROUPS) { set_error("n_groups must be 1..4"); return SEM_ERR_INVALID; }
    if (cfg->record_capacity > 0 && (!times || !states)) { set_error("record buffers missing"); return SEM_ERR_INVALID; }
    if (cfg->daily_rows > 0 && (!states || cfg->record_capacity > 0)) { set_error("daily_rows needs states and no event log"); return SEM_ERR_INVALID; }
    if ((replay_u == nullptr) != (replay_off == nullptr)) { set_error("replay_u and replay_off go together"); return SEM_ERR_INVALID; }
    SimDev P;
    P.n_sims = cfg->n_sims; P.shared_theta = cfg->shared_theta; P.shared_x0 = cfg->shared_x0;
    P.ntheta = model_ntheta(cfg->model, G); P.cap = cfg->record_capacity; P.max_time = cfg->max_time; P.daily = cfg->daily_rows;
    P.key = make_philox_key(cfg->seed); P.sim0 = cfg->sim_index0;
    P.x0 = x0; P.theta = theta; P.replay_u = replay_u; P.replay_off = (const long long *)replay_off;
    P.x_out = x_out; P.states = states; P.n_rows = (long long *)n_rows; P.times = times;
    const bool replay = replay_u != nullptr;
    if (cfg->arith == SEM_ARITH_UNIFORMIZED32 && !replay) { const int rc = ktab_bind(); if (rc) return rc; }
    cudaStream_t s = (cudaStream_t)stream;
    switch (cfg->model) {
        case SEM_MODEL_SIR: launch_sim<SirModel>(P, cfg->arith, replay, s); break;
        case SEM_MODEL_SEIR: launch_sim<SeirModel>(P, cfg->arith, replay, s); break;
        default:
            switch (G) {
                case 1: launch_sim<SubModel<1>>(P, cfg->arith, replay, s); break;
                case 2: launch_sim<SubModel<2>>(P, cfg->arith, replay, s); break;
                case 3: launch_sim<SubModel<3>>(P, cfg->arith, replay, s); break;
                default: launch_sim<SubModel<4>>(P, cfg->arith, replay, s); break;
            }
    }
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

int sem_abc_run(const sem_abc_config *cfg, const double *obs, const uint64_t *trial_ids, const double *theta_in,
                const int64_t *n_start_in, const double *replay_u, const int64_t *replay_off, double *theta_out,
                double *distance, int32_t *traj, uint64_t *n_events, uint64_t *work_counter, void *stream) {
    if (!cfg || !obs || !theta_out || !distance || !work_counter) { set_error("bad abc args"); return SEM_ERR_INVALID; }
    if (cfg->n_days < 1) { set_error("n_days must be >= 1"); return SEM_ERR_INVALID; }
    if (cfg->n_trials < 1) { set_error("n_trials must be >= 1"); return SEM_ERR_INVALID; }
    const bool replay = replay_u != nullptr;
    if (replay && (!theta_in || !n_start_in || !replay_off)) { set_error("replay needs theta_in, n_start_in, replay_off"); return SEM_ERR_INVALID; }
    AbcDev P;
    P.T = cfg->n_days; P.early_reject = cfg->early_reject; P.n_trials = cfg->n_trials; P.trial0 = cfg->trial0;
    P.threshold = cfg->threshold;
    for (int i = 0; i < 4; i++) P.prior[i] = cfg->prior[i];
    P.key = make_philox_key(cfg->seed);
    P.obs = obs; P.trial_ids = (const unsigned long long *)trial_ids; P.theta_in = theta_in; P.replay_u = replay_u;
    P.n_start_in = (const long long *)n_start_in; P.replay_off = (const long long *)replay_off;
    P.theta_out = theta_out; P.distance = distance; P.traj = traj;
    P.n_events = (unsigned long long *)n_events; P.work = (unsigned long long *)work_counter;
    cudaStream_t s = (cudaStream_t)stream;
    SEM_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(uint64_t), s));
    if (n_events) SEM_CUDA(cudaMemsetAsync(n_events, 0, sizeof(uint64_t), s));
    const int threads = 128;
    const void *fn = replay ? (const void *)abc_kernel<SEM_ARITH_REFERENCE, true>
                     : cfg->arith == SEM_ARITH_REFERENCE ? (const void *)abc_kernel<SEM_ARITH_REFERENCE, false>
                     : (cfg->arith == SEM_ARITH_FAST32 || cfg->arith == SEM_ARITH_UNIFORMIZED32) ? (const void *)abc_kernel<SEM_ARITH_FAST32, false>
                                                      : (const void *)abc_kernel<SEM_ARITH_FAST, false>;   // (UNIFORMIZED: the ABC loop needs event times)
    int per_sm = 0;
    const size_t smem = cfg->n_days <= kAbcSmemDays ? (size_t)cfg->n_days * 2 * sizeof(double) : 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 4; }
    const long long want = (cfg->n_trials + threads - 1) / threads;
    const long long cap = (long long)sm_count() * per_sm;                // persistent lanes: every CTA resident, no second wave
    const int blocks = (int)(want < cap ? want : cap);
    void *args[] = {(void *)&P};
    SEM_CUDA(cudaLaunchKernel(fn, dim3(blocks), dim3(threads), args, smem, s));
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}


int sem_ode_daily(const sem_ode_config *cfg, const double *y0, const double *theta, const double *t_grid,
                  const int32_t *row_of_grid, double *out, void *stream) {
    if (!cfg || !y0 || !theta || !t_grid || !row_of_grid || !out) { set_error("bad ode args"); return SEM_ERR_INVALID; }
    if (cfg->model < 0 || cfg->model > 3 || cfg->n_sets < 1 || cfg->n_grid < 1 || cfg->n_rows < 1 || cfg->substeps < 1) { set_error("bad ode config"); return SEM_ERR_INVALID; }
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
    if (G < 1 || G > SEM_MAX_GROUPS) { set_error("n_groups must be 1..4"); return SEM_ERR_INVALID; }
    OdeDev P;
    P.n_sets = cfg->n_sets; P.n_grid = cfg->n_grid; P.n_rows = cfg->n_rows; P.substeps = cfg->substeps;
    P.shared_y0 = cfg->shared_y0; P.shared_theta = cfg->shared_theta; P.G = G;
    P.y0 = y0; P.theta = theta; P.t = t_grid; P.row_of_grid = row_of_grid; P.out = out;
    cudaStream_t s = (cudaStream_t)stream;
    switch (cfg->model) {
        case SEM_MODEL_SIR: launch_ode<SEM_MODEL_SIR, 1>(P, s); break;
        case SEM_MODEL_SEIR: launch_ode<SEM_MODEL_SEIR, 1>(P, s); break;
        default:
            switch (G) {
                case 1: launch_ode<SEM_MODEL_SIR_SUBGROUPS, 1>(P, s); break;
                case 2: launch_ode<SEM_MODEL_SIR_SUBGROUPS, 2>(P, s); break;
                case 3: launch_ode<SEM_MODEL_SIR_SUBGROUPS, 3>(P, s); break;
                default: launch_ode<SEM_MODEL_SIR_SUBGROUPS, 4>(P, s); break;
            }
    }
    SEM_CUDA(cudaGetLastError());
    return SEM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ test hooks
__global__ void k_philox(uint4 c, const __grid_constant__ PhiloxKey k, uint32_t *out) {
    const uint4 w = philox4x32_10(c.x, c.y, c.z, c.w, k);
    out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = w.w;
}
__global__ void k_binom(const double *k, const double *n, const double *p, double *out, long long cnt) {
    __shared__ double2 s_tab[kLogTabSize];
    load_logtab(s_tab);
    __syncthreads();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < cnt) out[i] = binom_logpmf_obs(binom_obs(k[i], s_tab), n[i], p[i], s_tab);      // the filter's entry point
}
__global__ void k_norm(const double *y, const double *x, const double *pr, double *out, long long cnt) {
    __shared__ double2 s_tab[kLogTabSize];
    load_logtab(s_tab);
    __syncthreads();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < cnt) out[i] = norm_logpdf(y[i], x[i], pr[i], s_tab);
}
__global__ void k_poisson(double mu, const __grid_constant__ PhiloxKey key, uint32_t domain, uint32_t c2, double *out, long long cnt) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < cnt) { PairSource<false> s; s.init(key, (uint32_t)i, c2, stream_word(domain, 0)); out[i] = poisson_draw(s, mu); }
}

__global__ void k_fast_math(const double *x, const double *a, double *nl, double *rc, long long cnt) {
    __shared__ double2 s_tab[kLogTabSize];
    load_logtab(s_tab);
    __syncthreads();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < cnt) { nl[i] = neg_log_fast(x[i], s_tab); rc[i] = rcp_nr(a[i]); }
}

template <class F>
static int host_map3(const double *a, const double *b, const double *c, double *out, int64_t n, F launch) {
    double *d = nullptr;
    SEM_CUDA(cudaMalloc(&d, (size_t)n * 4 * sizeof(double)));
    cudaMemcpy(d, a, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(d + n, b, n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d + 2 * n, c, n * 8, cudaMemcpyHostToDevice);
    launch(d, d + n, d + 2 * n, d + 3 * n);
    cudaError_t e = cudaMemcpy(out, d + 3 * n, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("test map: %s", cudaGetErrorString(e)); return SEM_ERR_CUDA; }
    return SEM_OK;
}

extern "C" {

int sem_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t *d = nullptr;
    SEM_CUDA(cudaMalloc(&d, 16));
    k_philox<<<1, 1>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_philox_key(key[0], key[1]), d);
    cudaError_t e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("philox test: %s", cudaGetErrorString(e)); return SEM_ERR_CUDA; }
    return SEM_OK;
}
int sem_test_binom_logpmf(const double *k, const double *n, const double *p, double *out, int64_t count) {
    return host_map3(k, n, p, out, count, [&](double *a, double *b, double *c, double *o) { k_binom<<<(unsigned)((count + 127) / 128), 128>>>(a, b, c, o, count); });
}
int sem_test_norm_logpdf(const double *y, const double *x, const double *probs, double *out, int64_t count) {
    return host_map3(y, x, probs, out, count, [&](double *a, double *b, double *c, double *o) { k_norm<<<(unsigned)((count + 127) / 128), 128>>>(a, b, c, o, count); });
}
int sem_test_fast_math(const double *x, const double *a, double *neglog_out, double *rcp_out, int64_t count) {
    double *d = nullptr;
    SEM_CUDA(cudaMalloc(&d, (size_t)count * 4 * sizeof(double)));
    cudaMemcpy(d, x, count * 8, cudaMemcpyHostToDevice); cudaMemcpy(d + count, a, count * 8, cudaMemcpyHostToDevice);
    k_fast_math<<<(unsigned)((count + 127) / 128), 128>>>(d, d + count, d + 2 * count, d + 3 * count, count);
    cudaMemcpy(neglog_out, d + 2 * count, count * 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaMemcpy(rcp_out, d + 3 * count, count * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("fast math test: %s", cudaGetErrorString(e)); return SEM_ERR_CUDA; }
    return SEM_OK;
}
int sem_test_poisson(double mu, uint64_t seed, uint32_t domain, uint32_t c2, double *out, int64_t count) {
    double *d = nullptr;
    SEM_CUDA(cudaMalloc(&d, (size_t)count * 8));
    k_poisson<<<(unsigned)((count + 127) / 128), 128>>>(mu, make_philox_key(seed), domain, c2, d, count);
    cudaError_t e = cudaMemcpy(out, d, count * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("poisson test: %s", cudaGetErrorString(e)); return SEM_ERR_CUDA; }
    return SEM_OK;
}

}  // extern "C"
