/*
 * sem_oracle.c -- CPU oracle (plain C restatement) of the reference's particle-filter / SSA / ABC path.
 *
 * TEST INFRASTRUCTURE ONLY.  Linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by the product library.
 *
 * Parity status: PINNED.  Replay mode is checked bit-for-bit (integer states, ancestor indices) and to
 * 1e-12 (likelihoods) against tests/golden/*.npz, which were produced by running the unmodified reference
 * (GeorgeEfstathiadis/Stochastic-Epidemic-Modelling) under fixed seeds (tests/golden/make_golden.py).
 * The reference ships no golden vectors of its own and pins no numpy/scipy versions; the pins are to
 * numpy 2.3.5 / scipy 1.18.1 (tests/golden/manifest.json).
 *
 * What is restated (reference file:line):
 *   gillespie_algo.py:10-75    sir_simulate        -> ssa_event_* with model 0
 *   gillespie_algo.py:78-146   seir_simulate       -> model 1
 *   gillespie_algo.py:148-233  sir_subgroups_simulate -> model 2/3 (G groups)
 *   pmcmc.py:123-233           particle_filter     -> so_pf_run
 *   pmcmc.py:236-248           particle_path_sampler -> so_path_sample
 *   abc_algo.py:10-99          one ABC trial       -> so_abc_trials
 * Third-party arithmetic restated (not under /root/reference):
 *   numpy legacy RandomState: exponential(scale) = -log(1.0-u)*scale; choice(p) = searchsorted(cumsum(p)/last,u,'right')
 *   scipy.stats.binom.pmf (Boost.Math) -> Loader's saddle-point dbinom (C. Loader 2000, "Fast and accurate
 *   computation of binomial probabilities"), pinned to scipy's values in tests/golden/weights_known_answers.npz
 *   scipy.stats.norm.pdf -> exp(-z^2/2)/sqrt(2 pi)/scale
 *
 * Two uniform sources:
 *   replay : doubles handed in by the caller in the reference's consumption order (arith = REF reproduces the
 *            reference's fp64 operation order exactly)
 *   philox : Philox4x32-10 counter streams, the production RNG of the CUDA kernels; key/counter layout is
 *            the specification in DESIGN.md ("RNG streams") and is implemented here independently.
 * Two arithmetic orders: REF (reference's own operation order) and FAST (algebraically equal, fewer divisions);
 * both are exact Gillespie direct-method SSA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SO_MAX_G 4
#define SO_MAX_C (3 * SO_MAX_G)
#define SO_MAX_R (SO_MAX_G * SO_MAX_G + SO_MAX_G)

enum { M_SIR = 0, M_SEIR = 1, M_SUB = 2, M_SUB2 = 3 };
enum { ARITH_REF = 0, ARITH_FAST = 1, ARITH_UNIF = 2, ARITH_FAST32 = 3, ARITH_UNIF32 = 4 };
enum { DOM_SSA = 1, DOM_RESAMPLE = 2, DOM_INIT = 3, DOM_PATH = 4, DOM_ABC_PRIOR = 5, DOM_ABC_SSA = 6, DOM_SIM = 7, DOM_AUX = 8 };

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al., SC'11) */
void so_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* two 32-bit words -> double in [0,1) with 52 random mantissa bits: bits(0x3FF | mantissa) - 1.0 */
static inline double u52(uint32_t lo, uint32_t hi) {
    uint64_t m = (((uint64_t)hi << 32) | lo) >> 12;
    uint64_t b = 0x3FF0000000000000ull | m;
    double d;
    memcpy(&d, &b, 8);
    return d - 1.0;
}

/* one 32-bit word -> double in [0,1) with 32 random bits (the SSA streams of ARITH_FAST32) */
static inline double u32d(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

typedef struct {
    int philox;            /* 0 = replay buffer, 1 = philox */
    int bits32;            /* philox only: event k takes words (2(k&1), 2(k&1)+1) of call k>>1 (ARITH_FAST32) */
    const double *u;       /* replay: next doubles */
    int64_t pos, len;      /* replay cursor / limit */
    int overrun;           /* replay buffer exhausted */
    uint32_t key[2];
    uint32_t c1, c2, c3;   /* fixed counter words; c0 = draw index */
    uint32_t k;
} so_stream;

/* one pair of uniforms = one SSA event's draws */
static inline void stream_pair(so_stream *s, double *u1, double *u2) {
    if (s->philox && s->bits32) {
        uint32_t ctr[4] = {s->k >> 1, s->c1, s->c2, s->c3}, w[4];
        so_philox4x32(ctr, s->key, w);
        *u1 = u32d(w[2 * (s->k & 1)]);
        *u2 = u32d(w[2 * (s->k & 1) + 1]);
        s->k++;
    } else if (s->philox) {
        uint32_t ctr[4] = {s->k++, s->c1, s->c2, s->c3}, w[4];
        so_philox4x32(ctr, s->key, w);
        *u1 = u52(w[0], w[1]);
        *u2 = u52(w[2], w[3]);
    } else {
        if (s->pos + 2 > s->len) { s->overrun = 1; *u1 = 0.5; *u2 = 0.5; return; }
        *u1 = s->u[s->pos]; *u2 = s->u[s->pos + 1];
        s->pos += 2;
    }
}

static void philox_stream(so_stream *s, uint64_t seed, uint32_t c1, uint32_t c2, uint32_t domain, uint32_t fid) {
    memset(s, 0, sizeof(*s));
    s->philox = 1;
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->c1 = c1; s->c2 = c2; s->c3 = (domain << 24) | (fid & 0xFFFFFFu);
}

void so_philox_uniform_pair(uint64_t seed, uint32_t k, uint32_t c1, uint32_t c2, uint32_t domain, uint32_t fid,
                            double *u1, double *u2) {
    so_stream s; philox_stream(&s, seed, c1, c2, domain, fid); s.k = k; stream_pair(&s, u1, u2);
}

/* ------------------------------------------------------------------ SSA events */
typedef struct {
    int model, G, C, R;
    double th[SO_MAX_G * SO_MAX_G + 1]; /* SIR: beta,gamma; SEIR: beta,alpha,gamma; SUB: betas row-major, gamma */
} so_model;

static void model_init(so_model *m, int model, int G, const double *theta) {
    m->model = model;
    if (model == M_SIR) { m->G = 1; m->C = 3; m->R = 2; memcpy(m->th, theta, 2 * sizeof(double)); }
    else if (model == M_SEIR) { m->G = 1; m->C = 4; m->R = 3; memcpy(m->th, theta, 3 * sizeof(double)); }
    else { m->G = G; m->C = 3 * G; m->R = G * G + G; memcpy(m->th, theta, (G * G + 1) * sizeof(double)); }
}

static inline int model_alive(const so_model *m, const double *x) {
    if (m->model == M_SIR) return x[1] > 0;                        /* gillespie_algo.py:48 */
    if (m->model == M_SEIR) return x[1] > 0 || x[2] > 0;           /* :119 */
    double inf = 0; for (int g = 0; g < m->G; g++) inf = inf + x[3 * g + 1];
    return inf > 0;                                                /* :193 */
}

/* population size as the reference sums it (gillespie_algo.py:35,104,176+182) */
static double model_popsize(const so_model *m, const double *x) {
    if (m->model == M_SIR) return x[0] + x[1] + x[2];
    if (m->model == M_SEIR) return x[0] + x[1] + x[2] + x[3];
    double tot = 0;
    for (int g = 0; g < m->G; g++) { double ng = 0 + x[3 * g] + x[3 * g + 1] + x[3 * g + 2]; tot = tot + ng; }
    return tot;
}

/* propensities in the reference's reaction order and operation order */
static inline void model_rates_ref(const so_model *m, const double *x, double N, double *r) {
    const double *th = m->th;
    if (m->model == M_SIR) { r[0] = th[0] * x[0] * x[1] / N; r[1] = th[1] * x[1]; }                 /* :38-39 */
    else if (m->model == M_SEIR) { r[0] = th[0] * x[0] * x[2] / N; r[1] = th[1] * x[1]; r[2] = th[2] * x[2]; } /* :107-109 */
    else {
        int G = m->G, k = 0; double gamma = th[G * G];
        for (int a = 0; a < G; a++) {
            for (int b = 0; b < G; b++) r[k++] = th[a * G + b] * x[3 * b] * x[3 * a + 1] / N;        /* :182 */
            r[k++] = gamma * x[3 * a + 1];                                                          /* :184 */
        }
    }
}

/* FAST order: beta/N hoisted, no per-event division besides tau */
static inline void model_rates_fast(const so_model *m, const double *x, double invN, double *r) {
    const double *th = m->th;
    if (m->model == M_SIR) { r[0] = (th[0] * invN) * x[0] * x[1]; r[1] = th[1] * x[1]; }
    else if (m->model == M_SEIR) { r[0] = (th[0] * invN) * x[0] * x[2]; r[1] = th[1] * x[1]; r[2] = th[2] * x[2]; }
    else {
        int G = m->G, k = 0; double gamma = th[G * G];
        for (int a = 0; a < G; a++) {
            for (int b = 0; b < G; b++) r[k++] = (th[a * G + b] * invN) * x[3 * b] * x[3 * a + 1];
            r[k++] = gamma * x[3 * a + 1];
        }
    }
}

static inline void model_apply(const so_model *m, double *x, int j) {
    if (m->model == M_SIR) { if (j == 0) { x[0] -= 1; x[1] += 1; } else { x[1] -= 1; x[2] += 1; } }
    else if (m->model == M_SEIR) {
        if (j == 0) { x[0] -= 1; x[1] += 1; } else if (j == 1) { x[1] -= 1; x[2] += 1; } else { x[2] -= 1; x[3] += 1; }
    } else {
        int G = m->G, a = j / (G + 1), k = j % (G + 1);
        if (k < G) { x[3 * k] -= 1; x[3 * k + 1] += 1; } else { x[3 * a + 1] -= 1; x[3 * a + 2] += 1; }
    }
}

/* Run the direct method until max_time or extinction.  Returns number of uniform PAIRS drawn (events incl. the
 * discarded overshoot, gillespie_algo.py:62-66).  If times/states given, records accepted events. */
static int64_t ssa_run_unif(const so_model *m, double *x, double max_time, so_stream *s);
static int64_t ssa_run_unif32(const so_model *m, double *x, double max_time, so_stream *s);

static int64_t ssa_run(const so_model *m, double *x, double max_time, int arith, so_stream *s,
                       double *times, double *states, int64_t max_rec, int64_t *n_rec) {
    /* Philox orders only: a state holding a negative count fires no events.  The reference raises
     * ValueError there (negative probabilities in np.random.choice, gillespie_algo.py:63); the speculative / uniformized
     * loops would see a negative total propensity, i.e. time running backwards. */
    if (arith != ARITH_REF) for (int c = 0; c < m->C; c++) if (x[c] < 0) return 0;
    if (arith == ARITH_UNIF && s->philox && !times) return ssa_run_unif(m, x, max_time, s);
    if (arith == ARITH_UNIF32 && s->philox && !times) return ssa_run_unif32(m, x, max_time, s);
    if (arith == ARITH_UNIF) arith = ARITH_FAST;
    if (arith == ARITH_UNIF32) arith = ARITH_FAST32;            /* event times only exist in the direct method */
    if (arith == ARITH_FAST32) { arith = ARITH_FAST; s->bits32 = s->philox; }   /* same arithmetic, 32-bit streams */
    double r[SO_MAX_R], cdf[SO_MAX_R];
    const int R = m->R, C = m->C;
    double N = model_popsize(m, x), invN = 1.0 / N;
    double t = 0.0;
    int64_t pairs = 0, rec = 0;
    if (times && max_rec > 0) { times[0] = 0.0; memcpy(states, x, C * sizeof(double)); rec = 1; }
    while (model_alive(m, x)) {
        double u1, u2, a0 = 0, tau;
        int j = 0;
        if (arith == ARITH_REF) {
            model_rates_ref(m, x, N, r);
            for (int i = 0; i < R; i++) a0 = a0 + r[i];            /* builtin sum(): 0 + r0 + r1 ... */
            if (!(a0 > 0)) break;                                  /* reference would raise in choice(); we stop */
            stream_pair(s, &u1, &u2); pairs++;
            tau = -log(1.0 - u1) * (1 / a0);                       /* legacy exponential */
            double acc = 0;
            for (int i = 0; i < R; i++) { acc = acc + r[i] / a0; cdf[i] = acc; }   /* p = r/a0 ; cumsum */
            for (int i = 0; i < R; i++) cdf[i] = cdf[i] / acc;                       /* cdf /= cdf[-1] */
            for (int i = 0; i < R; i++) j += (cdf[i] <= u2);                        /* searchsorted right */
            if (j > R - 1) j = R - 1;
        } else {
            model_rates_fast(m, x, invN, r);
            for (int i = 0; i < R; i++) a0 = a0 + r[i];
            if (!(a0 > 0)) break;
            stream_pair(s, &u1, &u2); pairs++;
            tau = -log(1.0 - u1) / a0;
            double v = u2 * a0, acc = 0;
            for (int i = 0; i < R - 1; i++) { acc = acc + r[i]; j += (acc <= v); }
        }
        if (s->overrun) break;
        if (t + tau > max_time) break;                             /* gillespie_algo.py:65 */
        t = t + tau;
        model_apply(m, x, j);
        if (times && rec < max_rec) { times[rec] = t; memcpy(states + rec * C, x, C * sizeof(double)); }
        rec++;
    }
    if (n_rec) *n_rec = rec;
    return pairs;
}

/* Standalone simulation from a replay buffer (last_values_only=False semantics).  x: in/out state (C doubles).
 * Returns pairs consumed; *n_rec = events+1 rows recorded (clipped to max_rec in the buffers). */
int64_t so_ssa_replay(int model, int G, double *x, const double *theta, double max_time, int arith,
                      const double *u, int64_t n_u, double *times, double *states, int64_t max_rec, int64_t *n_rec,
                      int *overrun) {
    so_model m; model_init(&m, model, G, theta);
    so_stream s; memset(&s, 0, sizeof(s)); s.u = u; s.len = n_u;
    int64_t p = ssa_run(&m, x, max_time, arith, &s, times, states, max_rec, n_rec);
    if (overrun) *overrun = s.overrun;
    return p;
}

/* Standalone simulation from the philox stream DOM_SIM (item = sim index). */
int64_t so_ssa_philox(int model, int G, double *x, const double *theta, double max_time, int arith,
                      uint64_t seed, uint32_t sim_index, double *times, double *states, int64_t max_rec, int64_t *n_rec) {
    so_model m; model_init(&m, model, G, theta);
    so_stream s; philox_stream(&s, seed, sim_index, 0, DOM_SIM, 0);
    return ssa_run(&m, x, max_time, arith, &s, times, states, max_rec, n_rec);
}

/* ------------------------------------------------------------------ observation log-weights */
static const double SFE[16] = {0.0, 0.08106146679532726, 0.04134069595540929, 0.02767792568499834,
    0.02079067210376509, 0.01664469118982119, 0.01387612882307075, 0.01189670994589177,
    0.01041126526197209, 0.009255462182712733, 0.008330563433362871, 0.007573675487951841,
    0.006942840107209530, 0.006408994188004207, 0.005951370112758848, 0.005554733551962801};

/* stirlerr(n) = log(n!) - log(sqrt(2 pi n) (n/e)^n), integer n >= 0 (Loader 2000); one reciprocal for n >= 16 */
static double stirlerr(double n) {
    const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
    if (n < 16) return SFE[(int)n];
    double inv = 1.0 / n, i2 = inv * inv;
    if (n > 500) return (S0 - S1 * i2) * inv;
    if (n > 80) return (S0 - (S1 - S2 * i2) * i2) * inv;
    if (n > 35) return (S0 - (S1 - (S2 - S3 * i2) * i2) * i2) * inv;
    return (S0 - (S1 - (S2 - (S3 - S4 * i2) * i2) * i2) * i2) * inv;
}

/* bd0(x, np) = x log(x/np) + np - x.  Near x = np (|x-np| < 0.1 (x+np)) Loader's series
 * (x-np) v + 2 x v sum_{j>=1} v^(2j)/(2j+1), v = (x-np)/(x+np), summed as a fixed degree-8 polynomial in v^2
 * (|v| < 0.1: the first dropped term is < 1e-19 of the leading one). */
static double bd0(double x, double np) {
    double d = x - np;
    if (fabs(d) < 0.1 * (x + np)) {
        double v = d / (x + np), v2 = v * v;
        double q = 1.0 / 17;
        q = q * v2 + 1.0 / 15; q = q * v2 + 1.0 / 13; q = q * v2 + 1.0 / 11; q = q * v2 + 1.0 / 9;
        q = q * v2 + 1.0 / 7; q = q * v2 + 1.0 / 5; q = q * v2 + 1.0 / 3;
        return d * v + (2 * x * v) * (v2 * q);
    }
    return x * log(x / np) + np - x;
}

/* log binom.pmf(k | n, p) with scipy's support rules (pmcmc.py:179): k<0, k>n or non-integer k -> 0 */
double so_binom_logpmf(double k, double n, double p) {
    if (!(k >= 0) || k > n || k != floor(k)) return -INFINITY;
    double q = 1 - p;
    if (p == 0) return k == 0 ? 0.0 : -INFINITY;
    if (q == 0) return k == n ? 0.0 : -INFINITY;
    if (k == 0) {
        if (n == 0) return 0.0;
        return p < 0.1 ? -bd0(n, n * q) - n * p : n * log(q);
    }
    if (k == n) return q < 0.1 ? -bd0(n, n * p) - n * q : n * log(p);
    double lc = stirlerr(n) - stirlerr(k) - stirlerr(n - k) - bd0(k, n * p) - bd0(n - k, n * q);
    double lf = 1.8378770664093453 /* log(2 pi) */ + log(k) + log1p(-k / n);
    return lc - 0.5 * lf;
}

/* log norm.pdf(y | loc=x, scale=probs*x+1e-4)  (pmcmc.py:181) */
double so_norm_logpdf(double y, double x, double probs) {
    double sd = probs * x + .0001, z = (y - x) / sd;
    return -0.5 * z * z - log(sd) - 0.9189385332046727;
}

/* per-particle log-weight = min over observed columns (pmcmc.py:179,181).  x: C doubles. */
static double log_weight(int model, int G, int obs_kind, double probs, const double *Yrow, int Cobs, const double *x) {
    double lw = INFINITY;
    /* a particle holding a negative count (S0 = n_population - Poisson(mu) < 0, pmcmc.py:156-169) has no
     * weight: scipy returns nan for binom.pmf(k, n < 0, p) and for a negative scale of norm.pdf, np.random.choice refuses
     * the weights and the reference's filter returns (None, None, None) (pmcmc.py:187-192).  NaN = collapse at this step. */
    const int Call = model == M_SIR ? 3 : model == M_SEIR ? 4 : 3 * G;
    for (int c = 0; c < Call; c++) if (x[c] < 0) return NAN;
    for (int c = 0; c < Cobs; c++) {
        double xc;
        if (model == M_SUB2) { xc = 0; for (int g = 0; g < G; g++) xc = xc + x[3 * g + c]; }   /* pmcmc.py:172-173 */
        else xc = x[c];
        if (Yrow[c] != Yrow[c]) continue;   /* extension: NaN in Y = unobserved column (e.g. the hidden E of SEIR) */
        double l = obs_kind == 0 ? so_binom_logpmf(Yrow[c], xc, probs) : so_norm_logpdf(Yrow[c], xc, probs);
        if (l < lw || l != l) lw = l;
    }
    return lw == INFINITY ? 0.0 : lw;       /* nothing observed: weight 1 */
}

/* ------------------------------------------------------------------ Poisson sampler (philox init domain) */
static double log_factorial(double k) { /* log(k!) via Loader's stirlerr, exact table below 16 */
    if (k < 2) return 0.0;
    return k * log(k) - k + 0.5 * log(6.283185307179586 * k) + stirlerr(k);
}

/* Poisson(mu): mu < 10 sequential inversion with one uniform; else Hormann's PTRS (1993) transformed rejection.
 * Draws come as pairs from the stream. */
static double poisson_draw(so_stream *s, double mu) {
    double u1, u2;
    if (!(mu > 0)) return 0.0;
    if (mu < 10) {
        stream_pair(s, &u1, &u2);
        double pk = exp(-mu), F = pk, k = 0;
        while (u1 > F && k < 1000) { k += 1; pk *= mu / k; F += pk; }
        return k;
    }
    double slam = sqrt(mu), loglam = log(mu), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
    for (;;) {
        stream_pair(s, &u1, &u2);
        double U = u1 - 0.5, V = u2, us = 0.5 - fabs(U);
        double k = floor((2 * a / us + b) * U + mu + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -mu + k * loglam - log_factorial(k)) return k;
    }
}

double so_poisson_philox(double mu, uint64_t seed, uint32_t c1, uint32_t c2, uint32_t domain, uint32_t fid) {
    so_stream s; philox_stream(&s, seed, c1, c2, domain, fid);
    return poisson_draw(&s, mu);
}

/* ------------------------------------------------------------------ uniformized interval (ARITH_UNIF)
 * Exact law of the state at the end of the interval without waiting times (uniformization + restart rule); the
 * specification is the comment above ssa_run_unif in csrc/sem_common.cuh -- this is its independent C statement.
 * Candidates: the particle's SSA Philox stream, two 52-bit uniforms per call.  K ~ Poisson(B t_rem) and the
 * Marsaglia-Tsang gammas of the Beta(m, K-m+1) restart time: the DOM_AUX stream of the same particle/step. */
static double gamma_draw(so_stream *aux, double shape) {
    double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double u1, u2, u3, u4;
        stream_pair(aux, &u1, &u2);
        stream_pair(aux, &u3, &u4);
        double n = sqrt(-2.0 * log(1.0 - u1)) * cos(3.141592653589793 * (2.0 * u2));
        double v = 1.0 + c * n;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(1.0 - u3) < 0.5 * n * n + d - d * v + d * log(v)) return d * v;
    }
}

/* The candidate count of a uniformized32 batch: PTRS as above with the slow path's logarithms folded into three
 * (csrc/sem_common.cuh poisson_draw_u, which reads them from its table; libm here). */
static double poisson_draw_u(so_stream *s, double mu) {
    double u1, u2;
    if (!(mu > 0)) return 0.0;
    if (mu < 10) {
        stream_pair(s, &u1, &u2);
        double pk = exp(-mu), F = pk, k = 0;
        while (u1 > F && k < 1000) { k += 1; pk *= mu / k; F += pk; }
        return k;
    }
    double slam = sqrt(mu), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    double vr = 0.9277 - 3.6224 / (b - 2);
    for (;;) {
        stream_pair(s, &u1, &u2);
        double U = u1 - 0.5, V = u2, us = 0.5 - fabs(U);
        double k = floor((2 * a / us + b) * U + mu + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0 || (us < 0.013 && V > us)) continue;
        double invalpha = 1.1239 + 1.1328 / (b - 3.4);
        double q = V * invalpha / (a / (us * us) + b);
        if (k < 2 || !(q >= 2.3e-308)) {
            if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -mu + k * log(mu) - log_factorial(k)) return k;
            continue;
        }
        double rhs = (k - mu) + k * log(mu / k) - 0.5 * log(6.283185307179586 * k) - stirlerr(k);
        if (log(q) <= rhs) return k;
    }
}

double so_poisson_u_philox(double mu, uint64_t seed, uint32_t c1, uint32_t c2, uint32_t domain, uint32_t fid) {
    so_stream s; philox_stream(&s, seed, c1, c2, domain, fid);
    return poisson_draw_u(&s, mu);
}

/* diagnostics (single-threaded use): batches, violations, candidates, fired, direct fallbacks */
int64_t so_unif_counters[5];

static int64_t ssa_run_unif(const so_model *m, double *x, double max_time, so_stream *s) {
    const double c0 = 1.25, c1 = 3.0, direct_below = 2.0;
    so_stream aux = *s;
    aux.k = 0; aux.c3 = (s->c3 & 0xFFFFFFu) | ((uint32_t)DOM_AUX << 24);
    const int R = m->R;
    double r[SO_MAX_R], N = model_popsize(m, x), invN = 1.0 / N, t_rem = max_time;
    int64_t fired = 0;
    while (model_alive(m, x)) {
        double a0 = 0;
        model_rates_fast(m, x, invN, r);
        for (int i = 0; i < R; i++) a0 = a0 + r[i];
        if (!(a0 > 0)) break;
        double expect = a0 * t_rem;
        if (expect < direct_below) { so_unif_counters[4]++; return fired + ssa_run(m, x, t_rem, ARITH_FAST, s, NULL, NULL, 0, NULL); }
        so_unif_counters[0]++;
        double B = a0 * (c0 + c1 / sqrt(expect + 1.0));
        double K = poisson_draw(&aux, B * t_rem), done = 0.0;
        int violated = 0, half = 0;
        uint32_t w[4] = {0, 0, 0, 0};
        while (done < K) {
            if (!half) { uint32_t ctr[4] = {s->k++, s->c1, s->c2, s->c3}; so_philox4x32(ctr, s->key, w); }
            double u = half ? u52(w[2], w[3]) : u52(w[0], w[1]);
            half = !half;
            double v = fma(u + 1.0, B, -B);                    /* u*B with one rounding, as the kernel's fma(d,B,-B) */
            done += 1.0;
            so_unif_counters[2]++;
            if (v < a0) {
                so_unif_counters[3]++;
                double acc = r[0]; int j = (acc <= v);
                for (int i = 1; i < R - 1; i++) { acc = acc + r[i]; j += (acc <= v); }
                model_apply(m, x, j);
                fired++;
                if (!model_alive(m, x)) break;
                model_rates_fast(m, x, invN, r);
                a0 = 0; for (int i = 0; i < R; i++) a0 = a0 + r[i];
                if (a0 > B) { violated = 1; break; }
            }
        }
        if (!violated) break;
        so_unif_counters[1]++;
        double g1 = gamma_draw(&aux, done), g2 = gamma_draw(&aux, K - done + 1.0);
        t_rem = t_rem - t_rem * (g1 / (g1 + g2));
        if (!(t_rem > 0)) break;
    }
    return fired;
}

/* Uniformized interval with 32-bit candidates (arith 4; specification: the comment above ssa_unif32_leg in
 * csrc/sem_common.cuh).  Candidate c of the particle-step = word (c & 3) of Philox call c >> 2 of the SSA stream,
 * compared with the fixed-point thresholds 2^32 (r_0 + .. + r_j) / B; bound B = max(a0(x), a0(x + drift * t_rem)) * (c0 + c1 / sqrt(a0 t_rem + 1)); K and the gammas of the
 * restart time come from the DOM_AUX stream (52-bit pairs) as in ssa_run_unif; no direct-method tail. */
static void model_drift(const so_model *m, const double *x, const double *r, double t, double *xp) {
    if (m->model == M_SIR) {
        double f0 = r[0] * t, f1 = r[1] * t;
        xp[0] = fmax(x[0] - f0, 0.0); xp[1] = fmax(x[1] + (f0 - f1), 0.0); xp[2] = x[2];
    } else if (m->model == M_SEIR) {
        double f0 = r[0] * t, f1 = r[1] * t, f2 = r[2] * t;
        xp[0] = fmax(x[0] - f0, 0.0); xp[1] = fmax(x[1] + (f0 - f1), 0.0); xp[2] = fmax(x[2] + (f1 - f2), 0.0); xp[3] = x[3];
    } else {
        int G = m->G;
        for (int b = 0; b < G; b++) {
            double inflow = 0.0;
            for (int a = 0; a < G; a++) inflow = inflow + r[a * (G + 1) + b];
            double fS = inflow * t, fR = r[b * (G + 1) + G] * t;
            xp[3 * b] = fmax(x[3 * b] - fS, 0.0); xp[3 * b + 1] = fmax(x[3 * b + 1] + (fS - fR), 0.0); xp[3 * b + 2] = x[3 * b + 2];
        }
    }
}

/* Candidate-count tables (specification: "candidate-count tables" in csrc/sem_common.cuh, construction as in
 * csrc/sem_host.h): the mean B h of a batch is rounded up to the next double with six mantissa bits (641 grid means in
 * [4, 4096]); K is drawn from the alias table (Vose) of Poisson(grid mean) over mu +- 10 sigma with one pair of 52-bit
 * uniforms of the DOM_AUX stream.  Means above 4096 keep poisson_draw_u. */
#define KTAB_COUNT 641
typedef struct { double prob; int32_t alias; } ktab_entry;
static ktab_entry *g_kt_e = NULL;
static int32_t g_kt_meta[KTAB_COUNT][3];                   /* first entry, entries, first count */

static double ktab_grid_mean(int id) {
    uint64_t bits = (uint64_t)(((uint32_t)(0x40100000u >> 14) + (uint32_t)id) << 14) << 32;
    double mu; memcpy(&mu, &bits, 8);
    return mu;
}

static void ktab_build(void) {
    size_t cap = 0, used = 0;
    for (int id = 0; id < KTAB_COUNT; id++) { double mu = ktab_grid_mean(id); cap += (size_t)(20.0 * sqrt(mu)) + 24; }
    ktab_entry *e = (ktab_entry *)malloc(cap * sizeof(ktab_entry));
    double *p = (double *)malloc(4096 * sizeof(double)), *q = (double *)malloc(4096 * sizeof(double));
    int *small = (int *)malloc(8192 * sizeof(int)), *large = (int *)malloc(4096 * sizeof(int));
    for (int id = 0; id < KTAB_COUNT; id++) {
        const double mu = ktab_grid_mean(id), sd = sqrt(mu);
        long long k_lo = (long long)floor(mu - 10.0 * sd) - 4, k_hi = (long long)ceil(mu + 10.0 * sd) + 12;
        if (k_lo < 0) k_lo = 0;
        const int n = (int)(k_hi - k_lo + 1);
        const long long mode = (long long)floor(mu);
        p[mode - k_lo] = exp((double)mode * log(mu) - mu - lgamma((double)mode + 1.0));
        for (long long k = mode; k < k_hi; k++) p[k + 1 - k_lo] = p[k - k_lo] * mu / (double)(k + 1);
        for (long long k = mode; k > k_lo; k--) p[k - 1 - k_lo] = p[k - k_lo] * (double)k / mu;
        double sum = 0.0;
        for (int i = 0; i < n; i++) sum += p[i];
        int ns = 0, nl = 0, si = 0, li = 0;
        for (int i = 0; i < n; i++) {
            q[i] = p[i] * (double)n / sum;
            if (q[i] < 1.0) small[ns++] = i; else large[nl++] = i;
            e[used + i].prob = 1.0; e[used + i].alias = i;
        }
        while (si < ns && li < nl) {
            const int a = small[si++], g = large[li];
            e[used + a].prob = q[a]; e[used + a].alias = g;
            q[g] = (q[g] + q[a]) - 1.0;
            if (q[g] < 1.0) { small[ns++] = g; li++; }
        }
        g_kt_meta[id][0] = (int32_t)used; g_kt_meta[id][1] = n; g_kt_meta[id][2] = (int32_t)k_lo;
        used += (size_t)n;
    }
    free(p); free(q); free(small); free(large);
    g_kt_e = e;
}

static double ktab_round_up(double mu, int *id) {
    uint64_t bits; memcpy(&bits, &mu, 8);
    const uint32_t hi = (uint32_t)(bits >> 32), lo = (uint32_t)bits;
    if (!(mu > 4.0)) { *id = 0; return 4.0; }
    const int on_grid = (hi & 0x3FFFu) == 0u && lo == 0u;
    *id = (int)((hi >> 14) - (0x40100000u >> 14)) + (on_grid ? 0 : 1);
    return ktab_grid_mean(*id);
}

static uint32_t ktab_draw(so_stream *aux, int id) {
    if (!g_kt_e) {
#pragma omp critical(sem_ktab)
        { if (!g_kt_e) ktab_build(); }
    }
    double u1, u2;
    stream_pair(aux, &u1, &u2);
    const int n = g_kt_meta[id][1];
    int col = (int)(u1 * (double)n);
    if (col > n - 1) col = n - 1;
    const ktab_entry *e = g_kt_e + g_kt_meta[id][0] + col;
    return (uint32_t)(g_kt_meta[id][2] + (u2 < e->prob ? col : e->alias));
}

/* K of one batch (test hook) */
uint32_t so_ktab_philox(double mu, uint64_t seed, uint32_t c1, uint32_t c2, uint32_t domain, uint32_t fid, double *grid_mean) {
    so_stream s;
    philox_stream(&s, seed, c1, c2, domain, fid);
    int id;
    const double g = ktab_round_up(mu, &id);
    if (grid_mean) *grid_mean = g;
    return ktab_draw(&s, id);
}

/* Fixed-point thresholds of the candidate test (Model::thresholds in csrc/sem_common.cuh): T[j] = 2^52 + round(s (r_0 +
 * ... + r_j)), chained fused multiply-adds in the reaction order of model_rates_fast, s = 2^32 / B.  The low word of
 * T[j] is the integer the 32-bit candidate word is compared with. */
#define MAGIC52 4503599627370496.0
#define MAGIC52_HI 0x43300000u
static inline uint32_t lo32(double d) { uint64_t b; memcpy(&b, &d, 8); return (uint32_t)b; }
static inline uint32_t hi32(double d) { uint64_t b; memcpy(&b, &d, 8); return (uint32_t)(b >> 32); }

static void model_thresholds(const so_model *m, const double *x, double invN, double s, double *T) {
    const double *th = m->th;
    if (m->model == M_SIR) {
        T[0] = fma(x[1], ((th[0] * invN) * s) * x[0], MAGIC52);
        T[1] = fma(th[1] * s, x[1], T[0]);
    } else if (m->model == M_SEIR) {
        T[0] = fma(x[2], ((th[0] * invN) * s) * x[0], MAGIC52);
        T[1] = fma(th[1] * s, x[1], T[0]);
        T[2] = fma(th[2] * s, x[2], T[1]);
    } else {
        int G = m->G, k = 0; double gs = th[G * G] * s, acc = MAGIC52;
        for (int a = 0; a < G; a++) {
            for (int b = 0; b < G; b++) { acc = fma(x[3 * a + 1], ((th[a * G + b] * invN) * s) * x[3 * b], acc); T[k++] = acc; }
            acc = fma(gs, x[3 * a + 1], acc); T[k++] = acc;
        }
    }
}

static int64_t ssa_run_unif32(const so_model *m, double *x, double max_time, so_stream *s) {
    const double c0 = 1.0, c1 = 2.0, gmax = 1.25;             /* SEM_U32_C0 / SEM_U32_C1 / SEM_U32_GMAX */
    so_stream aux = *s;
    aux.k = 0; aux.c3 = (s->c3 & 0xFFFFFFu) | ((uint32_t)DOM_AUX << 24);
    const int R = m->R;
    double r[SO_MAX_R], rp[SO_MAX_R], xp[SO_MAX_C], T[SO_MAX_R], N = model_popsize(m, x), invN = 1.0 / N, t_rem = max_time;
    int64_t fired = 0;
    uint32_t cand = 0;
    while (model_alive(m, x)) {
        double a0 = 0, a0p = 0;
        model_rates_fast(m, x, invN, r);
        for (int i = 0; i < R; i++) a0 = a0 + r[i];
        if (!(a0 > 0)) break;
        model_drift(m, x, r, t_rem, xp);
        model_rates_fast(m, xp, invN, rp);
        for (int i = 0; i < R; i++) a0p = a0p + rp[i];
        double amax = a0p > a0 ? a0p : a0, h = t_rem, cap = gmax * a0;
        if (amax > cap) { h = t_rem * ((cap - a0) / (a0p - a0)); amax = cap; }     /* fast growth: a shorter batch */
        double expect = a0 * h;
        double B = amax * (c0 + c1 / sqrt(expect + 1.0));
        double mu = B * h;
        uint32_t K;
        if (mu <= 4096.0) { int id; mu = ktab_round_up(mu, &id); K = ktab_draw(&aux, id); }
        else { double Kd = poisson_draw_u(&aux, mu); K = Kd < 2.0e9 ? (uint32_t)Kd : 2000000000u; }
        uint32_t first = cand, last = cand + K;
        int violated = 0, absorbed = 0;
        const double sc = (4294967296.0 * h) / mu;             /* 2^32 / B for the bound B = mu / h actually used */
        model_thresholds(m, x, invN, sc, T);
        while (cand < last) {
            uint32_t ctr[4] = {cand >> 2, s->c1, s->c2, s->c3}, w[4];
            so_philox4x32(ctr, s->key, w);
            const uint32_t wq = w[cand & 3u];
            cand++;
            if (wq < lo32(T[R - 1])) {                         /* a real event: reaction j */
                int j = 0;
                for (int i = 0; i < R - 1; i++) j += (wq >= lo32(T[i]));
                model_apply(m, x, j);
                fired++;
                model_thresholds(m, x, invN, sc, T);
                if (!(hi32(T[R - 1]) == MAGIC52_HI && lo32(T[R - 1]) != 0u)) {
                    violated = hi32(T[R - 1]) != MAGIC52_HI; absorbed = !violated; break;
                }
            }
        }
        if (!violated) {                                       /* the batch covered its h exactly */
            if (absorbed || !(h < t_rem)) break;
            t_rem = t_rem - h;
            continue;
        }
        uint32_t done = cand - first;
        double g1 = gamma_draw(&aux, (double)done), g2 = gamma_draw(&aux, (double)(K - done) + 1.0);
        t_rem = t_rem - h * (g1 / (g1 + g2));
        if (!(t_rem > 0)) break;
    }
    s->k = (cand + 3) >> 2;
    return fired;
}

/* ------------------------------------------------------------------ particle filter */
typedef struct {
    int32_t model, obs_kind, resampler /*0 multinomial, 1 systematic*/, arith, philox;
    int32_t N, T, G, Cobs, init_poisson, n_threads, pad;
    double probs, dt;
    uint64_t seed;
    uint32_t filter_id, pad2;
} so_pf_cfg;

/*
 * pmcmc.py:123-233.  State layout (T,N,C) like the reference.  Replay inputs:
 *   flat_u != NULL   : one flat stream consumed exactly like numpy's global stream with jobs=1
 *                      ([N resample doubles][pairs of particle 0][pairs of particle 1]... per step); the
 *                      consumption map is written to res_u_out / ssa_off_out / ssa_u is flat_u itself.
 *   flat_u == NULL   : per-particle CSR buffers: res_u[(T-1)*N], ssa_u + ssa_off[(T-1)*N+1] (in doubles).
 * Philox inputs: seed, filter_id; X0 given, or init_poisson with mu[G], npop[G] (pmcmc.py:156-169).
 * Outputs: X_hist (T*N*C int32), ancestry (T*N int32), log_zetas[T], log_w (T*N, row p = weights used at
 * step p like the reference's weights[p]); returns 0, or p>0 = collapsed at step p (pmcmc.py:191-192),
 * or -1 = replay buffer exhausted.
 */
int so_pf_run(const so_pf_cfg *cfg, const double *Y, const double *theta, const int32_t *X0,
              const double *mu, const double *npop,
              const double *flat_u, int64_t flat_len, int64_t *flat_used,
              const double *res_u, const double *ssa_u, const int64_t *ssa_off,
              double *res_u_out, int64_t *ssa_off_out,
              int32_t *X_hist, int32_t *ancestry, double *log_zetas, double *log_w, int64_t *n_events) {
    const int N = cfg->N, T = cfg->T, G = cfg->G;
    so_model m; model_init(&m, cfg->model, G, theta);
    const int C = m.C, Cobs = cfg->Cobs;
    double *cur = (double *)malloc(sizeof(double) * N * C), *nxt = (double *)malloc(sizeof(double) * N * C);
    double *lw = (double *)malloc(sizeof(double) * N), *cdf = (double *)malloc(sizeof(double) * N);
    int64_t events = 0, cursor = 0;
    int rc = 0;
#ifdef _OPENMP
    int nthr = cfg->n_threads > 0 ? cfg->n_threads : omp_get_max_threads();
    if (flat_u) nthr = 1;
#endif
    /* X_0 */
    for (int j = 0; j < N; j++) {
        double *x = cur + (size_t)j * C;
        if (cfg->init_poisson) {
            memset(x, 0, sizeof(double) * C);
            for (int g = 0; g < (cfg->model >= M_SUB ? G : 1); g++) {
                so_stream s; philox_stream(&s, cfg->seed, (uint32_t)j, (uint32_t)g, DOM_INIT, cfg->filter_id);
                double i0 = poisson_draw(&s, mu[g]);
                int icol = cfg->model == M_SEIR ? 2 : 3 * g + 1;       /* pmcmc.py:157,161,167 */
                x[icol] = i0; x[3 * g * (cfg->model >= M_SUB)] = npop[g] - i0;
            }
        } else for (int c = 0; c < C; c++) x[c] = X0[(size_t)j * C + c];
        for (int c = 0; c < C; c++) X_hist[(size_t)j * C + c] = (int32_t)x[c];
        ancestry[j] = 0;
    }
    log_zetas[0] = 0.0;
    if (ssa_off_out) ssa_off_out[0] = 0;
    for (int p = 1; p < T; p++) {
        /* weights of X[p-1] against Y[p-1] */
        double M = -INFINITY; int bad = 0;
        for (int j = 0; j < N; j++) {
            lw[j] = log_weight(cfg->model, G, cfg->obs_kind, cfg->probs, Y + (size_t)(p - 1) * Cobs, Cobs, cur + (size_t)j * C);
            if (lw[j] != lw[j]) bad = 1;
            if (lw[j] > M) M = lw[j];
            if (log_w) log_w[(size_t)p * N + j] = lw[j];
        }
        if (bad || !(M > -INFINITY)) { rc = p; break; }             /* all weights zero / NaN -> collapse */
        double acc = 0;
        for (int j = 0; j < N; j++) { acc = acc + exp(lw[j] - M); cdf[j] = acc; }
        log_zetas[p] = log_zetas[p - 1] + M + log(acc) - log((double)N);   /* zetas[p]=zetas[p-1]*mean(w), :183 */
        /* resample (pmcmc.py:187-193): first index with cdf > u*total  == searchsorted(cdf/total, u, 'right') */
        double u0 = 0, dummy;
        if (cfg->philox && cfg->resampler == 1)
            so_philox_uniform_pair(cfg->seed, 0, 0, (uint32_t)p, DOM_RESAMPLE, cfg->filter_id, &u0, &dummy);
        for (int j = 0; j < N; j++) {
            double u;
            if (cfg->philox) {
                if (cfg->resampler == 1) u = ((double)j + u0) / (double)N;
                else so_philox_uniform_pair(cfg->seed, 0, (uint32_t)j, (uint32_t)p, DOM_RESAMPLE, cfg->filter_id, &u, &dummy);
            } else if (flat_u) {
                if (cursor >= flat_len) { rc = -1; break; }
                u = flat_u[cursor++];
            } else u = res_u[(size_t)(p - 1) * N + j];
            if (res_u_out) res_u_out[(size_t)(p - 1) * N + j] = u;
            double v = u * acc;
            int lo = 0, hi = N;                                     /* first index with cdf[i] > v */
            while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= v) lo = mid + 1; else hi = mid; }
            if (lo > N - 1) lo = N - 1;
            ancestry[(size_t)p * N + j] = lo;
        }
        if (rc) break;
        /* propagate (pmcmc.py:195-225) */
        int fail = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : events) reduction(| : fail) num_threads(nthr)
#endif
        for (int j = 0; j < N; j++) {
            double *x = nxt + (size_t)j * C;
            memcpy(x, cur + (size_t)ancestry[(size_t)p * N + j] * C, sizeof(double) * C);
            so_stream s;
            if (cfg->philox) philox_stream(&s, cfg->seed, (uint32_t)j, (uint32_t)p, DOM_SSA, cfg->filter_id);
            else {
                memset(&s, 0, sizeof(s));
                if (flat_u) { s.u = flat_u; s.pos = cursor; s.len = flat_len; }
                else { size_t q = (size_t)(p - 1) * N + j; s.u = ssa_u; s.pos = ssa_off[q]; s.len = ssa_off[q + 1]; }
            }
            int64_t pairs = ssa_run(&m, x, cfg->dt, cfg->arith, &s, NULL, NULL, 0, NULL);
            events += pairs;
            if (s.overrun) fail = 1;
            if (flat_u) { cursor = s.pos; }
            if (ssa_off_out) ssa_off_out[(size_t)(p - 1) * N + j + 1] = flat_u ? cursor : 0;
            for (int c = 0; c < C; c++) X_hist[((size_t)p * N + j) * C + c] = (int32_t)x[c];
        }
        if (fail) { rc = -1; break; }
        double *tmp = cur; cur = nxt; nxt = tmp;
    }
    if (flat_used) *flat_used = cursor;
    if (n_events) *n_events = events;
    free(cur); free(nxt); free(lw); free(cdf);
    return rc;
}

/* pmcmc.py:236-248 (bug-compatible) or exact genealogy. X_hist (T,N,C) int32, ancestry (T,N). */
void so_path_sample(const int32_t *X_hist, const int32_t *ancestry, int T, int N, int C, int chosen, int exact,
                    int32_t *traj) {
    for (int c = 0; c < C; c++) traj[(size_t)(T - 1) * C + c] = X_hist[((size_t)(T - 1) * N + chosen) * C + c];
    for (int p = T - 2; p >= 0; p--) {
        chosen = ancestry[(size_t)(exact ? p + 1 : p) * N + chosen];
        for (int c = 0; c < C; c++) traj[(size_t)p * C + c] = X_hist[((size_t)p * N + chosen) * C + c];
    }
}

/* ------------------------------------------------------------------ ABC (abc_algo.py:17-109) */
typedef struct {
    int32_t philox, T, arith, early_reject, n_threads, pad;
    double threshold, prior[4]; /* beta lo,hi ; gamma lo,hi */
    uint64_t seed;
} so_abc_cfg;

/*
 * Runs n_trials independent trials.  Replay: theta[n][2], n_start[n][3], per-trial ssa_u/ssa_off CSR.
 * Philox: trial id = trial0 + i; prior draws + Poisson start from DOM_ABC_PRIOR, SSA from DOM_ABC_SSA.
 * obs: (T,3) [S,I,R].  Row k of a trajectory = state at integer time k (abc_algo.py:58-93), so the simulation
 * only needs to reach T-1 (the reference runs to max_time=T and discards the last day, :23,:93).
 * Outputs per trial: theta_out[n][2], distance[n], traj[n][T][3] (S,I,R int32; may be NULL), events.
 */
int so_abc_trials(const so_abc_cfg *cfg, const double *obs, int64_t n_trials, uint64_t trial0,
                  const double *theta_in, const int64_t *n_start_in, const double *ssa_u, const int64_t *ssa_off,
                  double *theta_out, double *distance, int32_t *traj, int64_t *n_events) {
    const int T = cfg->T;
    int64_t events = 0; int fail = 0;
#ifdef _OPENMP
    int nthr = cfg->n_threads > 0 ? cfg->n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : events) reduction(| : fail) num_threads(nthr)
#endif
    for (int64_t i = 0; i < n_trials; i++) {
        double beta, gamma, x[3];
        so_stream s;
        uint64_t id = trial0 + (uint64_t)i;
        if (cfg->philox) {
            so_stream ps; philox_stream(&ps, cfg->seed, (uint32_t)id, (uint32_t)(id >> 32), DOM_ABC_PRIOR, 0);
            double u1, u2; stream_pair(&ps, &u1, &u2);
            beta = cfg->prior[0] + (cfg->prior[1] - cfg->prior[0]) * u1;      /* abc_algo.py:36 */
            gamma = cfg->prior[2] + (cfg->prior[3] - cfg->prior[2]) * u2;     /* :37 */
            for (int c = 0; c < 3; c++) x[c] = poisson_draw(&ps, (double)(int64_t)obs[c]);   /* :39-40 */
            philox_stream(&s, cfg->seed, (uint32_t)id, (uint32_t)(id >> 32), DOM_ABC_SSA, 0);
            s.bits32 = (cfg->arith == ARITH_FAST32 || cfg->arith == ARITH_UNIF32);   /* (the ABC loop needs event times: direct method) */
        } else {
            beta = theta_in[2 * i]; gamma = theta_in[2 * i + 1];
            for (int c = 0; c < 3; c++) x[c] = (double)n_start_in[3 * i + c];
            memset(&s, 0, sizeof(s)); s.u = ssa_u; s.pos = ssa_off[i]; s.len = ssa_off[i + 1];
        }
        theta_out[2 * i] = beta; theta_out[2 * i + 1] = gamma;
        double th[2] = {beta, gamma};
        so_model m; model_init(&m, M_SIR, 1, th);
        double N = x[0] + x[1] + x[2], invN = 1.0 / N, t = 0.0, dsum = 0.0, sI = 0.0, sR = 0.0;
        const double t_stop = (double)(T - 1), reject_at = cfg->threshold * 2.0 * T;
        int day = 0, rejected = 0;
        int32_t *tr = traj ? traj + (size_t)i * T * 3 : NULL;
        /* day 0 row */
        #define RECORD_DAY() do { if (tr) { tr[day*3]=(int32_t)x[0]; tr[day*3+1]=(int32_t)x[1]; tr[day*3+2]=(int32_t)x[2]; } \
              sI += fabs(x[1] - obs[day*3+1]); sR += fabs(x[2] - obs[day*3+2]); dsum = sI + sR; day++; } while (0)
        RECORD_DAY();
        while (x[1] > 0 && day < T) {
            double r[2], a0, u1, u2, tau; int j;
            if (cfg->arith == ARITH_REF) {
                model_rates_ref(&m, x, N, r); a0 = 0 + r[0] + r[1];
                if (!(a0 > 0)) break;
                stream_pair(&s, &u1, &u2); events++;
                tau = -log(1.0 - u1) * (1 / a0);
                double c0 = r[0] / a0, c1 = c0 + r[1] / a0; c0 = c0 / c1;
                j = (c0 <= u2);
            } else {
                model_rates_fast(&m, x, invN, r); a0 = 0 + r[0] + r[1];
                if (!(a0 > 0)) break;
                stream_pair(&s, &u1, &u2); events++;
                tau = -log(1.0 - u1) / a0;
                j = (r[0] <= u2 * a0);
            }
            if (s.overrun) { fail = 1; break; }
            if (t + tau > t_stop) break;
            t = t + tau;
            /* days strictly before this event's time keep the pre-event state (ceil(t) > day) */
            while (day < T && (double)day < t) RECORD_DAY();
            model_apply(&m, x, j);
            if (cfg->early_reject && dsum > reject_at) { rejected = 1; break; }
        }
        if (!rejected) while (day < T) RECORD_DAY();
        #undef RECORD_DAY
        /* distance_function (abc_algo.py:10-13): (mean|dI| + mean|dR|)/2 */
        distance[i] = rejected ? INFINITY : (sI / T + sR / T) / 2;
    }
    if (n_events) *n_events = events;
    return fail ? -1 : 0;
}

int so_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
