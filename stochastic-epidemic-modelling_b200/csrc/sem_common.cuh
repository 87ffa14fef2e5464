// sem_common.cuh -- device-side building blocks shared by the particle-filter, ABC and simulate kernels.
//   * Philox4x32-10 counter RNG and the stream layout (DESIGN.md "RNG streams")
//   * uniform sources (Philox / replay buffer)
//   * Poisson sampler for X_0 and the ABC start (pmcmc.py:157-169, abc_algo.py:39-40)
//   * observation log-weights: binomial (Loader saddle point) and normal (pmcmc.py:178-181)
//   * epidemic models: reaction tables of gillespie_algo.py in the reference's reaction order
//   * the Gillespie direct-method loop (gillespie_algo.py:48-70)
// sm_100a only.  All arithmetic that decides an integer outcome uses explicit round-to-nearest intrinsics so
// that FMA contraction can never change a trajectory relative to the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <string.h>

#include "../../include/sem_b200.h"
#include "sem_host.h"

namespace sem {

enum : uint32_t { DOM_SSA = 1, DOM_RESAMPLE = 2, DOM_INIT = 3, DOM_PATH = 4, DOM_ABC_PRIOR = 5, DOM_ABC_SSA = 6, DOM_SIM = 7, DOM_AUX = 8 };

// ------------------------------------------------------------------------------------------ Philox4x32-10
// The key travels as its expanded round-key schedule (rk[2r] = k0 + r*W0, rk[2r+1] = k1 + r*W1), filled on the host:
// as part of a kernel parameter it sits in the constant bank, where LOP3 reads it as a direct operand -- no
// registers and no per-call key arithmetic (the Weyl additions cost 18 integer adds per call otherwise).
struct PhiloxKey { uint32_t rk[20]; };

__host__ __device__ inline PhiloxKey make_philox_key(uint32_t k0, uint32_t k1) {
    PhiloxKey key;
    for (int r = 0; r < 10; r++) { key.rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u; key.rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u; }
    return key;
}
__host__ __device__ inline PhiloxKey make_philox_key(uint64_t seed) { return make_philox_key((uint32_t)seed, (uint32_t)(seed >> 32)); }

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey &key) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;     // IMAD.WIDE.U32
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.rk[2 * r];     // LOP3 with a constant-bank operand
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.rk[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

// two words -> double in [1,2) carrying 52 random mantissa bits; u = d - 1 in [0,1), and 1 - u = 2 - d exactly
__device__ __forceinline__ double bits_to_d12(uint32_t lo, uint32_t hi) {
    return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
}

// one word -> double in [1,2) carrying 32 random bits (the 32-bit streams of SEM_ARITH_FAST32)
__device__ __forceinline__ double word_to_d12(uint32_t w) {
    return __hiloint2double((int)(0x3FF00000u | (w >> 12)), (int)(w << 20));
}

__host__ __device__ __forceinline__ uint32_t stream_word(uint32_t domain, uint32_t fid) { return (domain << 24) | (fid & 0xFFFFFFu); }

// A sequential source of (u1,u2) pairs: one Philox call, or two doubles of a replay buffer.
template <bool REPLAY>
struct PairSource;

template <>
struct PairSource<false> {
    const PhiloxKey *key; uint32_t k, c1, c2, c3;         // key points into the kernel parameters (constant bank)
    __device__ __forceinline__ void init(const PhiloxKey &key_, uint32_t c1_, uint32_t c2_, uint32_t c3_) { key = &key_; k = 0; c1 = c1_; c2 = c2_; c3 = c3_; }
    __device__ __forceinline__ uint4 raw() { return philox4x32_10(k++, c1, c2, c3, *key); }
    __device__ __forceinline__ bool next(double &u1, double &u2) {
        const uint4 w = raw();
        u1 = bits_to_d12(w.x, w.y) - 1.0; u2 = bits_to_d12(w.z, w.w) - 1.0;
        return true;
    }
};

// ------------------------------------------------------------------------------------------ fast fp64 helpers
// -log(x) for x in (0,1] without a division: x = z * 2^k with z in [0.6875, 1.375); a 1024-entry table gives
// (1/c, log c) for the mantissa-bit subinterval of z; r = z/c - 1 (one FMA, |r| < 2^-11) and
// log1p(r) = r - r^2/2 + r^3/3 - r^4/4 (truncation r^5/5 < 6e-18).  Absolute error ~1e-16 + 1e-16*|log x| (checked
// against libm on the GPU by tests/test_gpu_parity.py::test_fast_math).  The table (16 KB) lives in shared memory.
// (global memory, not __constant__: the copy into shared memory indexes it by thread id, which the constant cache
// would serialise)
constexpr int kLogTabBits = 10, kLogTabSize = 1 << kLogTabBits;
static __device__ const double2 kLogTab[kLogTabSize] = {
#include "sem_logtab.inc"
};

__device__ __forceinline__ void load_logtab(double2 *smem_tab) {
    for (int i = threadIdx.x; i < kLogTabSize; i += blockDim.x) smem_tab[i] = kLogTab[i];
}

static __constant__ double kLogPoly[4] = {-1.0 / 4, 1.0 / 3, -1.0 / 2, 0x1.62e42fefa39efp-1 /* ln 2 */};

__device__ __forceinline__ double neg_log_fast(double x, const double2 *tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int tmp = hi - 0x3fe60000;
    const int i = (tmp >> (20 - kLogTabBits)) & (kLogTabSize - 1);
    const int k = tmp >> 20;                                   // arithmetic shift: floor exponent offset (<= 0 here)
    const double z = __hiloint2double(hi - (tmp & 0xfff00000), lo);
    const double2 tc = tab[i];
    const double r = __fma_rn(z, tc.x, -1.0);
    const double w = __fma_rn((double)k, kLogPoly[3], tc.y);     // (I2F beats the 2^52 magic-number conversion here, measured)
    const double r2 = __dmul_rn(r, r);
    double p = __fma_rn(r, kLogPoly[0], kLogPoly[1]);
    p = __fma_rn(r, p, kLogPoly[2]);
    return -__fma_rn(r2, p, __dadd_rn(w, r));
}

// 1/a for normal positive a, no IEEE fix-up path: hardware seed y0 (rcp.approx.ftz.f64 looks at the high word only,
// rel. error e <~ 2^-20), then 1/a = y0 (1 + e + e^2 + ...) cut after e^2 (a 3rd-order step, 2^-60) and rounded:
// <= 1 ulp.  a = 0 gives NaN (0 * inf), which the SSA loops use as "no event".
__device__ __forceinline__ double rcp_nr(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = __fma_rn(-a, y, 1.0);
    const double q = __fma_rn(e, e, e);                      // e + e^2
    return __fma_rn(y, q, y);
}

template <>
struct PairSource<true> {
    const double *u; long long pos, end;
    __device__ __forceinline__ void init(const double *u_, long long pos_, long long end_) { u = u_; pos = pos_; end = end_; }
    __device__ __forceinline__ bool next(double &u1, double &u2) {
        if (pos + 2 > end) { u1 = 0.5; u2 = 0.5; return false; }
        u1 = u[pos]; u2 = u[pos + 1]; pos += 2;
        return true;
    }
};

// ------------------------------------------------------------------------------------------ log-weights
static __constant__ double kSfe[16] = {0.0, 0.08106146679532726, 0.04134069595540929, 0.02767792568499834,
    0.02079067210376509, 0.01664469118982119, 0.01387612882307075, 0.01189670994589177,
    0.01041126526197209, 0.009255462182712733, 0.008330563433362871, 0.007573675487951841,
    0.006942840107209530, 0.006408994188004207, 0.005951370112758848, 0.005554733551962801};

__device__ __forceinline__ double stirlerr(double n) {   // log(n!) - log(sqrt(2 pi n)(n/e)^n), Loader (2000)
    const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
    if (n < 16) return kSfe[(int)n];
    const double inv = 1.0 / n, i2 = inv * inv;
    if (n > 500) return (S0 - S1 * i2) * inv;
    if (n > 80) return (S0 - (S1 - S2 * i2) * i2) * inv;
    if (n > 35) return (S0 - (S1 - (S2 - S3 * i2) * i2) * i2) * inv;
    return (S0 - (S1 - (S2 - (S3 - S4 * i2) * i2) * i2) * i2) * inv;
}

// bd0(x, np) = x log(x/np) + np - x; near x = np Loader's series as a fixed degree-8 polynomial in v^2,
// v = (x-np)/(x+np), |v| < 0.1 (no data-dependent loop, one division)
__device__ __forceinline__ double bd0(double x, double np) {
    const double d = x - np;
    if (fabs(d) < 0.1 * (x + np)) {
        const double v = d / (x + np), v2 = v * v;
        double q = 1.0 / 17;
        q = q * v2 + 1.0 / 15; q = q * v2 + 1.0 / 13; q = q * v2 + 1.0 / 11; q = q * v2 + 1.0 / 9;
        q = q * v2 + 1.0 / 7; q = q * v2 + 1.0 / 5; q = q * v2 + 1.0 / 3;
        return d * v + (2 * x * v) * (v2 * q);
    }
    return x * log(x / np) + np - x;
}

// Division-free, libm-free variants for the observation weights (every particle, every step): quotients through the
// Newton reciprocal, logarithms through the shared-memory table.  Each agrees with the IEEE / libm statement above to a
// few ulp; the weights agree with scipy / Boost to ~1e-12 (tests/test_gpu_parity.py::test_weights_vs_scipy_golden).
__device__ __forceinline__ double log_tab(double x, const double2 *tab) { return -neg_log_fast(x, tab); }   // x > 0, normal

__device__ __forceinline__ double stirlerr_w(double n) {
    const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
    if (n < 16) return kSfe[(int)n];
    const double inv = rcp_nr(n), i2 = inv * inv;
    if (n > 500) return (S0 - S1 * i2) * inv;
    if (n > 80) return (S0 - (S1 - S2 * i2) * i2) * inv;
    if (n > 35) return (S0 - (S1 - (S2 - S3 * i2) * i2) * i2) * inv;
    return (S0 - (S1 - (S2 - (S3 - S4 * i2) * i2) * i2) * i2) * inv;
}

__device__ __forceinline__ double bd0_w(double x, double np, const double2 *tab) {
    const double d = x - np, s = x + np;
    if (fabs(d) < 0.1 * s) {
        const double v = d * rcp_nr(s), v2 = v * v;
        double q = 1.0 / 17;
        q = q * v2 + 1.0 / 15; q = q * v2 + 1.0 / 13; q = q * v2 + 1.0 / 11; q = q * v2 + 1.0 / 9;
        q = q * v2 + 1.0 / 7; q = q * v2 + 1.0 / 5; q = q * v2 + 1.0 / 3;
        return d * v + (2 * x * v) * (v2 * q);
    }
    return x * log_tab(x * rcp_nr(np), tab) + np - x;
}

// log binom.pmf(k | n, p) with scipy's support rules (pmcmc.py:179): k<0, k>n, non-integer k -> -inf
static __device__ __noinline__ double binom_logpmf(double k, double n, double p, const double2 *tab) {
    if (!(k >= 0) || k > n || k != floor(k)) return -CUDART_INF;
    const double q = 1 - p;
    if (p == 0) return k == 0 ? 0.0 : -CUDART_INF;
    if (q == 0) return k == n ? 0.0 : -CUDART_INF;
    if (k == 0) {
        if (n == 0) return 0.0;
        return p < 0.1 ? -bd0_w(n, n * q, tab) - n * p : n * log(q);
    }
    if (k == n) return q < 0.1 ? -bd0_w(n, n * p, tab) - n * q : n * log(p);
    const double nk = n - k;
    const double lc = stirlerr_w(n) - stirlerr_w(k) - stirlerr_w(nk) - bd0_w(k, n * p, tab) - bd0_w(nk, n * q, tab);
    // 0.5 log(2 pi k (n-k)/n): log k + log1p(-k/n), the second as the log of the exactly known ratio (n-k)/n
    const double lf = 1.8378770664093453 + log_tab(k, tab) + log_tab(nk * rcp_nr(n), tab);
    return lc - 0.5 * lf;
}

// The filter's hot form of the same function: the observed count k (one per column and step) is shared by all
// particles, so its terms are prepared once (BinomObs), and the common case 1 <= k < n, 0 < p < 1 runs straight-line:
// 4 Newton reciprocals, two 5-term Stirling tails, two bd0 (series near the mode) and one table logarithm.  Everything
// else (k = 0, k = n, k > n, non-integer k, p in {0,1}) goes to the general function above.
struct BinomObs { double k, sk, lk; bool regular; };

static __device__ const double kSfeG[16] = {0.0, 0.08106146679532726, 0.04134069595540929, 0.02767792568499834,
    0.02079067210376509, 0.01664469118982119, 0.01387612882307075, 0.01189670994589177,
    0.01041126526197209, 0.009255462182712733, 0.008330563433362871, 0.007573675487951841,
    0.006942840107209530, 0.006408994188004207, 0.005951370112758848, 0.005554733551962801};

// stirlerr(n) from inv = 1/n: the full 5-term tail for every n >= 16 (the extra terms of the shorter forms used for
// large n are below 1e-17), the exact table below 16
__device__ __forceinline__ double stirlerr_inv(double n, double inv) {
    const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
    const double i2 = inv * inv;
    const double tail = (S0 - (S1 - (S2 - (S3 - S4 * i2) * i2) * i2) * i2) * inv;
    return n < 16 ? kSfeG[(int)n] : tail;
}

__device__ __forceinline__ BinomObs binom_obs(double k, const double2 *tab) {
    BinomObs o;
    o.k = k;
    o.regular = (k >= 1.0) && (k == floor(k)) && (k < 4.0e15);
    const double kk = o.regular ? k : 1.0;
    o.sk = stirlerr_inv(kk, rcp_nr(kk));
    o.lk = log_tab(kk, tab);
    return o;
}

__device__ __forceinline__ double binom_logpmf_obs(const BinomObs &o, double n, double p, const double2 *tab) {
    const double q = 1 - p;
    if (o.k == 0.0 && n >= 1.0 && p > 0 && q > 0)            // nothing observed: q^n (frequent: a compartment that is still / again empty)
        return p < 0.1 ? -bd0_w(n, n * q, tab) - n * p : n * log_tab(q, tab);
    if (!(o.regular && n > o.k && p > 0 && q > 0)) return binom_logpmf(o.k, n, p, tab);
    const double k = o.k, nk = n - k;
    const double rn = rcp_nr(n), rnk = rcp_nr(nk);
    const double lc = stirlerr_inv(n, rn) - o.sk - stirlerr_inv(nk, rnk) - bd0_w(k, n * p, tab) - bd0_w(nk, n * q, tab);
    const double lf = 1.8378770664093453 + o.lk + log_tab(nk * rn, tab);
    return lc - 0.5 * lf;
}

// log norm.pdf(y | loc = x, scale = probs*x + 1e-4)  (pmcmc.py:181)
__device__ __forceinline__ double norm_logpdf(double y, double x, double probs, const double2 *tab) {
    const double sd = probs * x + .0001;
    if (!(sd > 0)) { const double z = (y - x) / sd; return -0.5 * z * z - log(sd) - 0.9189385332046727; }   // scipy's nan / inf rules
    const double z = (y - x) * rcp_nr(sd);
    return -0.5 * z * z - log_tab(sd, tab) - 0.9189385332046727;
}

__device__ __forceinline__ double log_factorial(double k) {
    if (k < 2) return 0.0;
    return k * log(k) - k + 0.5 * log(6.283185307179586 * k) + stirlerr(k);
}

// Poisson(mu): mu < 10 sequential inversion with one uniform, else Hormann's PTRS transformed rejection.
template <class Src>
__device__ __noinline__ double poisson_draw(Src &src, double mu) {
    double u1, u2;
    if (!(mu > 0)) return 0.0;
    if (mu < 10) {
        src.next(u1, u2);
        double pk = exp(-mu), F = pk, k = 0;
        while (u1 > F && k < 1000) { k += 1; pk *= mu / k; F += pk; }
        return k;
    }
    const double slam = sqrt(mu), loglam = log(mu), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
    for (;;) {
        src.next(u1, u2);
        const double U = u1 - 0.5, V = u2, us = 0.5 - fabs(U);
        const double k = floor((2 * a / us + b) * U + mu + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -mu + k * loglam - log_factorial(k)) return k;
    }
}

// The candidate count of a uniformized batch (every particle, every step): the same PTRS sampler, arranged for the
// 86 % of draws that the squeeze accepts -- nothing but vr is prepared up front, log(mu) and invalpha move into the
// slow path, whose six logarithms are folded into three (log of the quotient; k log(mu/k); log(2 pi k)) and read from
// the shared-memory table.  The C oracle states the same formulas with libm logarithms and IEEE divisions;
// the two can disagree only when the acceptance test is decided within a few ulp (~1e-15 per draw).
template <class Src>
__device__ __noinline__ double poisson_draw_u(Src &src, double mu, const double2 *tab) {
    double u1, u2;
    if (!(mu > 0)) return 0.0;
    if (mu < 10) {
        src.next(u1, u2);
        double pk = exp(-mu), F = pk, k = 0;
        while (u1 > F && k < 1000) { k += 1; pk *= mu / k; F += pk; }
        return k;
    }
    // quotients through the Newton reciprocal (<= 1 ulp from the IEEE division the oracle states)
    const double slam = sqrt(mu), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 * rcp_nr(b - 2);
    for (;;) {
        src.next(u1, u2);
        const double U = u1 - 0.5, V = u2, us = 0.5 - fabs(U);
        const double k = floor((2 * a * rcp_nr(us) + b) * U + mu + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (!(k >= 0) || (us < 0.013 && V > us)) continue;
        const double invalpha = 1.1239 + 1.1328 * rcp_nr(b - 3.4);
        const double q = V * invalpha * rcp_nr(a * rcp_nr(us * us) + b);
        if (k < 2 || !(q >= 2.3e-308)) {                       // (never at these means; stated for completeness)
            if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -mu + k * log(mu) - log_factorial(k)) return k;
            continue;
        }
        const double ik = rcp_nr(k);
        const double serr = k < 16 ? stirlerr(k) : (1.0 / 12 - (1.0 / 360 - (1.0 / 1260 - (1.0 / 1680 - 1.0 / 1188 * (ik * ik)) * (ik * ik)) * (ik * ik)) * (ik * ik)) * ik;
        const double rhs = (k - mu) + k * log_tab(mu * ik, tab) - 0.5 * log_tab(6.283185307179586 * k, tab) - serr;
        if (log_tab(q, tab) <= rhs) return k;
    }
}

// ------------------------------------------------------------------------------------------ candidate-count tables
// The candidate count of a uniformized batch is K ~ Poisson(B h), and the bound B is OURS to choose (any B above the
// propensity is valid): the mean is rounded UP to the next double with six mantissa bits (<= 1.6 % more candidates, 0.6 %
// on average) so that it falls on a grid of 641 values in [4, 4096], each with a precomputed alias table (Walker / Vose)
// of its Poisson law over mu +- 10 sigma (truncated mass < 1e-20, renormalised; built on the host in fp64, sem_host.h).
// A draw is one Philox call, one 16-byte metadata load and one 16-byte entry load -- no rejection loop, no logarithms,
// no divergence: 0.5 us instead of 3.4 us per step of the headline filter.  Means above 4096 keep the PTRS sampler.
struct KTabEntry { double prob; int32_t alias, pad; };      // accept column i with probability prob, else take alias
struct KTab { const KTabEntry *e; const int4 *meta; };      // meta[id] = (first entry, entries, first count, 0)
static __device__ KTab g_ktab;                              // bound per translation unit and device (ktab_bind)
constexpr double kKTabMaxMean = 4096.0;
constexpr int kKTabShift = 14, kKTabBase = 0x40100000 >> kKTabShift, kKTabCount = ((0x40B00000 >> kKTabShift) - kKTabBase) + 1;

// smallest grid mean >= mu (mu <= 4096): returns the grid mean, id = its table
__host__ __device__ inline double ktab_round_up(const double mu, int &id) {
#ifdef __CUDA_ARCH__
    const uint32_t hi = (uint32_t)__double2hiint(mu), lo = (uint32_t)__double2loint(mu);
#else
    uint64_t bits; memcpy(&bits, &mu, 8);
    const uint32_t hi = (uint32_t)(bits >> 32), lo = (uint32_t)bits;
#endif
    if (!(mu > 4.0)) { id = 0; return 4.0; }
    const uint32_t t = hi >> kKTabShift;
    const bool on_grid = (hi & ((1u << kKTabShift) - 1u)) == 0u && lo == 0u;
    id = (int)(t - (uint32_t)kKTabBase) + (on_grid ? 0 : 1);
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)(((uint32_t)kKTabBase + (uint32_t)id) << kKTabShift), 0);
#else
    const uint64_t ob = (uint64_t)(((uint32_t)kKTabBase + (uint32_t)id) << kKTabShift) << 32;
    double out; memcpy(&out, &ob, 8);
    return out;
#endif
}

// Point this translation unit's g_ktab at the current device's tables (once per unit and device); called by the host
// entry points before a kernel that may run uniformized intervals.
static int ktab_bind() {
    static bool bound[64] = {false};
    static std::mutex mu;
    int dev = 0;
    SEM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index out of range"); return SEM_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(mu);
    if (bound[dev]) return SEM_OK;
    const void *e = nullptr, *m = nullptr;
    const int rc = ktab_device(&e, &m);
    if (rc) return rc;
    KTab kt; kt.e = (const KTabEntry *)e; kt.meta = (const int4 *)m;
    SEM_CUDA(cudaMemcpyToSymbol(g_ktab, &kt, sizeof(KTab)));
    bound[dev] = true;
    return SEM_OK;
}

template <class Src>
__device__ __forceinline__ uint32_t ktab_draw(Src &aux, const int id) {
    const int4 mt = __ldg(&g_ktab.meta[id]);
    double u1, u2;
    aux.next(u1, u2);
    const int col = min((int)__dmul_rn(u1, (double)mt.y), mt.y - 1);
    const KTabEntry *ep = g_ktab.e + (size_t)mt.x + col;
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(ep));
    const double prob = __hiloint2double(raw.y, raw.x);
    return (uint32_t)(mt.z + (u2 < prob ? col : raw.z));
}

// ------------------------------------------------------------------------------------------ models
// State is kept as fp64 integers in registers (exact up to 2^53); converted to int32 at observation boundaries.
// rates<ARITH>() fills r[] in the reference's reaction order; REF divides by N per event like
// gillespie_algo.py:38, FAST multiplies by the hoisted beta/N.

constexpr double kMagic52 = 4503599627370496.0;          // 2^52: a double 2^52 + n (0 <= n < 2^32) holds n in its low word
constexpr uint32_t kMagic52Hi = 0x43300000u;

struct SirModel {
    static constexpr int C = 3, R = 2, NTHETA = 2, G = 1;
    double beta, gamma, N, bN;
    __device__ __forceinline__ void setup(const double *th, const double *x) {
        beta = th[0]; gamma = th[1];
        N = __dadd_rn(__dadd_rn(x[0], x[1]), x[2]);                      // gillespie_algo.py:35
        bN = __dmul_rn(beta, __ddiv_rn(1.0, N));
    }
    __device__ __forceinline__ bool alive(const double *x) const { return x[1] > 0; }      // :48
    template <int ARITH>
    __device__ __forceinline__ void rates(const double *x, double *r) const {
        if (ARITH == SEM_ARITH_REFERENCE) r[0] = __ddiv_rn(__dmul_rn(__dmul_rn(beta, x[0]), x[1]), N);   // :38
        else r[0] = __dmul_rn(__dmul_rn(bN, x[0]), x[1]);
        r[1] = __dmul_rn(gamma, x[1]);                                                                  // :39
    }
    // TRACK_R = false leaves the removed count to fix_removed() (R = N - S - I): one fp64 op less per event
    template <bool TRACK_R = true>
    __device__ __forceinline__ void apply(double *x, int j) const {                                     // :43-46
        const bool inf = (j == 0);                                       // +-1 / 0 are exact: one add per compartment
        x[0] = x[0] + (inf ? -1.0 : 0.0);
        x[1] = x[1] + (inf ? 1.0 : -1.0);
        if (TRACK_R) x[2] = x[2] + (inf ? 0.0 : 1.0);
    }
    __device__ __forceinline__ void fix_removed(double *x) const { x[2] = N - x[0] - x[1]; }
    // Fixed-point thresholds of the uniformized candidate test (ssa_unif32_leg): T[j] = 2^52 + round(s (r_0 + .. + r_j)),
    // reaction order of rates(); the low word of T[j] is the integer a 32-bit candidate word is compared with.
    struct Scaled { double bNs, gs; };
    __device__ __forceinline__ void scale(Scaled &sc, const double s) const { sc.bNs = __dmul_rn(bN, s); sc.gs = __dmul_rn(gamma, s); }
    __device__ __forceinline__ void thresholds(const Scaled &sc, const double *x, double *T) const {
        T[0] = __fma_rn(x[1], __dmul_rn(sc.bNs, x[0]), kMagic52);
        T[1] = __fma_rn(sc.gs, x[1], T[0]);
    }
    // apply reaction j if `hit` (a select of the +-1 / 0 increment, no branch): same arithmetic as apply()
    template <bool TRACK_R = true>
    __device__ __forceinline__ void apply_if(double *x, const bool hit, const int j) const {
        const bool inf = hit && j == 0;
        x[0] = x[0] + (inf ? -1.0 : 0.0);
        x[1] = x[1] + (hit ? (inf ? 1.0 : -1.0) : 0.0);
        if (TRACK_R) x[2] = x[2] + ((hit && !inf) ? 1.0 : 0.0);
    }
    // state after the expected net change over a time t (Euler step of the mean-field drift, clamped at 0): the
    // uniformized interval uses it to anticipate growth of the total propensity
    __device__ __forceinline__ void drift(const double *x, const double *r, double t, double *xp) const {
        const double f0 = __dmul_rn(r[0], t), f1 = __dmul_rn(r[1], t);
        xp[0] = fmax(__dsub_rn(x[0], f0), 0.0);
        xp[1] = fmax(__dadd_rn(x[1], __dsub_rn(f0, f1)), 0.0);
        xp[2] = x[2];
    }
};

struct SeirModel {
    static constexpr int C = 4, R = 3, NTHETA = 3, G = 1;
    double beta, alpha, gamma, N, bN;
    __device__ __forceinline__ void setup(const double *th, const double *x) {
        beta = th[0]; alpha = th[1]; gamma = th[2];                                        // :92
        N = __dadd_rn(__dadd_rn(__dadd_rn(x[0], x[1]), x[2]), x[3]);                       // :104
        bN = __dmul_rn(beta, __ddiv_rn(1.0, N));
    }
    __device__ __forceinline__ bool alive(const double *x) const { return x[1] > 0 || x[2] > 0; }   // :119
    template <int ARITH>
    __device__ __forceinline__ void rates(const double *x, double *r) const {
        if (ARITH == SEM_ARITH_REFERENCE) r[0] = __ddiv_rn(__dmul_rn(__dmul_rn(beta, x[0]), x[2]), N);   // :107
        else r[0] = __dmul_rn(__dmul_rn(bN, x[0]), x[2]);
        r[1] = __dmul_rn(alpha, x[1]);                                                                  // :108
        r[2] = __dmul_rn(gamma, x[2]);                                                                  // :109
    }
    template <bool TRACK_R = true>
    __device__ __forceinline__ void apply(double *x, int j) const {                                     // :113-117
        x[0] = x[0] + ((j == 0) ? -1.0 : 0.0);
        x[1] = x[1] + ((j == 0) ? 1.0 : ((j == 1) ? -1.0 : 0.0));
        x[2] = x[2] + ((j == 1) ? 1.0 : ((j == 2) ? -1.0 : 0.0));
        if (TRACK_R) x[3] = x[3] + ((j == 2) ? 1.0 : 0.0);
    }
    __device__ __forceinline__ void fix_removed(double *x) const { x[3] = N - x[0] - x[1] - x[2]; }
    struct Scaled { double bNs, as, gs; };
    __device__ __forceinline__ void scale(Scaled &sc, const double s) const {
        sc.bNs = __dmul_rn(bN, s); sc.as = __dmul_rn(alpha, s); sc.gs = __dmul_rn(gamma, s);
    }
    __device__ __forceinline__ void thresholds(const Scaled &sc, const double *x, double *T) const {
        T[0] = __fma_rn(x[2], __dmul_rn(sc.bNs, x[0]), kMagic52);
        T[1] = __fma_rn(sc.as, x[1], T[0]);
        T[2] = __fma_rn(sc.gs, x[2], T[1]);
    }
    template <bool TRACK_R = true>
    __device__ __forceinline__ void apply_if(double *x, const bool hit, const int j) const {
        const bool r0 = hit && j == 0, r1 = hit && j == 1, r2 = hit && j == 2;
        x[0] = x[0] + (r0 ? -1.0 : 0.0);
        x[1] = x[1] + (r0 ? 1.0 : (r1 ? -1.0 : 0.0));
        x[2] = x[2] + (r1 ? 1.0 : (r2 ? -1.0 : 0.0));
        if (TRACK_R) x[3] = x[3] + (r2 ? 1.0 : 0.0);
    }
    __device__ __forceinline__ void drift(const double *x, const double *r, double t, double *xp) const {
        const double f0 = __dmul_rn(r[0], t), f1 = __dmul_rn(r[1], t), f2 = __dmul_rn(r[2], t);
        xp[0] = fmax(__dsub_rn(x[0], f0), 0.0);
        xp[1] = fmax(__dadd_rn(x[1], __dsub_rn(f0, f1)), 0.0);
        xp[2] = fmax(__dadd_rn(x[2], __dsub_rn(f1, f2)), 0.0);
        xp[3] = x[3];
    }
};

template <int G_>
struct SubModel {
    static constexpr int G = G_, C = 3 * G_, R = G_ * G_ + G_, NTHETA = G_ * G_ + 1;
    double betas[G_ * G_], bN[G_ * G_], gamma, N, Ng[G_];
    __device__ __forceinline__ void setup(const double *th, const double *x) {
#pragma unroll
        for (int i = 0; i < G * G; i++) betas[i] = th[i];                 // betas[a*G+b]: infector a -> susceptible b (:182)
        gamma = th[G * G];
        double tot = 0.0;
#pragma unroll
        for (int g = 0; g < G; g++) {                                      // :176 per-group builtin sum, then sum(N) (:182)
            const double ng = __dadd_rn(__dadd_rn(__dadd_rn(0.0, x[3 * g]), x[3 * g + 1]), x[3 * g + 2]);
            tot = __dadd_rn(tot, ng);
            Ng[g] = ng;
        }
        N = tot;
        const double invN = __ddiv_rn(1.0, N);
#pragma unroll
        for (int i = 0; i < G * G; i++) bN[i] = __dmul_rn(betas[i], invN);
    }
    __device__ __forceinline__ bool alive(const double *x) const {        // :192-193
        double inf = 0.0;
#pragma unroll
        for (int g = 0; g < G; g++) inf = __dadd_rn(inf, x[3 * g + 1]);
        return inf > 0;
    }
    template <int ARITH>
    __device__ __forceinline__ void rates(const double *x, double *r) const {
#pragma unroll
        for (int a = 0; a < G; a++) {
#pragma unroll
            for (int b = 0; b < G; b++) {
                if (ARITH == SEM_ARITH_REFERENCE)
                    r[a * (G + 1) + b] = __ddiv_rn(__dmul_rn(__dmul_rn(betas[a * G + b], x[3 * b]), x[3 * a + 1]), N);
                else
                    r[a * (G + 1) + b] = __dmul_rn(__dmul_rn(bN[a * G + b], x[3 * b]), x[3 * a + 1]);
            }
            r[a * (G + 1) + G] = __dmul_rn(gamma, x[3 * a + 1]);         // :184
        }
    }
    template <bool TRACK_R = true>
    __device__ __forceinline__ void apply(double *x, int j) const {       // :183,185
        const int a = j / (G + 1), k = j - a * (G + 1);
#pragma unroll
        for (int g = 0; g < G; g++) {
            const bool infect = (k == g);                   // susceptible of group g infected (by group a)
            const bool recover = (k == G) && (a == g);
            x[3 * g] = infect ? x[3 * g] - 1.0 : x[3 * g];
            x[3 * g + 1] = infect ? x[3 * g + 1] + 1.0 : (recover ? x[3 * g + 1] - 1.0 : x[3 * g + 1]);
            if (TRACK_R) x[3 * g + 2] = recover ? x[3 * g + 2] + 1.0 : x[3 * g + 2];
        }
    }
    __device__ __forceinline__ void fix_removed(double *x) const {
#pragma unroll
        for (int g = 0; g < G; g++) x[3 * g + 2] = Ng[g] - x[3 * g] - x[3 * g + 1];
    }
    struct Scaled { double bNs[G_ * G_], gs; };
    __device__ __forceinline__ void scale(Scaled &sc, const double s) const {
#pragma unroll
        for (int i = 0; i < G * G; i++) sc.bNs[i] = __dmul_rn(bN[i], s);
        sc.gs = __dmul_rn(gamma, s);
    }
    __device__ __forceinline__ void thresholds(const Scaled &sc, const double *x, double *T) const {
        double acc = kMagic52;
#pragma unroll
        for (int a = 0; a < G; a++) {
#pragma unroll
            for (int b = 0; b < G; b++) { acc = __fma_rn(x[3 * a + 1], __dmul_rn(sc.bNs[a * G + b], x[3 * b]), acc); T[a * (G + 1) + b] = acc; }
            acc = __fma_rn(sc.gs, x[3 * a + 1], acc);
            T[a * (G + 1) + G] = acc;
        }
    }
    template <bool TRACK_R = true>
    __device__ __forceinline__ void apply_if(double *x, const bool hit, const int j) const { if (hit) apply<TRACK_R>(x, j); }
    __device__ __forceinline__ void drift(const double *x, const double *r, double t, double *xp) const {
#pragma unroll
        for (int b = 0; b < G; b++) {
            double inflow = 0.0;                                          // new infections of group b, from every infector group
#pragma unroll
            for (int a = 0; a < G; a++) inflow = __dadd_rn(inflow, r[a * (G + 1) + b]);
            const double fS = __dmul_rn(inflow, t), fR = __dmul_rn(r[b * (G + 1) + G], t);
            xp[3 * b] = fmax(__dsub_rn(x[3 * b], fS), 0.0);
            xp[3 * b + 1] = fmax(__dadd_rn(x[3 * b + 1], __dsub_rn(fS, fR)), 0.0);
            xp[3 * b + 2] = x[3 * b + 2];
        }
    }
};

struct NoRec { __device__ __forceinline__ void operator()(double, const double *) const {} };

// One Gillespie draw (split so that no uniform is consumed when the total propensity is not positive):
//   ssa_total     : propensities r[] in the reference's reaction order and a0 = builtin sum() = 0 + r0 + r1 ...
//   ssa_pick_ref  : numpy's legacy exponential / choice arithmetic (gillespie_algo.py:62-63):
//                   tau = -log(1-u1) * (1/a0);  p = r/a0; cdf = cumsum(p); cdf /= cdf[-1]; j = #(cdf <= u2)
//   ssa_pick_fast : tau = -log(1-u1) / a0 with the table log and a Newton reciprocal; j = #(prefix(r) <= u2*a0).
//                   Takes the Philox doubles d = 1+u in [1,2): 1-u1 = 2-d1 and u2*a0 = fma(d2,a0,-a0), both exact
//                   rewrites.  tau agrees with the IEEE value to ~2 ulp; an integer outcome can differ from the CPU
//                   oracle's FAST order only when t+tau hits the interval end to within those ulps.
template <class Model, int ARITH>
__device__ __forceinline__ double ssa_total(const Model &m, const double *x, double *r) {
    m.template rates<ARITH>(x, r);
    double a0 = (ARITH == SEM_ARITH_REFERENCE) ? __dadd_rn(0.0, r[0]) : r[0];   // 0 + r0 is r0 (rates are never -0)
#pragma unroll
    for (int i = 1; i < Model::R; i++) a0 = __dadd_rn(a0, r[i]);
    return a0;
}

template <class Model>
__device__ __forceinline__ void ssa_pick_ref(const double *r, double a0, double u1, double u2, double &tau, int &j) {
    const double E = -log(__dsub_rn(1.0, u1));
    tau = __dmul_rn(E, __ddiv_rn(1.0, a0));
    double cdf[Model::R], acc = 0.0;
#pragma unroll
    for (int i = 0; i < Model::R; i++) { acc = __dadd_rn(acc, __ddiv_rn(r[i], a0)); cdf[i] = acc; }
    j = 0;
#pragma unroll
    for (int i = 0; i < Model::R; i++) j += (__ddiv_rn(cdf[i], acc) <= u2) ? 1 : 0;
    j = min(j, Model::R - 1);
}

template <class Model>
__device__ __forceinline__ void ssa_pick_fast(const double *r, double a0, double d1, double d2, const double2 *tab,
                                              double &tau, int &j) {
    const double E = neg_log_fast(__dsub_rn(2.0, d1), tab);
    tau = __dmul_rn(E, rcp_nr(a0));
    const double v = __fma_rn(d2, a0, -a0);
    double acc = r[0];
    j = (acc <= v) ? 1 : 0;
#pragma unroll
    for (int i = 1; i < Model::R - 1; i++) { acc = __dadd_rn(acc, r[i]); j += (acc <= v) ? 1 : 0; }
}

// Direct method from t = 0 to max_time (gillespie_algo.py:48-70).  Both uniforms are drawn before the
// overshoot test, so the discarded last event consumes a pair too (:62-66).  Returns pairs drawn, or -1 when a
// replay buffer ran dry.  Rec(t, x) is called after every accepted event.
template <class Model, bool REPLAY, class Rec>
__device__ __forceinline__ long long ssa_run_ref(const Model &m, double *x, double max_time, PairSource<REPLAY> &src, Rec rec) {
    double t = 0.0;
    long long pairs = 0;
    while (m.alive(x)) {
        double r[Model::R], u1, u2, tau;
        int j;
        const double a0 = ssa_total<Model, SEM_ARITH_REFERENCE>(m, x, r);
        if (!(a0 > 0)) break;                                              // (the reference would raise inside choice())
        if (!src.next(u1, u2)) return -1;
        pairs++;
        ssa_pick_ref<Model>(r, a0, u1, u2, tau, j);
        const double tn = __dadd_rn(t, tau);
        if (tn > max_time) break;                                          // :65
        t = tn;
        m.apply(x, j);
        rec(t, x);
    }
    return pairs;
}

// FAST order, Philox only.  The words of event k+1 are generated while event k's fp64 chain runs (they do not
// depend on the state), which gives every thread two independent instruction streams.
template <class Model, bool TRACK_R, class Rec>
__device__ __forceinline__ long long ssa_run_fast(const Model &m, double *x, double max_time, PairSource<false> &src,
                                                  const double2 *tab, Rec rec) {
    double t = 0.0;
    int pairs = 0;                                                         // < 2^31 events per particle-step
    PairSource<false> loc = src;                                           // register copy of the stream state
    uint4 w = loc.raw();
    bool go = m.alive(x);
    // Single basic block per event (the overshoot test is a predicate, not a loop exit) so that the scheduler can
    // interleave the integer Philox chain of event k+1 with the fp64 chain of event k.
    while (go) {
        double r[Model::R], tau;
        int j;
        const double a0 = ssa_total<Model, SEM_ARITH_FAST>(m, x, r);
        const uint4 wn = loc.raw();
        ssa_pick_fast<Model>(r, a0, bits_to_d12(w.x, w.y), bits_to_d12(w.z, w.w), tab, tau, j);
        const double tn = __dadd_rn(t, tau);
        const bool drew = a0 > 0;                                          // no draw when nothing can happen
        const bool fire = drew && !(tn > max_time);                        // gillespie_algo.py:65
        pairs += drew ? 1 : 0;
        if (fire) {
            t = tn;
            m.template apply<TRACK_R>(x, j);
            rec(t, x);
        }
        w = wn;
        go = fire && m.alive(x);
    }
    src.k = loc.k;
    if (!TRACK_R) m.fix_removed(x);
    return pairs;
}

// FAST order, Philox only: the same arithmetic as ssa_run_fast event by event (bit-identical states and draw counts),
// executed U events at a time.  Only the reaction choice feeds back into the state (rates -> a0 -> u2*a0 -> compare ->
// +-1); the reciprocal / logarithm / waiting-time chain is feed-forward, and the interval end is decided by the running
// time alone.  So a block advances the state speculatively through U events (a short dependent chain), evaluates the U
// waiting times with U-fold instruction-level parallelism, and tests the interval end ONCE (times are non-decreasing,
// and a non-positive a0 -- an extinct or frozen state -- makes its time NaN, which poisons every later sum).  The
// block in which the interval ends rolls back to the last event that fired; the Philox draws after it are simply unused
// (streams restart at counter 0 every particle-step, so nothing downstream shifts).
// BITS32 (SEM_ARITH_FAST32): event k takes words (2(k&1), 2(k&1)+1) of Philox call k>>1 -- u1 and u2 carry 32 random
// bits each, one call serves two events.  The 32x32->64 multiplies of Philox are the most expensive instructions of
// the loop on sm_100a (IMAD.WIDE issues once per ~4 cycles, tools/micro/pipes2.cu), so this halves its largest cost.
// One speculative block: from state xs[0] at time t, the states xs[1..U] after each of the next U events (whether or
// not they fit in the interval), their firing times tn[0..U) (non-decreasing; NaN from the first event whose total
// propensity is not positive) and the propensities a0[0..U).  Event i fires iff tn[i] <= max_time.
template <class Model, int U, bool BITS32, bool TRACK_R>
__device__ __forceinline__ void ssa_block(const Model &m, double (&xs)[U + 1][Model::C], const double t, PairSource<false> &src,
                                          const double2 *tab, double (&tn)[U], double (&a0)[U]) {
    static_assert(!BITS32 || U % 2 == 0, "32-bit streams serve two events per call");
    double d1[U], d2[U];
    if constexpr (BITS32) {
#pragma unroll
        for (int i = 0; i < U; i += 2) {
            const uint4 w = src.raw();
            d1[i] = word_to_d12(w.x); d2[i] = word_to_d12(w.y); d1[i + 1] = word_to_d12(w.z); d2[i + 1] = word_to_d12(w.w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < U; i++) { const uint4 w = src.raw(); d1[i] = bits_to_d12(w.x, w.y); d2[i] = bits_to_d12(w.z, w.w); }
    }
#pragma unroll
    for (int i = 0; i < U; i++) {                                          // state chain (speculative)
        double r[Model::R];
        a0[i] = ssa_total<Model, SEM_ARITH_FAST>(m, xs[i], r);
        const double v = __fma_rn(d2[i], a0[i], -a0[i]);
        double acc = r[0];
        int j = (acc <= v) ? 1 : 0;
#pragma unroll
        for (int k = 1; k < Model::R - 1; k++) { acc = __dadd_rn(acc, r[k]); j += (acc <= v) ? 1 : 0; }
#pragma unroll
        for (int c = 0; c < Model::C; c++) xs[i + 1][c] = xs[i][c];
        m.template apply<TRACK_R>(xs[i + 1], j);
    }
    double tt = t;
#pragma unroll
    for (int i = 0; i < U; i++) {                                          // waiting times (feed-forward, independent)
        const double E = neg_log_fast(__dsub_rn(2.0, d1[i]), tab);
        tt = __dadd_rn(tt, __dmul_rn(E, rcp_nr(a0[i])));
        tn[i] = tt;
    }
}

template <class Model, int U, bool BITS32, bool TRACK_R, class Rec>
__device__ __forceinline__ long long ssa_run_spec(const Model &m, double *x, double max_time, PairSource<false> &src,
                                                  const double2 *tab, Rec rec) {
    if (!m.alive(x)) { if (!TRACK_R) m.fix_removed(x); return 0; }
    double t = 0.0;
    int pairs = 0;
    PairSource<false> loc = src;
    double xs[U + 1][Model::C];
#pragma unroll
    for (int c = 0; c < Model::C; c++) xs[0][c] = x[c];
    for (;;) {
        double a0[U], tn[U];
        ssa_block<Model, U, BITS32, TRACK_R>(m, xs, t, loc, tab, tn, a0);
        if (tn[U - 1] <= max_time) {                                       // all U events fired (gillespie_algo.py:65)
            pairs += U;
            t = tn[U - 1];
#pragma unroll
            for (int i = 0; i < U; i++) rec(tn[i], xs[i + 1]);
#pragma unroll
            for (int c = 0; c < Model::C; c++) xs[0][c] = xs[U][c];
            continue;
        }
        int nf = 0;                                                        // the interval ends inside this block
#pragma unroll
        for (int i = 0; i < U; i++) nf += (tn[i] <= max_time) ? 1 : 0;
#pragma unroll
        for (int i = 0; i < U - 1; i++) if (i < nf) rec(tn[i], xs[i + 1]);
        double a_stop = a0[0];
#pragma unroll
        for (int i = 1; i < U; i++) a_stop = (nf == i) ? a0[i] : a_stop;
#pragma unroll
        for (int c = 0; c < Model::C; c++) {
            double v = xs[0][c];
#pragma unroll
            for (int i = 1; i < U; i++) v = (nf == i) ? xs[i][c] : v;
            x[c] = v;
        }
        pairs += nf + ((a_stop > 0) ? 1 : 0);                              // the discarded draw counts when one was made
        break;
    }
    src.k = loc.k;
    if (!TRACK_R) m.fix_removed(x);
    return pairs;
}

// events per speculative block of the production loops (measured, tools/micro/ssa_loop.cu): 4 with 32-bit uniforms
// (two Philox calls in flight), 2 with 52-bit uniforms (4 spills under the 80-register cap); models with many
// compartments keep fewer speculative states in registers (0 = the one-event-per-iteration loop)
// (SEIR: blocks of 4 = a 311-instruction loop, 12.25 ms on config 3; blocks of 2 = 161 instructions, 12.55 ms)
template <class Model> struct SpecBlock { static constexpr int bits32 = Model::C <= 4 ? 4 : 2, bits52 = Model::C <= 6 ? 2 : 0; };

// The same loop in two legs, for a particle whose interval is shared by two warps (pf_persistent's scheduler
// balancing): it runs from time t (0 on the first leg) until the interval ends (finished = true) or -- at a block
// boundary -- until t >= handoff, and returns the continuation (state, t, stream counter).  The blocks stay aligned to
// the same stream counters as in ssa_run_spec, so two legs reproduce the single run bit for bit.
template <class Model, int U, bool BITS32>
__device__ __forceinline__ long long ssa_run_spec_leg(const Model &m, double *x, double &t, const double handoff, const double max_time,
                                                      PairSource<false> &src, const double2 *tab, bool &finished) {
    finished = true;
    if (!m.alive(x)) { m.fix_removed(x); return 0; }
    int pairs = 0;
    PairSource<false> loc = src;
    double xs[U + 1][Model::C];
#pragma unroll
    for (int c = 0; c < Model::C; c++) xs[0][c] = x[c];
    for (;;) {
        double a0[U], tn[U];
        ssa_block<Model, U, BITS32, false>(m, xs, t, loc, tab, tn, a0);
        if (tn[U - 1] <= max_time) {
            pairs += U;
            t = tn[U - 1];
#pragma unroll
            for (int c = 0; c < Model::C; c++) xs[0][c] = xs[U][c];
            if (t >= handoff) {                                            // pass the particle on
#pragma unroll
                for (int c = 0; c < Model::C; c++) x[c] = xs[0][c];
                finished = false;
                break;
            }
            continue;
        }
        int nf = 0;
#pragma unroll
        for (int i = 0; i < U; i++) nf += (tn[i] <= max_time) ? 1 : 0;
        double a_stop = a0[0];
#pragma unroll
        for (int i = 1; i < U; i++) a_stop = (nf == i) ? a0[i] : a_stop;
#pragma unroll
        for (int c = 0; c < Model::C; c++) {
            double v = xs[0][c];
#pragma unroll
            for (int i = 1; i < U; i++) v = (nf == i) ? xs[i][c] : v;
            x[c] = v;
        }
        pairs += nf + ((a_stop > 0) ? 1 : 0);
        break;
    }
    src.k = loc.k;
    m.fix_removed(x);                                                      // keeps S + I + R = N for the next leg's setup
    return pairs;
}

// out-of-line entry (keeps the legs' registers out of the callers' main loop)
template <class Model, int U, bool BITS32>
__device__ __noinline__ long long ssa_run_spec_leg_call(const Model &m, double *x, double &t, const double handoff, const double max_time,
                                                        PairSource<false> &src, const double2 *tab, bool &finished) {
    return ssa_run_spec_leg<Model, U, BITS32>(m, x, t, handoff, max_time, src, tab, finished);
}

template <class Model, int ARITH>
struct LegLoop {                                                           // which (block size, stream width) ssa_run<ARITH> uses
    static constexpr bool available = (ARITH == SEM_ARITH_FAST32) || (ARITH == SEM_ARITH_FAST && SpecBlock<Model>::bits52 > 0);
    static constexpr bool bits32 = ARITH == SEM_ARITH_FAST32;
    static constexpr int U = bits32 ? SpecBlock<Model>::bits32 : (SpecBlock<Model>::bits52 > 0 ? SpecBlock<Model>::bits52 : 2);
};

// ------------------------------------------------------------------------------------------ uniformized interval
// Exact simulation of the state at the end of an interval WITHOUT waiting times (Jensen's uniformization with a
// restart rule).  While the total propensity a0(x) stays <= B, the jump process is a rate-B Poisson stream of
// candidates thinned with probability r_j(x)/B (null event otherwise); only the NUMBER K ~ Poisson(B * t_rem) of
// candidates in the remaining time matters, not their times.  If a fired event lifts a0 above B, the bound was valid
// up to that candidate (the m-th of K): its time is the m-th order statistic of K uniforms, t_rem * Beta(m, K-m+1),
// drawn exactly (ratio of Marsaglia-Tsang gammas); the later candidates are discarded (strong Markov property) and
// the rest of the interval restarts from the current state with a fresh bound.  Per candidate: half a Philox call
// (one 52-bit uniform), no logarithm, no reciprocal.
// Streams: candidates consume the particle's SSA stream (two per call); K and the gammas use the DOM_AUX stream.
template <class Src>
__device__ __noinline__ double gamma_draw(Src &aux, double shape) {        // shape >= 1 (Marsaglia & Tsang 2000)
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double u1, u2, u3, u4;
        aux.next(u1, u2);
        aux.next(u3, u4);
        const double n = sqrt(-2.0 * log(1.0 - u1)) * cos(3.141592653589793 * (2.0 * u2));   // Box-Muller normal
        double v = 1.0 + c * n;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(1.0 - u3) < 0.5 * n * n + d - d * v + d * log(v)) return d * v;
    }
}

// out-of-line copy of the direct loop for the tail of a uniformized interval (keeps its registers out of the batch loop)
template <class Model, bool TRACK_R>
__device__ __noinline__ long long ssa_run_fast_call(const Model &m, double *x, double max_time, PairSource<false> &src,
                                                    const double2 *tab) {
    return ssa_run_fast<Model, TRACK_R>(m, x, max_time, src, tab, NoRec());
}

struct UnifTuning { double c0, c1, direct_below; };
// B = a0 * (c0 + c1 / sqrt(a0 * t_rem + 1)); intervals expecting fewer than direct_below events use the direct method
__device__ __forceinline__ UnifTuning unif_tuning() { return UnifTuning{1.25, 3.0, 2.0}; }


// one candidate of the thinned stream, branch-free: returns true when the batch must stop (absorbed or bound violated)
template <class Model, bool TRACK_R>
__device__ __forceinline__ bool unif_candidate(const Model &m, double *x, double *r, double &a0, const double B, const double d,
                                               int &fired, bool &violated) {
    const double v = __fma_rn(d, B, -B);                                   // u * B
    const bool hit = v < a0;                                               // a real event (else a null candidate)
    double acc = r[0];
    int j = (acc <= v) ? 1 : 0;
#pragma unroll
    for (int i = 1; i < Model::R - 1; i++) { acc = __dadd_rn(acc, r[i]); j += (acc <= v) ? 1 : 0; }
    if (hit) m.template apply<TRACK_R>(x, j);
    fired += hit ? 1 : 0;
    a0 = ssa_total<Model, SEM_ARITH_FAST>(m, x, r);
    const bool alive = m.alive(x);
    violated = hit && alive && (a0 > B);
    return hit && (!alive || a0 > B);
}

template <class Model, bool TRACK_R>
__device__ __forceinline__ long long ssa_run_unif(const Model &m, double *x, double max_time, PairSource<false> &src,
                                                  PairSource<false> &aux, const double2 *tab) {
    const UnifTuning tune = unif_tuning();
    double t_rem = max_time;
    long long total_fired = 0;
    while (m.alive(x)) {
        double r[Model::R];
        double a0 = ssa_total<Model, SEM_ARITH_FAST>(m, x, r);
        if (!(a0 > 0)) break;
        const double expect = a0 * t_rem;
        if (expect < tune.direct_below) {                                  // nearly nothing left: direct method
            total_fired += ssa_run_fast_call<Model, TRACK_R>(m, x, t_rem, src, tab);
            return total_fired;
        }
        const double B = a0 * (tune.c0 + tune.c1 / sqrt(expect + 1.0));
        const double Kd = poisson_draw(aux, B * t_rem);
        const int K = Kd < 2.0e9 ? (int)Kd : 2000000000;
        int done = 0, fired = 0;                                           // candidates processed / events fired in this batch
        bool violated = false, stop = false;
        PairSource<false> loc = src;                                       // register copy (src's address escapes to out-of-line calls)
        while (done < K && !stop) {                                        // two candidates per Philox call
            const uint4 w = loc.raw();
            stop = unif_candidate<Model, TRACK_R>(m, x, r, a0, B, bits_to_d12(w.x, w.y), fired, violated);
            done++;
            if (!stop && done < K) {
                stop = unif_candidate<Model, TRACK_R>(m, x, r, a0, B, bits_to_d12(w.z, w.w), fired, violated);
                done++;
            }
        }
        src.k = loc.k;
        total_fired += fired;
        if (!violated) break;                                              // the batch covered the rest of the interval
        // the bound held up to candidate `done`; its time is the done-th order statistic of K uniforms on [0, t_rem]
        const double g1 = gamma_draw(aux, (double)done), g2 = gamma_draw(aux, (double)(K - done) + 1.0);
        t_rem = t_rem - t_rem * (g1 / (g1 + g2));
        if (!(t_rem > 0)) break;
    }
    if (!TRACK_R) m.fix_removed(x);
    return total_fired;
}

// Uniformized interval with 32-bit candidate uniforms (SEM_ARITH_UNIFORMIZED32).  Candidate c of a particle-step takes
// word (c & 3) of Philox call c >> 2, so one call serves FOUR candidates (a candidate needs one uniform: is it a real
// event, and which).  A group of four candidates is straight-line code: once the batch has to stop (absorbed state,
// bound violated, K reached) the remaining candidates of the group are predicated off and their words are served again
// after the restart.  The bound anticipates growth: B = max(a0(x), a0(x + expected drift over h)) * (c0 + c1 /
// sqrt(expected events + 1)), and a batch covers h <= t_rem chosen so that the anticipated growth stays below GMAX (fast
// epidemics take several batches per interval; a batch that completes without violating its bound has covered exactly h).
// Every interval is uniformized (no direct-method tail).
// The loop is re-entrant (Unif32State) so that two warps can share one particle's interval: with `handoff` the call
// serves the first half of the current batch's candidates and returns false; a second call with the same state
// finishes the interval (pf_persistent's scheduler balancing).
#ifndef SEM_U32_C0
#define SEM_U32_C0 1.0
#define SEM_U32_C1 2.0
#define SEM_U32_GMAX 1.25          /* largest anticipated growth of the total propensity within one batch */
#endif
struct Unif32State { double t_rem, h, s; uint32_t cand, first, last, aux_k; int in_batch; };   // s = 2^32 / B (B = bound of the batch)

__device__ __forceinline__ void unif32_begin(Unif32State &st, double max_time) {
    st.t_rem = max_time; st.h = max_time; st.s = 0.0; st.cand = 0; st.first = 0; st.last = 0; st.aux_k = 0; st.in_batch = 0;
}

// First-batch / next-batch setup of the uniformized loop: bound B, covered time h, candidate count K ~ Poisson(B h).
// Returns false when the state is absorbed (nothing left to simulate).  K is known BEFORE the loop runs, which is what
// lets pf_persistent sort a CTA's particles by their exact amount of work (see there).
template <class Model>
__device__ __forceinline__ bool unif32_batch_setup(const Model &m, const double *x, Unif32State &st, PairSource<false> &aux,
                                                   double *r, double &a0, const double2 *tab) {
    if (!m.alive(x)) return false;
    a0 = ssa_total<Model, SEM_ARITH_FAST>(m, x, r);
    if (!(a0 > 0)) return false;
    double xp[Model::C], rp[Model::R];
    m.drift(x, r, st.t_rem, xp);
    const double a0p = ssa_total<Model, SEM_ARITH_FAST>(m, xp, rp);
    double amax = a0p > a0 ? a0p : a0;
    st.h = st.t_rem;
    const double cap = __dmul_rn(SEM_U32_GMAX, a0);
    if (amax > cap) {                                             // fast growth: a shorter batch, so that the (linearised)
        st.h = __dmul_rn(st.t_rem, __ddiv_rn(__dsub_rn(cap, a0), __dsub_rn(a0p, a0)));   // drift lifts a0 by GMAX at most
        amax = cap;
    }
    const double expect = __dmul_rn(a0, st.h);
    const double B = __dmul_rn(amax, __dadd_rn(SEM_U32_C0, __ddiv_rn(SEM_U32_C1, sqrt(__dadd_rn(expect, 1.0)))));
    double mu = __dmul_rn(B, st.h);                               // mean candidate count; rounded up to the tables' grid
    uint32_t K;
    if (mu <= kKTabMaxMean) {
        int id;
        mu = ktab_round_up(mu, id);
        K = ktab_draw(aux, id);
    } else {
        const double Kd = poisson_draw_u(aux, mu, tab);
        K = Kd < 2.0e9 ? (uint32_t)Kd : 2000000000u;
    }
    st.s = __ddiv_rn(__dmul_rn(4294967296.0, st.h), mu);          // 2^32 / B for the bound B = mu / h actually used
    st.first = st.cand; st.last = st.cand + K; st.in_batch = 1;
    return true;
}

// Every reaction of these models lowers sum_g (3 S_g + 2 E_g + I_g) by exactly one (infection: S-1 and E+1 or I+1;
// E -> I; recovery: I-1; the SIR models have no E), so the number of fired events of a leg is the drop of that tally --
// the loop need not count (two instructions per candidate).  Exact: the counts are integers far below 2^53.
template <class Model>
__device__ __forceinline__ double event_tally(const double *x) {
    double t = 0.0;
    if constexpr (Model::C == 4) t = 3.0 * x[0] + 2.0 * x[1] + x[2];          // SEIR: S, E, I, R
    else {
#pragma unroll
        for (int g = 0; g < Model::G; g++) t += 2.0 * x[3 * g] + x[3 * g + 1]; // SIR (sub)groups: S, I, R per group
    }
    return t;
}

// One candidate of the thinned stream against the fixed-point thresholds T (see Model::thresholds): the 32-bit word w is
// a real event iff w < lo32(T[R-1]) -- probability round(2^32 a0 / B) / 2^32 -- namely reaction j = #{i < R-1 : w >=
// lo32(T[i])}; integer compares on the low words, no conversion of the word, no fp64 compare.  After a fired event the
// thresholds are recomputed and the batch stops unless 0 < n(a0) < 2^32, i.e. unless the high word of T[R-1] is still
// that of 2^52 (the bound holds; n = 2^32 exactly counts as a violation, which is conservative and therefore exact) and
// its low word is non-zero (not absorbed).
template <class Model, bool TRACK_R>
__device__ __forceinline__ void unif32_candidate(const Model &m, const typename Model::Scaled &sc, double *x, double *T,
                                                 const uint32_t w, const bool live, bool &stop) {
    int j = 0;
#pragma unroll
    for (int i = 0; i < Model::R - 1; i++) j += (w >= (uint32_t)__double2loint(T[i])) ? 1 : 0;
    const bool hit = live && w < (uint32_t)__double2loint(T[Model::R - 1]);
    m.template apply_if<TRACK_R>(x, hit, j);
    m.thresholds(sc, x, T);
    const bool ok = (uint32_t)__double2hiint(T[Model::R - 1]) == kMagic52Hi && __double2loint(T[Model::R - 1]) != 0;
    stop = stop || (hit && !ok);
}

// Serve the candidates [cand, last) of a batch, four per Philox call (candidate c = word c & 3 of call c >> 2), until the
// batch stops.  An aligned group is straight-line code (live = not stopped yet and inside the batch); only the first group
// after a restart or a hand-over takes the general path (words before cand were served already).  (The last, partial
// group of a batch used to take the general path too: the lanes of a sorted warp end within ~10 consecutive iterations,
// each of which then ran both paths -- 7-9 % of the loop.)  (Computing the
// next group's Philox words behind this group's candidates was measured: the lone-warp latency of a group stays 540
// cycles -- ptxas does not interleave the two chains -- and the filter gets 6 % slower; tools/micro/cand_loop.cu.)
template <class Model, bool TRACK_R>
__device__ __forceinline__ void unif32_serve(const Model &m, const typename Model::Scaled &sc, double *x, double *T,
                                             PairSource<false> &loc, uint32_t &cand, const uint32_t last, bool &stop) {
    while (cand < last && !stop) {
        loc.k = cand >> 2;
        const uint4 w = loc.raw();
        if ((cand & 3u) == 0u) {                                       // an aligned group: straight-line; the last group of a batch
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};              // has fewer than four candidates left (rem)
            const uint32_t rem = last - cand;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const bool live = !stop && (uint32_t)q < rem;
                cand += live ? 1u : 0u;
                unif32_candidate<Model, TRACK_R>(m, sc, x, T, words[q], live, stop);
            }
        } else {
            const uint32_t base = cand & ~3u;
#pragma unroll 1
            for (uint32_t q = 0; q < 4u; q++) {
                const uint32_t c = base + q;
                const bool live = !stop && c >= cand && c < last;
                const uint32_t wq = q == 0u ? w.x : q == 1u ? w.y : q == 2u ? w.z : w.w;
                unif32_candidate<Model, TRACK_R>(m, sc, x, T, wq, live, stop);
                cand += live ? 1u : 0u;
            }
        }
    }
}

template <class Model, bool TRACK_R>
__device__ __forceinline__ bool ssa_unif32_leg(const Model &m, double *x, Unif32State &st, long long &fired_total, const bool handoff,
                                               PairSource<false> &src, PairSource<false> &aux, const double2 *tab) {
    PairSource<false> loc = src;
    aux.k = st.aux_k;
    bool finished = true;
    const double tally0 = event_tally<Model>(x);
    for (;;) {
        if (!st.in_batch) {                                            // (else: resumed -- second leg, or set up by the caller)
            double r[Model::R], a0;
            if (!unif32_batch_setup(m, x, st, aux, r, a0, tab)) break;
        }
        typename Model::Scaled sc;
        m.scale(sc, st.s);                                             // candidate words against 2^32 r / B
        double T[Model::R];
        m.thresholds(sc, x, T);
        // a first leg serves the first half of the batch's candidates only: the loop is the same, its bound differs
        const uint32_t last = handoff ? st.first + ((st.last - st.first) >> 1) : st.last;
        uint32_t cand = st.cand;
        bool stop = false;
        unif32_serve<Model, TRACK_R>(m, sc, x, T, loc, cand, last, stop);
        st.cand = cand;
        if (!stop && cand < st.last) { finished = false; break; }      // hand-over point reached (first leg only)
        st.in_batch = 0;
        const bool violated = stop && (uint32_t)__double2hiint(T[Model::R - 1]) != kMagic52Hi;
        if (!violated) {                                               // the batch covered its h exactly
            if (stop || !(st.h < st.t_rem)) break;                     // absorbed, or h was the rest of the interval
            st.t_rem = __dsub_rn(st.t_rem, st.h);
            continue;
        }
        // the bound held up to candidate `done`; its time is the done-th order statistic of K uniforms on [0, h]
        const uint32_t done = cand - st.first, K = st.last - st.first;
        const double g1 = gamma_draw(aux, (double)done), g2 = gamma_draw(aux, (double)(K - done) + 1.0);
        st.t_rem = __dsub_rn(st.t_rem, __dmul_rn(st.h, __ddiv_rn(g1, __dadd_rn(g1, g2))));
        if (!(st.t_rem > 0)) break;
    }
    st.aux_k = aux.k;
    src.k = loc.k;
    fired_total += (long long)(tally0 - event_tally<Model>(x));
    if (!TRACK_R) m.fix_removed(x);                                      // (also on a hand-over: the next leg's setup sums the state)
    return finished;
}

template <class Model, bool TRACK_R>
__device__ __forceinline__ long long ssa_run_unif32(const Model &m, double *x, double max_time, PairSource<false> &src,
                                                    PairSource<false> &aux, const double2 *tab) {
    Unif32State st;
    unif32_begin(st, max_time);
    long long fired = 0;
    ssa_unif32_leg<Model, TRACK_R>(m, x, st, fired, false, src, aux, tab);
    return fired;
}

template <class Model, int ARITH, bool REPLAY, bool TRACK_R, class Rec>
__device__ __forceinline__ long long ssa_run(const Model &m, double *x, double max_time, PairSource<REPLAY> &src,
                                             const double2 *tab, Rec rec) {
    if constexpr (ARITH != SEM_ARITH_REFERENCE) {            // a state holding a negative count fires no events (the
        bool neg = false;                                    // reference raises ValueError, gillespie_algo.py:63; the blocked
#pragma unroll                                               // loops would run time backwards under a negative propensity)
        for (int c = 0; c < Model::C; c++) neg = neg || x[c] < 0.0;
        if (neg) return 0;
    }
    if constexpr (ARITH == SEM_ARITH_FAST && !REPLAY) {
        if constexpr (SpecBlock<Model>::bits52 > 0) return ssa_run_spec<Model, SpecBlock<Model>::bits52, false, TRACK_R>(m, x, max_time, src, tab, rec);
        else return ssa_run_fast<Model, TRACK_R>(m, x, max_time, src, tab, rec);
    }
    else if constexpr (ARITH == SEM_ARITH_FAST32 && !REPLAY) return ssa_run_spec<Model, SpecBlock<Model>::bits32, true, TRACK_R>(m, x, max_time, src, tab, rec);
    else if constexpr (ARITH == SEM_ARITH_UNIFORMIZED && !REPLAY) {
        PairSource<false> aux; aux.init(*src.key, src.c1, src.c2, (src.c3 & 0xFFFFFFu) | (DOM_AUX << 24));
        return ssa_run_unif<Model, TRACK_R>(m, x, max_time, src, aux, tab);
    }
    else if constexpr (ARITH == SEM_ARITH_UNIFORMIZED32 && !REPLAY) {
        PairSource<false> aux; aux.init(*src.key, src.c1, src.c2, (src.c3 & 0xFFFFFFu) | (DOM_AUX << 24));
        return ssa_run_unif32<Model, TRACK_R>(m, x, max_time, src, aux, tab);
    }
    else return ssa_run_ref<Model, REPLAY>(m, x, max_time, src, rec);
}


// ------------------------------------------------------------------------------------------ block primitives
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_xor_d(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmax(v, shfl_xor_d(v, d));
    return v;
}
__device__ __forceinline__ double warp_incl_scan_d(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double o = shfl_up_d(v, d); if (lane >= d) v += o; }
    return v;
}

}  // namespace sem
