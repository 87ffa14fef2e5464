"""Posterior summaries of the reference's helpers.py, without arviz.

Same names and semantics: mean_credible_interval (helpers.py:5-13, HDI of a 1-D sample), gelman_rubin_test
(helpers.py:15-43), running_mean (helpers.py:46-48), posterior_mse (helpers.py:51-54); plus effective_sample_size and
acceptance_rate, which the reference's result-review scripts take from arviz / compute inline
(tests/test_pmcmc_seir.py:257-282).  O(chain length) host post-processing in numpy -- not a device path.
"""
import numpy as np

__all__ = ["hdi", "mean_credible_interval", "gelman_rubin_test", "running_mean", "posterior_mse", "effective_sample_size",
           "acceptance_rate"]


def hdi(a, hdi_prob=0.95):
    """Highest-density interval of a 1-D sample: the shortest window holding ceil(hdi_prob*n) sorted points (what
    arviz.hdi computes for unimodal samples)."""
    x = np.sort(np.asarray(a, dtype=float).ravel())
    n = x.size
    k = int(np.floor(hdi_prob * n))
    if k >= n:
        return x[0], x[-1]
    widths = x[k:] - x[:n - k]
    i = int(np.argmin(widths))
    return x[i], x[i + k]


def mean_credible_interval(data, alpha=0.95):
    """helpers.py:5-13: (mean, lo, hi) with the 95% HDI (the reference ignores `alpha` and always uses 0.95)."""
    a = np.array(data)
    m = np.mean(a, axis=1) if a.ndim == 2 else np.mean(a)
    lo, hi = hdi(a, 0.95)
    return m, lo, hi


def gelman_rubin_test(chains):
    """helpers.py:15-43: potential scale reduction sqrt(V/W) per parameter for M chains of shape (N, P)."""
    chains = [np.asarray(c, dtype=float) for c in chains]
    M = len(chains)
    N = chains[0].shape[0]
    means = np.array([c.mean(axis=0) for c in chains])
    variances = np.array([c.var(axis=0, ddof=1) for c in chains])
    W = variances.mean(axis=0)
    B = N / (M - 1) * np.sum((means - means.mean(axis=0)) ** 2, axis=0)
    V = (N - 1) / N * W + (M + 1) / (M * N) * B
    return np.sqrt(V / W)


def running_mean(x, N):
    """helpers.py:46-48."""
    cumsum = np.cumsum(np.insert(x, 0, 0))
    return (cumsum[N:] - cumsum[:-N]) / float(N)


def posterior_mse(true_prm, chain):
    """helpers.py:51-54."""
    return np.mean((chain - true_prm) ** 2)


def effective_sample_size(x):
    """ESS of a 1-D chain from the initial positive sequence of autocorrelations (Geyer)."""
    x = np.asarray(x, dtype=float)
    n = x.size
    xc = x - x.mean()
    var = xc.var()
    if var == 0 or n < 4:
        return float(n)
    f = np.fft.rfft(xc, 2 * n)
    acf = np.fft.irfft(f * np.conj(f))[:n] / (var * n)
    # Geyer's initial positive sequence: Gamma_m = rho_{2m} + rho_{2m+1}, m = 0, 1, ... summed while positive;
    # integrated autocorrelation time tau = -1 + 2 sum Gamma_m  (rho_0 = 1)
    s = 0.0
    for k in range(0, n - 1, 2):
        pair = acf[k] + acf[k + 1]
        if pair < 0:
            break
        s += pair
    tau = max(2 * s - 1, 1e-12)
    return float(min(n / tau, n))


def acceptance_rate(thetas):
    """Fraction of MH iterations that moved (distinct consecutive rows of the chain)."""
    t = np.asarray(thetas)
    return float(np.mean(np.any(t[1:] != t[:-1], axis=1))) if len(t) > 1 else 0.0
