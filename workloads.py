"""Synthetic inputs for bench.py, __graft_entry__.smoke() and the full-size tests (numpy only, deterministic).

The reference fabricates its "true" epidemics by integrating the deterministic ODE and keeping one row per
integer day (pmcmc.py:16-113, sir_simulate_discrete & co.), then observes them with binomial thinning
(tests/test_particles.py:31-38) or relative normal noise (tests/test_noise.py:23-32).  This file restates that
recipe with a fixed-step RK4 so the bench does not need scipy.odeint or the reference on the GPU box.
"""
import numpy as np


def _rk4(f, y0, T, substeps=40):
    y = np.array(y0, dtype=float)
    rows = [y.copy()]
    h = 1.0 / substeps
    for _ in range(T - 1):
        for _ in range(substeps):
            k1 = f(y); k2 = f(y + h / 2 * k1); k3 = f(y + h / 2 * k2); k4 = f(y + h * k3)
            y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        rows.append(y.copy())
    return np.array(rows)


def sir_truth(y0, T, beta, gamma):
    N = float(sum(y0))
    return _rk4(lambda v: np.array([-beta * v[0] * v[1] / N, beta * v[0] * v[1] / N - gamma * v[1], gamma * v[1]]), y0, T)


def seir_truth(y0, T, beta, alpha, gamma):
    N = float(sum(y0))
    return _rk4(lambda v: np.array([-beta * v[0] * v[2] / N, beta * v[0] * v[2] / N - alpha * v[1],
                                    alpha * v[1] - gamma * v[2], gamma * v[2]]), y0, T)


def observe_binomial(truth, p, seed=0):
    return np.random.RandomState(seed).binomial(np.maximum(truth, 0).astype(int), p).astype(float)


def observe_normal(truth, ratio, seed=0):
    rng = np.random.RandomState(seed)
    return (truth + rng.normal(0, ratio * np.maximum(truth, 0) + 1e-12)).astype(int).astype(float)


# BASELINE.json metric workload: SIR, population 1e4, 1e5 particles, 100 observation intervals (config 4(ii),
# y0=(9980,20,0), beta=.4, gamma=.2, binomial p=.1 under-reporting)
HEADLINE = dict(name="pf_sir_pop1e4_N1e5_T101_binomial_p0.1", model="SIR", y0=(9980, 20, 0), theta=(.4, .2), T=101,
                n_particles=100_000, n_population=10_000, mu=20, probs=.1, observations=False)


def headline_Y():
    w = HEADLINE
    return observe_binomial(sir_truth(w["y0"], w["T"], *w["theta"]), w["probs"], seed=0)
