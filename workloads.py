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


def subgroups_truth(y0, T, betas, gamma):
    """G-group SIR ODE in the SSA's convention (gillespie_algo.py:182-183): beta[a, b] = infector group a ->
    susceptible group b.  y0 (G,3) rows [S,I,R]; returns (T, 3G) columns [S0,I0,R0,S1,...]."""
    y0 = np.asarray(y0, dtype=float); betas = np.asarray(betas, dtype=float)
    G, N = y0.shape[0], float(y0.sum())

    def f(v):
        x = v.reshape(G, 3)
        inf = (betas * x[:, 1][:, None]).sum(0) * x[:, 0] / N          # new infections per susceptible group b
        return np.stack([-inf, inf - gamma * x[:, 1], gamma * x[:, 1]], 1).reshape(-1)
    return _rk4(f, y0.reshape(-1), T)


# The other BASELINE.json configurations at sizes one GPU runs in a bench (extra measurements, bench.py --workload configs)
CONFIGS = {
    # config 3: SEIR, binomial p=.1, time-rescaled x10: beta=.4, alpha=.1, gamma=.1, 101 days, pop 1e4
    "seir_pop1e4_N1e5_T101": dict(model=1, G=1, y0=(9980, 0, 20, 0), theta=(.4, .1, .1), T=101, n_particles=100_000,
                                  n_population=[10_000], mu=[20], probs=.1, observations=False),
    # config 4(i): SIR with noisy counts (normal observation model, ratio .1)
    "sir_noisy_pop1e4_N1e5_T101": dict(model=0, G=1, y0=(9980, 20, 0), theta=(.4, .2), T=101, n_particles=100_000,
                                       n_population=[10_000], mu=[20], probs=.1, observations=True),
    # config 5 shape (tests/test_pmcmc_sir_subgrps2.py:13-15) at pop 1e4 instead of 1e6: 2 groups, group-summed observations
    "sub2_pop1e4_N1e5_T15": dict(model=3, G=2, y0=((3940, 60, 0), (5920, 80, 0)), theta=(5, 2, 1, 3, .5), T=15,
                                 n_particles=100_000, n_population=[4000, 6000], mu=[60, 80], probs=.1, observations=False),
}


def config_Y(name):
    w = CONFIGS[name]
    if w["model"] == 1:
        truth = seir_truth(w["y0"], w["T"], *w["theta"])
    elif w["model"] == 0:
        truth = sir_truth(w["y0"], w["T"], *w["theta"])
    else:
        G = w["G"]
        full = subgroups_truth(w["y0"], w["T"], np.array(w["theta"][:G * G]).reshape(G, G), w["theta"][-1])
        truth = full.reshape(w["T"], G, 3).sum(1) if w["model"] == 3 else full       # SUBGROUPS2 observes the group sums
    if w["observations"]:
        return observe_normal(truth, w["probs"], seed=0)
    return observe_binomial(truth, w["probs"], seed=0)
