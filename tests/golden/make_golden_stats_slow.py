"""Statistical goldens from the UNMODIFIED reference in the SLOW-GROWTH regime (beta = .4, gamma = .2: growth 0.2 per
observation interval), i.e. where the engine's default `arith="auto"` resolves to the uniformized interval simulation
(engine.resolve_arith) -- the regime of the BASELINE headline workload, at sizes the Python reference can run.

    python tests/golden/make_golden_stats_slow.py [logz] [pmcmc]      (about an hour on 8 cores; worker processes)

1. stat_logz_sir_slow.npz  : 320 independent runs of the reference particle_filter (pmcmc.py:123) at a fixed theta
                             -> sampling distribution of zetas[-1]; every run in its own np.random.seed(stream).
2. stat_pmcmc_sir_slow.npz : reference particle_mcmc chains (pmcmc.py:251), one per worker process
                             -> posterior draws of (beta, gamma) for the mean / 95 % HDI comparison.
The reference is imported from /root/reference where it lies (oracle/ref_harness.py); only outputs are stored.
"""
import multiprocessing as mp
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

BETA, GAMMA, POP, MU, T, N, PROBS = .4, .2, 800, 20, 16, 64, .1
LOGZ_WORKERS, LOGZ_PER_WORKER = 8, 40
PMCMC_CHAINS, PMCMC_ITERS, PMCMC_H = 4, 2500, 6e-4


def _ref():
    from oracle import ref_harness as rh
    pm = rh.load_reference()["pmcmc"]
    pm.tqdm = lambda *a, **k: types.SimpleNamespace(update=lambda *_: None, set_description_str=lambda *_: None)
    return pm


def make_Y():
    pm = _ref()
    t = np.linspace(0, T - 1, num=200 * (T - 1))
    truth = pm.sir_simulate_discrete((POP - MU, MU, 0), t, BETA, GAMMA).iloc[:, 1:].to_numpy(dtype=float)
    assert truth.shape == (T, 3), truth.shape
    return np.random.RandomState(15).binomial(truth.astype(int), PROBS).astype(float)


def logz_worker(args):
    w, Y = args
    pm = _ref()
    theta = np.array([BETA, GAMMA])
    out = []
    for i in range(LOGZ_PER_WORKER):
        np.random.seed(100_000 + 1000 * w + i)
        zetas, _, _ = pm.particle_filter(Y, pm.ModelType.SIR, theta, False, PROBS, N, POP, MU, 1)
        out.append(np.nan if zetas is None else zetas[-1])
    return out


def pmcmc_worker(args):
    c, Y = args
    pm = _ref()
    np.random.seed(777 + c)
    thetas, lik, _ = pm.particle_mcmc(Y, pm.ModelType.SIR, [BETA, GAMMA], PMCMC_H, n_chains=PMCMC_ITERS, probs=PROBS,
                                      n_particles=N, n_population=POP, mu=MU, jobs=1)
    return thetas, lik


def main():
    what = sys.argv[1:] or ["logz", "pmcmc"]
    Y = make_Y()
    print("Y =", Y.tolist())
    if "logz" in what:
        t0 = time.time()
        with mp.Pool(LOGZ_WORKERS) as pool:
            z = np.concatenate(pool.map(logz_worker, [(w, Y) for w in range(LOGZ_WORKERS)]))
        print(f"logZ runs: {time.time() - t0:.0f}s, mean Z {np.nanmean(z):.4e}, sd log Z {np.nanstd(np.log(z)):.3f}, collapsed {np.isnan(z).sum()}")
        np.savez_compressed(os.path.join(HERE, "stat_logz_sir_slow.npz"), Y=Y, theta=np.array([BETA, GAMMA]), n_particles=N,
                            n_population=POP, mu=MU, probs=PROBS, zetas_last=z)
    if "pmcmc" in what:
        t0 = time.time()
        with mp.Pool(PMCMC_CHAINS) as pool:
            res = pool.map(pmcmc_worker, [(c, Y) for c in range(PMCMC_CHAINS)])
        thetas = np.stack([r[0] for r in res])
        lik = np.stack([r[1] for r in res])
        print(f"pmcmc chains: {time.time() - t0:.0f}s, means {thetas[:, 300:].mean(1).tolist()}, "
              f"acc {[len(np.unique(th[:, 0])) / len(th) for th in thetas]}")
        np.savez_compressed(os.path.join(HERE, "stat_pmcmc_sir_slow.npz"), Y=Y, parameters=np.array([BETA, GAMMA]), h=PMCMC_H,
                            n_chains=PMCMC_ITERS, n_particles=N, n_population=POP, mu=MU, probs=PROBS, thetas=thetas,
                            likelihoods=lik)


if __name__ == "__main__":
    main()
