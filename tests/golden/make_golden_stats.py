"""Statistical goldens from the UNMODIFIED reference (slow: ~15-25 min on one core).

    python tests/golden/make_golden_stats.py

1. stat_logz_sir.npz   : 300 independent runs of the reference particle_filter (pmcmc.py:123) at a fixed theta
                         -> the sampling distribution of its likelihood estimate zetas[-1] (Pattern A of SURVEY 4).
2. stat_pmcmc_sir.npz  : one reference particle_mcmc chain (pmcmc.py:251) -> posterior draws of (beta, gamma).
The GPU tests compare the B200 engine's estimates with these within Monte-Carlo error (north_star, third check).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
import types  # noqa: E402


def main():
    pm = rh.load_reference()["pmcmc"]
    pm.tqdm = lambda *a, **k: types.SimpleNamespace(update=lambda *_: None, set_description_str=lambda *_: None)
    t = np.linspace(0, 7, num=200)
    truth = pm.sir_simulate_discrete((480, 20, 0), t, 2.0, 1.0).iloc[:, 1:].to_numpy(dtype=float)
    Y = np.random.RandomState(5).binomial(truth.astype(int), .1).astype(float)
    theta = np.array([2.0, 1.0])
    N, npop, mu = 50, 500, 20

    t0 = time.time()
    np.random.seed(1234)
    z = []
    for i in range(300):
        zetas, _, _ = pm.particle_filter(Y, pm.ModelType.SIR, theta, False, .1, N, npop, mu, 1)
        z.append(np.nan if zetas is None else zetas[-1])
    print(f"logZ runs: {time.time() - t0:.0f}s, mean Z {np.nanmean(z):.4e}, collapsed {np.isnan(z).sum()}")
    np.savez_compressed(os.path.join(HERE, "stat_logz_sir.npz"), Y=Y, theta=theta, n_particles=N, n_population=npop, mu=mu,
                        probs=.1, zetas_last=np.array(z))

    t0 = time.time()
    np.random.seed(4321)
    n_chains = 1500
    thetas, lik, _ = pm.particle_mcmc(Y, pm.ModelType.SIR, [2.0, 1.0], .02, n_chains=n_chains, probs=.1, n_particles=N,
                                      n_population=npop, mu=mu, jobs=1)
    print(f"pmcmc chain: {time.time() - t0:.0f}s, mean theta {thetas[300:].mean(0)}, acc {len(np.unique(thetas[:, 0])) / n_chains:.2f}")
    np.savez_compressed(os.path.join(HERE, "stat_pmcmc_sir.npz"), Y=Y, parameters=np.array([2.0, 1.0]), h=.02, n_chains=n_chains,
                        n_particles=N, n_population=npop, mu=mu, probs=.1, thetas=thetas, likelihoods=lik)


if __name__ == "__main__":
    main()
