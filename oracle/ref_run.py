"""Run the UNMODIFIED Python reference from oracle/_ref/ (see oracle/make_ref.py) -- the CPU baseline of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing in the product path imports this file.
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return all(os.path.exists(os.path.join(REF, f)) for f in ("gillespie_algo.py", "pmcmc.py", "abc_algo.py"))


_mods = {}


def load():
    """Import pmcmc / gillespie_algo / abc_algo from oracle/_ref (matplotlib stubbed, pmcmc.py:5)."""
    if _mods:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref is empty: run `python oracle/make_ref.py` in the build container")
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    saved = {k: sys.modules.pop(k) for k in ("gillespie_algo", "pmcmc", "abc_algo") if k in sys.modules}
    sys.path.insert(0, REF)
    try:
        import gillespie_algo, pmcmc, abc_algo  # noqa: E401
        assert os.path.dirname(os.path.abspath(pmcmc.__file__)) == REF
        _mods.update(gillespie_algo=gillespie_algo, pmcmc=pmcmc, abc_algo=abc_algo)
    finally:
        sys.path.remove(REF)
        for k in ("gillespie_algo", "pmcmc", "abc_algo"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return _mods


def time_particle_filter(Y, theta, observations, probs, n_particles, n_population, mu, jobs=1, seed=0):
    """One call of the reference particle_filter (pmcmc.py:123) on host cores.  Returns dict(seconds, particle_steps, collapsed)."""
    pm = load()["pmcmc"]
    np.random.seed(seed)
    t0 = time.perf_counter()
    zetas, hidden, anc = pm.particle_filter(np.asarray(Y, dtype=float), pm.ModelType.SIR, np.asarray(theta, dtype=float), observations,
                                            probs, n_particles, n_population, mu, jobs)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, particle_steps=n_particles * (len(Y) - 1), collapsed=zetas is None,
                log_z=None if zetas is None else float(np.log(zetas[-1])) if zetas[-1] > 0 else float("-inf"))
