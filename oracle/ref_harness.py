"""Harness that runs the UNMODIFIED reference (/root/reference) to manufacture golden vectors.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this file.  It is used by
``tests/golden/make_golden.py`` in the build container (where /root/reference exists); it cannot
run on the GPU box.  The reference is never copied: it is imported from where it lies.

How determinism is obtained:
  * ``np.random.seed(s)`` + ``jobs=1`` makes the reference consume numpy's global legacy MT19937
    stream strictly sequentially (joblib with n_jobs=1 runs in-process).
  * Every draw on the hot path after initialisation is a ``random_sample`` double:
      - ``np.random.exponential(scale)``   = ``-log(1.0 - u) * scale``          (1 double)
      - ``np.random.choice(R, p=p)``       = ``searchsorted(cumsum(p)/cumsum(p)[-1], u, 'right')`` (1 double)
      - ``np.random.choice(range(N), N, p=w)`` = N doubles, same searchsorted rule
    (gillespie_algo.py:62-63,133-134,208-212; pmcmc.py:188-190).
  * Draws whose word consumption is variable (``poisson``, ``randint``) happen *before* the part we
    replay, so we snapshot the MT19937 state right after them and take their results as inputs.
So a golden vector is (inputs, MT19937 state at the start of the replayed part, outputs): the uniform
stream is regenerated at test time with ``np.random.RandomState().set_state(...)``.
"""
import sys
import types

import numpy as np

REFERENCE_PATH = "/root/reference"

_ref = {}


def load_reference():
    """Import gillespie_algo / pmcmc / abc_algo from /root/reference.

    pmcmc.py:5 imports matplotlib (absent here, never used on the hot path) -> empty stub modules.
    """
    if _ref:
        return _ref
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    if REFERENCE_PATH not in sys.path:
        sys.path.insert(0, REFERENCE_PATH)
    import gillespie_algo  # noqa
    import pmcmc  # noqa
    import abc_algo  # noqa

    assert gillespie_algo.__file__.startswith(REFERENCE_PATH)
    assert pmcmc.__file__.startswith(REFERENCE_PATH)
    _ref.update(gillespie_algo=gillespie_algo, pmcmc=pmcmc, abc_algo=abc_algo)
    return _ref


def mt_state():
    """Current global MT19937 state as (key[624] uint32, pos)."""
    st = np.random.get_state()
    assert st[0] == "MT19937"
    return np.array(st[1], dtype=np.uint32), int(st[2])


class _Patch:
    """Context manager: temporarily replace attributes, restore on exit."""

    def __init__(self):
        self._saved = []

    def set(self, obj, name, value):
        self._saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        for obj, name, val in reversed(self._saved):
            setattr(obj, name, val)
        return False


def run_ssa(model, population, theta, max_time, seed):
    """Run one reference SSA with full trajectory.  model in {'sir','seir','subgroups'}.

    For 'subgroups' theta = (betas[G,G], gamma).  Returns dict with the state snapshot and outputs.
    """
    g = load_reference()["gillespie_algo"]
    np.random.seed(seed)
    key, pos = mt_state()
    if model == "sir":
        out = g.sir_simulate(list(population), np.array(theta, dtype=float), max_time, False)
        cols = ["s", "i", "r"]
    elif model == "seir":
        out = g.seir_simulate(list(population), np.array(theta, dtype=float), max_time, False)
        cols = ["s", "e", "i", "r"]
    else:
        betas, gamma = theta
        pop = np.array(population)
        out = g.sir_subgroups_simulate(pop, np.array(betas, dtype=float), float(gamma), max_time, False)
        cols = [f"{c}_{grp}" for grp in range(pop.shape[0]) for c in ("s", "i", "r")]
    states = np.array([out[c] for c in cols], dtype=np.float64).T  # (events+1, C)
    times = np.array(out["time"], dtype=np.float64)
    return dict(mt_key=key, mt_pos=pos, times=times, states=states)


def run_particle_filter(Y, model_name, theta, observations, probs, n_particles, n_population, mu, seed):
    """Run the reference particle_filter (pmcmc.py:123-233) with jobs=1 under a fixed seed.

    Captures the MT19937 state right after the Poisson initialisation (pmcmc.py:156-169), i.e. at
    the start of the time loop (pmcmc.py:177).  Returns inputs+outputs; collapsed filters return
    zetas=None (pmcmc.py:191-192).
    """
    ref = load_reference()
    pm = ref["pmcmc"]
    mt = getattr(pm.ModelType, model_name)
    snap = {}
    real_poisson = np.random.poisson

    def poisson_spy(*a, **k):
        r = real_poisson(*a, **k)
        snap["key"], snap["pos"] = mt_state()
        return r

    np.random.seed(seed)
    with _Patch() as p:
        p.set(np.random, "poisson", poisson_spy)
        zetas, hidden, anc = pm.particle_filter(
            np.array(Y), mt, theta, observations, probs, n_particles, n_population, mu, 1
        )
    out = dict(mt_key=snap["key"], mt_pos=snap["pos"], collapsed=zetas is None)
    if zetas is not None:
        out.update(zetas=zetas, hidden_process=hidden, ancestry_matrix=anc)
    return out


def run_path_sampler(hidden_process, ancestry_matrix, seed):
    """Reference particle_path_sampler (pmcmc.py:236-248); records the randint pick."""
    pm = load_reference()["pmcmc"]
    picked = {}
    real_randint = np.random.randint

    def randint_spy(*a, **k):
        r = real_randint(*a, **k)
        picked["v"] = int(r)
        return r

    np.random.seed(seed)
    with _Patch() as p:
        p.set(np.random, "randint", randint_spy)
        traj = pm.particle_path_sampler(hidden_process, ancestry_matrix)
    return dict(chosen=picked["v"], trajectory=traj)


def run_abc(observed_data, no_of_samples, threshold, priors, seed):
    """Reference abc_algo (abc_algo.py:17-109) with per-trial capture.

    For every trial records beta, gamma, the Poisson-perturbed start (abc_algo.py:39-40), the MT19937
    state at entry to sir_simulate (abc_algo.py:41), the discretised [T,4] trajectory handed to the
    distance function (abc_algo.py:93-99) and the distance.
    """
    ref = load_reference()
    ab = ref["abc_algo"]
    trials = []
    real_sim = ab.sir_simulate
    real_dist = ab.distance_function

    def sim_spy(population, theta, max_time, last_values_only):
        key, pos = mt_state()
        trials.append(dict(n_start=np.array(population, dtype=np.int64), theta=np.array(theta, dtype=float),
                           mt_key=key, mt_pos=pos))
        return real_sim(population, theta, max_time, last_values_only)

    def dist_spy(I_1, I_2, R_1, R_2):
        d = real_dist(I_1, I_2, R_1, R_2)
        trials[-1]["I_sim"] = np.array(I_1, dtype=float)
        trials[-1]["R_sim"] = np.array(R_1, dtype=float)
        trials[-1]["distance"] = float(d)
        return d

    np.random.seed(seed)
    with _Patch() as p:
        p.set(ab, "sir_simulate", sim_spy)
        p.set(ab, "distance_function", dist_spy)
        # silence tqdm bars
        p.set(ab, "tqdm", lambda *a, **k: types.SimpleNamespace(update=lambda *_: None,
                                                                 set_description_str=lambda *_: None))
        post, trajs = ab.abc_algo(np.array(observed_data), no_of_samples, threshold, priors)
    return dict(trials=trials, posterior=post, trajectories=np.array(trajs))
