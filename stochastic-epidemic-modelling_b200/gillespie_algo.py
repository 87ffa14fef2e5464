"""Drop-in for the reference's gillespie_algo.py: exact Gillespie SSA on the GPU (sem_ssa_simulate).

Signatures follow gillespie_algo.py:10-12, 78-80, 148-151.  Extra keyword-only arguments keep their
behaviour-preserving defaults.  A single simulation occupies one GPU thread; use the *_batch helpers to
run many at once (that is where the device pays off).
"""
import numpy as np

from . import engine

__all__ = ["sir_simulate", "seir_simulate", "sir_subgroups_simulate", "simulate_batch", "predict_forward"]


def _max_rows(model, x0):
    """Upper bound on events+1: every susceptible can be infected once, every infected recover once."""
    x = np.asarray(x0, dtype=np.int64)
    if model == 0:
        return int(2 * x[0] + x[1]) + 2
    if model == 1:
        return int(3 * x[0] + 2 * x[1] + x[2]) + 2
    x = x.reshape(-1, 3)
    return int((2 * x[:, 0] + x[:, 1]).sum()) + 2


def _check_propensities(model, G, x0, theta):
    """The reference raises ValueError("probabilities are not non-negative") from np.random.choice when a propensity is
    negative (gillespie_algo.py:63,134,209-212) -- a negative count next to a live infection, or a negative rate
    constant.  Same error here, before anything is launched (the kernels themselves fire no events from such a state,
    DESIGN section 2, D9)."""
    x = np.asarray(x0, dtype=np.float64).reshape(-1)
    th = np.asarray(theta, dtype=np.float64).reshape(-1)
    n_tot = x.sum()
    if not n_tot > 0:
        return
    if model == 0:
        alive, rates = x[1] > 0, [th[0] * x[0] * x[1] / n_tot, th[1] * x[1]]
    elif model == 1:
        alive, rates = (x[1] > 0) or (x[2] > 0), [th[0] * x[0] * x[2] / n_tot, th[1] * x[1], th[2] * x[2]]
    else:
        xs = x.reshape(G, 3)
        alive = xs[:, 1].sum() > 0
        rates = [th[a * G + b] * xs[b, 0] * xs[a, 1] / n_tot for a in range(G) for b in range(G)] + [th[G * G] * xs[a, 1] for a in range(G)]
    if alive and min(rates) < 0:
        raise ValueError("probabilities are not non-negative")


def _run_one(model, G, x0, theta, max_time, last_values_only, names, seed, arith, replay_u):
    seed = engine.new_seed() if seed is None else seed
    _check_propensities(model, G, x0, theta)
    replay = None
    if replay_u is not None:
        replay_u = np.asarray(replay_u, dtype=np.float64)
        replay = dict(u=replay_u, off=np.array([0, replay_u.size], dtype=np.int64))
        arith = "reference"
    cap = 0 if last_values_only else _max_rows(model, x0)
    out = engine.simulate(model, np.asarray(x0, dtype=np.int64).reshape(-1), np.asarray(theta, dtype=np.float64).reshape(-1),
                          max_time, G=G, arith=arith, seed=seed, record_capacity=cap, replay=replay, n_sims=1)
    rows = int(out["n_rows"].cpu()[0])
    if rows < 0:
        raise RuntimeError("replay buffer exhausted")
    if last_values_only:
        return [int(v) for v in out["x"].cpu().numpy()[0]]
    states = out["states"][0, :rows].cpu().numpy()
    times = out["times"][0, :rows].cpu().numpy()
    cond = {name: states[:, c].tolist() for c, name in enumerate(names)}
    cond["time"] = times.tolist()
    return cond


def sir_simulate(population, theta_proposal, max_time, last_values_only, *, seed=None, arith="fast32", replay_u=None):
    """gillespie_algo.py:10-75.  population [S,I,R]; theta_proposal array [beta,gamma].
    Returns (S,I,R) at max_time, or the dict of event-by-event lists {"s","i","r","time"}."""
    r = _run_one(0, 1, population, theta_proposal, max_time, last_values_only, ["s", "i", "r"], seed, arith, replay_u)
    return tuple(r) if last_values_only else r


def seir_simulate(population, theta_proposal, max_time, last_values_only, *, seed=None, arith="fast32", replay_u=None):
    """gillespie_algo.py:78-146.  population [S,E,I,R]; theta_proposal [beta,alpha,gamma] (:92)."""
    r = _run_one(1, 1, population, theta_proposal, max_time, last_values_only, ["s", "e", "i", "r"], seed, arith, replay_u)
    return tuple(r) if last_values_only else r


def sir_subgroups_simulate(population, betas_proposal, gamma_proposal, max_time, last_values_only, *, seed=None,
                           arith="fast32", replay_u=None):
    """gillespie_algo.py:148-233.  population (G,3); betas (G,G) with betas[a,b] = infector group a ->
    susceptible group b (:182-183).  Returns list of G [S,I,R] lists (:224-231) or the dict of lists keyed
    "s_g","i_g","r_g","time" (:169-174)."""
    pop = np.asarray(population)
    G = pop.shape[0]
    theta = np.concatenate([np.asarray(betas_proposal, dtype=np.float64).reshape(-1), [float(gamma_proposal)]])
    names = [f"{c}_{g}" for g in range(G) for c in ("s", "i", "r")]
    r = _run_one(2, G, pop.reshape(-1), theta, max_time, last_values_only, names, seed, arith, replay_u)
    if last_values_only:
        return [r[3 * g:3 * g + 3] for g in range(G)]
    return r


def simulate_batch(model, populations, thetas, max_time, *, n_groups=1, seed=None, arith="fast32", n_sims=None):
    """Many independent SSA runs at once: populations (n,C) or (C,), thetas (n,P) or (P,) -> (n,C) int32 CUDA
    tensor of final states.  model: 'sir' | 'seir' | 'sir_subgroups'."""
    seed = engine.new_seed() if seed is None else seed
    m = engine.MODEL_NAMES[model] if isinstance(model, str) else int(model)
    out = engine.simulate(m, populations, thetas, max_time, G=n_groups, arith=arith, seed=seed, n_sims=n_sims)
    return out["x"]


def predict_forward(model, thetas, last_states, horizon, *, n_groups=1, seed=None, arith="fast32"):
    """Forward-prediction fan-out of tests/pred_tmps.py:55-73: from every posterior sample's last
    state simulate `horizon` more days with its own theta; returns (n, horizon, C) int32 CUDA tensor whose row d is
    the state at integer time d+1 (the state the reference extracts as "last event with floor(time) == d",
    forward-filled on days without events)."""
    seed = engine.new_seed() if seed is None else seed
    m = engine.MODEL_NAMES[model] if isinstance(model, str) else int(model)
    out = engine.simulate(m, last_states, thetas, float(horizon), G=n_groups, arith=arith, seed=seed, daily_rows=int(horizon))
    return out["states"]
