"""Drop-in for the hot-path functions of the reference's pmcmc.py (pmcmc.py:116-408).

particle_filter / particle_path_sampler / particle_mcmc keep the reference's positional and keyword
signatures; extra arguments are keyword-only with behaviour-preserving defaults.  The whole filter (X_0,
weights, likelihood, resampling, SSA propagation, all T steps) runs on the GPU through one sem_pf_run call;
the Metropolis-Hastings loop stays on the host like the reference's, but in the log domain.

Deliberate, documented differences from the reference:
  D2  likelihoods are accumulated as log-sums on the device; `zetas` is exp() of that (same numbers where the
      reference does not underflow, finite log-likelihood where it does).
  D3  resampler='systematic' (one uniform per step) is the default; resampler='multinomial' is the reference's
      np.random.choice rule.  Both leave the likelihood estimator unbiased.
  D4  the MH ratio of pmcmc.py:376-391 reduces algebraically to Z'/Z_prev (symmetric proposal, the "prior" terms
      cancel); it is evaluated as log u < logZ' - logZ_prev.  The reference's min(1, nan) -> accept quirk on a
      0/0 underflow is not reproduced.
  D8  particle_path_sampler keeps the reference's off-by-one ancestry indexing by default (exact_genealogy=False).
  `jobs` is accepted and ignored (no process pool).
The ODE data synthesiser (pmcmc.py:16-113) is a batched GPU utility here (RK4): *_simulate_discrete.
"""
from enum import Enum

import numpy as np
import torch

from . import engine
from .gillespie_algo import *  # noqa: F401,F403  (pmcmc.py:13 re-exports the simulators)
from .gillespie_algo import __all__ as _g_all

__all__ = ["ModelType", "particle_filter", "particle_path_sampler", "particle_mcmc", "pf_loglik", "sir_simulate_discrete",
           "seir_simulate_discrete", "sir_subgroups_simulate_discrete", "simulate_discrete_batch"] + list(_g_all)


class ModelType(Enum):          # pmcmc.py:116-120
    SIR = "sir"
    SEIR = "seir"
    SIR_SUBGROUPS = "sir_subgroups"
    SIR_SUBGROUPS2 = "sir_subgroups2"


def _model_id(type_model):
    """pmcmc.py:134-141: SIR, SEIR, SIR_SUBGROUPS2 are recognised; anything else takes the subgroup branch."""
    v = getattr(type_model, "value", type_model)
    return {"sir": 0, "seir": 1, "sir_subgroups2": 3}.get(v, 2)


def _flatten_theta(model, theta_proposal):
    if model >= 2:      # (beta[G,G], gamma)  pmcmc.py:214-215
        betas, gamma = theta_proposal
        return np.concatenate([np.asarray(betas, dtype=np.float64).reshape(-1), [float(gamma)]])
    return np.asarray(theta_proposal, dtype=np.float64).reshape(-1)


def simulate_discrete_batch(type_model, y0, theta, t, *, substeps=8):
    """The reference's ODE data synthesiser for a BATCH of parameter sets: mean-field ODE of the model
    integrated over the grid t (RK4 on the GPU, `substeps` per grid interval; odeint in the reference), one row per integer
    day (pmcmc.py:54-113).  y0 (B,C) or (C,), theta (B,P) or (P,) with the subgroup betas row-major as pmcmc.py:47 uses
    them, then gamma.  Returns float64 numpy (B, days+1, C)."""
    model = _model_id(type_model)
    y0 = np.asarray(y0, dtype=np.float64)
    G = (y0.shape[-1] // 3) if model >= 2 else 1
    return engine.ode_daily(model, y0, np.asarray(theta, dtype=np.float64), t, G=G, substeps=substeps).cpu().numpy()


def _discrete_frame(rows, names):
    import pandas as pd
    return pd.DataFrame({"time": np.arange(rows.shape[0]), **{n: rows[:, c] for c, n in enumerate(names)}})


def sir_simulate_discrete(y0, t, beta, gamma):
    """pmcmc.py:54-74: DataFrame(time, susceptible, infected, removed), one row per integer day."""
    rows = simulate_discrete_batch(ModelType.SIR, np.asarray(y0, dtype=float), [beta, gamma], t)[0]
    return _discrete_frame(rows, ["susceptible", "infected", "removed"])


def seir_simulate_discrete(y0, t, beta, alpha, gamma):
    """pmcmc.py:77-98: DataFrame(time, susceptible, exposed, infected, removed)."""
    rows = simulate_discrete_batch(ModelType.SEIR, np.asarray(y0, dtype=float), [beta, alpha, gamma], t)[0]
    return _discrete_frame(rows, ["susceptible", "exposed", "infected", "removed"])


def sir_subgroups_simulate_discrete(y0, t, beta, gamma):
    """pmcmc.py:101-113: y0 (G,3), beta (G,G); DataFrame(susceptible0, infected0, removed0, ..., time)."""
    import pandas as pd
    y0 = np.asarray(y0, dtype=float)
    G = y0.shape[0]
    rows = simulate_discrete_batch(ModelType.SIR_SUBGROUPS, y0.reshape(-1), np.concatenate([np.asarray(beta, dtype=float).reshape(-1), [gamma]]), t)[0]
    cols = {}
    for g in range(G):
        for c, n in enumerate(("susceptible", "infected", "removed")):
            cols[f"{n}{g}"] = rows[:, 3 * g + c]
    cols["time"] = np.arange(rows.shape[0])
    return pd.DataFrame(cols)


class _Mvn:
    """np.random.multivariate_normal(mean, cov) with the factorisation of cov cached while cov does not change: the same
    standard normals from numpy's global stream, the same arithmetic (numpy's legacy algorithm: x = z (sqrt(s) v) + mean
    with u, s, v = svd(cov)), hence the same proposals bit for bit -- without an SVD per MH iteration (60 us of a
    0.57 ms iteration at 10^3 particles)."""

    def __init__(self):
        self.cov = None
        self.factor = None

    def __call__(self, mean, cov, rng=np.random):
        cov = np.array(cov, dtype=np.double)
        if self.cov is None or cov.shape != self.cov.shape or not np.array_equal(cov, self.cov):
            _, sv, v = np.linalg.svd(cov)
            self.cov, self.factor = cov, np.sqrt(sv)[:, None] * v
        mean = np.array(mean, dtype=np.double)
        x = rng.standard_normal(mean.shape[0]).reshape(-1, mean.shape[0])
        x = np.dot(x, self.factor)
        x += mean
        return x.reshape(mean.shape[0])


def _setup(Y, type_model, theta_proposal, observations, probs, n_particles, n_population, mu, resampler, seed, arith,
           filter_id, store_history=True, block_particles=0):
    model = _model_id(type_model)
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError("Y must be (T, columns)")
    G = len(mu) if model >= 2 else 1
    theta = _flatten_theta(model, theta_proposal)
    arith = engine.resolve_arith(model, arith, theta=theta, n_particles=n_particles)   # 'auto' looks at the dynamics and the size
    cfg = engine.make_pf_config(model, n_particles, Y.shape[0], G=G, observations=observations, probs=probs,
                                resampler=resampler, arith=arith, seed=seed, filter_id0=filter_id,
                                mu=np.atleast_1d(mu), n_population=np.atleast_1d(n_population),
                                store_history=store_history, block_particles=block_particles)
    if Y.shape[1] != cfg.n_obs_cols:
        raise ValueError(f"Y has {Y.shape[1]} columns, the model observes {cfg.n_obs_cols}")
    return cfg, Y, theta


def particle_filter(Y, type_model, theta_proposal, observations=False, probs=.1, n_particles=1000, n_population=4820,
                    mu=20, jobs=4, *, resampler="systematic", seed=None, arith="auto", X0=None, replay=None,
                    output="numpy", filter_id=0, block_particles=0):
    """Bootstrap particle filter (pmcmc.py:123-233).

    Returns (zetas[T], hidden_process[T,N,C], ancestry_matrix[T,N]) as float64 numpy arrays like the reference,
    or (None, None, None) when the filter collapses (all weights zero, pmcmc.py:191-192).
    output='torch' returns CUDA tensors instead (zetas fp64; states / ancestry int32 (T,N,C) / (T,N));
    output='result' returns the engine.PfResult (adds log_zetas, n_events) or None on collapse.
    replay = dict(res_u, ssa_u, ssa_off) feeds the reference's uniform stream (needs X0) for bit-exact parity.
    """
    seed = engine.new_seed() if seed is None else seed
    cfg, Y, theta = _setup(Y, type_model, theta_proposal, observations, probs, n_particles, n_population, mu,
                           "multinomial" if replay is not None else resampler, seed, arith, filter_id,
                           block_particles=block_particles)
    res = engine.run_pf(cfg, Y, theta, X0=X0, replay=replay)
    status = int(res.status.cpu()[0])
    if status == -3:
        raise RuntimeError("replay buffer exhausted")
    if status != 0:
        return None if output == "result" else (None, None, None)
    if output == "result":
        return res
    zetas = torch.exp(res.log_zetas[0])
    if output == "torch":
        return zetas, res.X_hist[0].permute(0, 2, 1), res.ancestry[0]
    # device -> pinned host memory (torch's caching host allocator recycles the blocks once the arrays are dropped)
    hd = res.hidden_process(0)
    ad = res.ancestry[0].to(torch.float64)
    hidden = torch.empty(hd.shape, dtype=torch.float64, pin_memory=True)
    anc = torch.empty(ad.shape, dtype=torch.float64, pin_memory=True)
    zh = torch.empty(zetas.shape, dtype=torch.float64, pin_memory=True)
    hidden.copy_(hd, non_blocking=True); anc.copy_(ad, non_blocking=True); zh.copy_(zetas, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return zh.numpy(), hidden.numpy(), anc.numpy()


def pf_loglik(Y, type_model, theta_proposal, observations=False, probs=.1, n_particles=1000, n_population=4820,
              mu=20, *, resampler="systematic", seed=None, arith="auto", filter_id=0):
    """log of the likelihood estimate zetas[-1] only (no history leaves the device); -inf on collapse."""
    seed = engine.new_seed() if seed is None else seed
    cfg, Y, theta = _setup(Y, type_model, theta_proposal, observations, probs, n_particles, n_population, mu,
                           resampler, seed, arith, filter_id, store_history=False)
    res = engine.run_pf(cfg, Y, theta)
    st = int(res.status.cpu()[0])
    return float("-inf") if st != 0 else float(res.log_zetas[0, -1].cpu())


def particle_path_sampler(hidden_process, ancestry_matrix, *, exact_genealogy=False, chosen_path=None, seed=None):
    """pmcmc.py:236-248: pick a final particle uniformly and chase its ancestry backwards.  (T,N,C),(T,N) -> (T,C).

    The index chase runs in the sem_path_sample kernel; host arrays are uploaded first (pass CUDA tensors to avoid
    that).  exact_genealogy=False keeps the reference's indexing of ancestry_matrix[p]."""
    dev = engine.require_cuda()
    hp = torch.as_tensor(np.asarray(hidden_process) if not isinstance(hidden_process, torch.Tensor) else hidden_process)
    am = torch.as_tensor(np.asarray(ancestry_matrix) if not isinstance(ancestry_matrix, torch.Tensor) else ancestry_matrix)
    T, N, Cn = hp.shape
    X = hp.to(dev).to(torch.int32).permute(0, 2, 1).contiguous().unsqueeze(0)      # [1][T][C][N]
    A = am.to(dev).to(torch.int32).contiguous().unsqueeze(0)
    if chosen_path is None and seed is None:
        chosen_path = int(np.random.randint(0, N))                                # pmcmc.py:241
    res = engine.PfResult(None, X, A, None, None, None, 0)
    traj = res.path_sample(0, chosen=-1 if chosen_path is None else chosen_path, exact=exact_genealogy, seed=seed or 0)
    return traj.to(torch.float64).cpu().numpy()


def particle_mcmc(Y, type_model, parameters, h, adaptive=False, sigma=None, n_chains=1000, observations=False,
                  probs=.1, n_particles=1000, n_population=4820, mu=20, jobs=4, *, resampler="systematic", seed=None,
                  arith="auto", exact_genealogy=False, return_log=False, progress=False, stats=None, sharded=False,
                  group=None, exchange="auto", lookahead=1):
    """Particle marginal Metropolis-Hastings (pmcmc.py:251-408).

    Returns (thetas[n_chains,P], likelihoods[n_chains], sampled_trajs[T,n_chains,C]).  likelihoods are the linear
    zetas[-1] like the reference (log-likelihoods with return_log=True).  `stats`, if a dict, receives
    'filter_runs', 'acceptances', 'launches'.

    lookahead = L > 1 (or "auto"): small filters leave most of the GPU idle and one iteration is pure latency (a filter of
    10^3 particles takes as long as one of 3*10^4), so the next L proposals are evaluated in ONE launch (L filters with
    their own theta and p_obs) under the assumption that the earlier ones are rejected -- after a rejection the chain has
    not moved, so proposal k+1 is drawn from the same state, adaptive covariance included.  The first accepted
    proposal ends the batch and the later evaluations are discarded (they are independent of everything kept), so the
    chain has exactly the law of the sequential loop; with acceptance rate a a launch advances (1 - (1-a)^L) / a
    iterations instead of one.  lookahead=1 is the sequential loop of the reference (same numpy draw order).

    sharded=True (every rank of an initialised torch.distributed group calls with the same arguments): ONE chain whose
    filter of n_particles (global count) is sharded over the ranks with global systematic resampling (BASELINE config 5;
    globalises pmcmc.py:183-199 and the path sampler of pmcmc.py:236-248).  The MH draws come from a generator seeded
    once from rank 0, every rank receives the same likelihood and sampled trajectory from the device and takes the same
    decisions, so there is no host collective per iteration; all ranks return the same arrays.  exchange = "device"
    (peer-memory exchange inside one launch per rank), "host" (NCCL collectives per step, for shards too large to be
    co-resident) or "auto".
    """
    if sharded:
        return _particle_mcmc_sharded(Y, type_model, parameters, h, adaptive, sigma, n_chains, observations, probs,
                                      n_particles, n_population, mu, seed=seed, arith=arith, exact_genealogy=exact_genealogy,
                                      return_log=return_log, stats=stats, group=group, exchange=exchange)
    dev = engine.require_cuda()
    Y = np.asarray(Y, dtype=np.float64)
    model = _model_id(type_model)
    n_par = len(parameters)
    T = Y.shape[0]
    G = len(mu) if model >= 2 else 1
    Cn = engine.model_dims(model, G)[0]
    if lookahead == "auto":                                   # fill the GPU's idle share, up to 16 proposals per launch
        lookahead = int(max(1, min(16, 32768 // max(int(n_particles), 1))))
    if int(lookahead) > 1:
        return _particle_mcmc_lookahead(Y, type_model, parameters, h, adaptive, sigma, n_chains, observations, probs,
                                        n_particles, n_population, mu, resampler=resampler, seed=seed, arith=arith,
                                        exact_genealogy=exact_genealogy, return_log=return_log, stats=stats,
                                        lookahead=int(lookahead))
    thetas = np.zeros((n_chains, n_par))
    loglik = np.zeros(n_chains)
    sampled_trajs = np.zeros((T, n_chains, Cn))                                  # pmcmc.py:269-272
    std = np.eye(n_par) if sigma is None else sigma                               # :273-275
    seed = engine.new_seed() if seed is None else seed
    counters = dict(filter_runs=0, acceptances=1, launches=0, arith=set())

    # One iteration = ONE launch (the filter, the path sample of pmcmc.py:371 and the packing of the results happen in
    # the whole-filter kernel) + one small H2D (theta) + one small D2H (log-likelihood, status, sampled trajectory);
    # buffers, config and pointers are prepared once (engine.PreparedIteration).
    cfg0, _, _ = _setup(Y, type_model, _split(model, G, np.asarray(parameters, dtype=float), probs)[0], observations,
                        .5, n_particles, n_population, mu, resampler, seed, arith, 0)
    cfg0.path_exact = int(bool(exact_genealogy))
    prep = engine.PreparedIteration(cfg0, Y, dev)

    def run_filter(theta_vec, it):
        theta2, probs2 = _split(model, G, theta_vec, probs)
        th = _flatten_theta(model, theta2)
        a_now = engine.resolve_arith(model, arith, theta=th, n_particles=n_particles)    # 'auto' may pick another interval simulation
        r = prep.run(th, it, probs=probs2, arith=a_now)[0]                                # for another proposal: both are exact
        counters["arith"].add(engine.ARITH_NAMES.get(a_now, a_now))
        counters["filter_runs"] += 1
        counters["launches"] += prep.launches + (0 if prep.launches == 1 else 1)
        if int(r[1]) != 0:
            return None, None
        return float(r[0]), r[engine.ITER_HEADER:].reshape(T, Cn).copy()

    def finish_theta(theta_vec):
        if probs is None:                                                         # :313-314,373-374: stored p_obs is clipped
            out_t = np.array(theta_vec, dtype=float)
            out_t[-1] = max(min(out_t[-1], 1), 0)
            return out_t
        return theta_vec

    it = 0
    mvn = _Mvn()
    while True:                                                                   # :276-310 initial draw
        theta_proposal = mvn(np.array(parameters), h * std)
        if np.sum(theta_proposal < 0) > 0:
            continue
        lz, traj = run_filter(theta_proposal, it)
        it += 1
        if lz is not None:
            break
    thetas[0] = finish_theta(theta_proposal)
    loglik[0] = lz
    sampled_trajs[:, 0, :] = traj

    bar = None
    if progress:
        from tqdm import tqdm
        bar = tqdm(total=n_chains - 1, desc="Chains")
    for i in range(1, n_chains):                                                  # :325
        if adaptive and i > 1e3:
            std = np.cov(thetas[:i].T, ddof=0) + 1e-4 * np.eye(n_par)             # :327-328
        theta_proposal = mvn(thetas[i - 1], h * std)                              # :330
        lz = None
        if not np.sum(theta_proposal < 0) > 0:                                    # :333
            lz, traj = run_filter(theta_proposal, it)
            it += 1
        if lz is None:                                                            # negative proposal or collapse (:333-337,365-369)
            thetas[i] = thetas[i - 1]; loglik[i] = loglik[i - 1]; sampled_trajs[:, i, :] = sampled_trajs[:, i - 1, :]
        elif np.log(np.random.uniform()) < lz - loglik[i - 1]:                    # :376-395, see D4
            counters["acceptances"] += 1
            thetas[i] = finish_theta(theta_proposal); loglik[i] = lz; sampled_trajs[:, i, :] = traj
        else:
            thetas[i] = thetas[i - 1]; loglik[i] = loglik[i - 1]; sampled_trajs[:, i, :] = sampled_trajs[:, i - 1, :]
        if bar is not None:
            bar.update(1)
            bar.set_postfix_str(f"theta={thetas[i]}, logZ={loglik[i]:.3f}, acc={100 * counters['acceptances'] / (i + 1):.1f}%")
    if isinstance(stats, dict):
        stats.update(counters)
    return thetas, (loglik if return_log else np.exp(loglik)), sampled_trajs


def _particle_mcmc_lookahead(Y, type_model, parameters, h, adaptive, sigma, n_chains, observations, probs, n_particles,
                             n_population, mu, *, resampler, seed, arith, exact_genealogy, return_log, stats, lookahead):
    """particle_mcmc with `lookahead` proposals per launch (see particle_mcmc): pmcmc.py:276-403 with the filter calls of
    consecutive iterations batched under the rejection assumption."""
    dev = engine.require_cuda()
    model = _model_id(type_model)
    n_par = len(parameters)
    T = Y.shape[0]
    G = len(mu) if model >= 2 else 1
    Cn = engine.model_dims(model, G)[0]
    L = int(lookahead)
    thetas = np.zeros((n_chains, n_par))
    loglik = np.zeros(n_chains)
    sampled_trajs = np.zeros((T, n_chains, Cn))
    std = np.eye(n_par) if sigma is None else sigma
    seed = engine.new_seed() if seed is None else seed
    counters = dict(filter_runs=0, acceptances=1, launches=0, lookahead=L, discarded=0)
    th0, pr0 = _split(model, G, np.asarray(parameters, dtype=float), probs)
    th0 = _flatten_theta(model, th0)
    while True:                                               # the whole batch must be ONE launch (all filters co-resident)
        cfg0 = engine.make_pf_config(model, n_particles, T, G=G, n_filters=L, observations=observations, probs=.5,
                                     resampler=resampler, arith=engine.resolve_arith(model, arith, theta=th0, n_particles=n_particles * L), seed=seed,
                                     mu=np.atleast_1d(mu), n_population=np.atleast_1d(n_population), path_exact=exact_genealogy)
        if Y.shape[1] != cfg0.n_obs_cols:
            raise ValueError(f"Y has {Y.shape[1]} columns, the model observes {cfg0.n_obs_cols}")
        if L == 1 or engine._lib.load().sem_pf_launch_count(engine.C.byref(cfg0)) == 1:
            break
        L = max(1, L // 2)
    counters["lookahead"] = L
    prep = engine.PreparedIteration(cfg0, Y, dev)
    n_th = th0.size                                           # model parameters (the optional trailing p_obs excluded)
    th_batch = np.tile(th0, (L, 1))
    pr_batch = np.full(L, .5 if pr0 is None else float(pr0))
    state = dict(it=0)

    def run_batch(props, theta_ref):
        """Filters for the non-negative proposals of the batch (the others are rejected without a run, pmcmc.py:333).
        props (n, n_par).  Returns (ok[n], loglik[n], packed results) -- host work is vectorised: with 10^3 particles a
        launch takes 0.57 ms and every 40 us of Python per proposal would show."""
        n = len(props)
        live = ~np.any(props < 0, axis=1)
        th_batch[:] = theta_ref[:n_th]                        # idle slots re-run the current state (ignored)
        th_batch[:n][live] = props[live][:, :n_th]            # (beta[G,G] row-major, gamma) is the vector itself, pmcmc.py:289-296
        if probs is None:
            pr_batch[:] = min(max(theta_ref[-1], 0), 1)
            pr_batch[:n][live] = np.clip(props[live][:, -1], 0, 1)                # :313-314
        if not live.any():
            return live, np.zeros(n), None
        first = int(np.argmax(live))
        r = prep.run(th_batch, state["it"], probs=pr_batch, arith=engine.resolve_arith(model, arith, theta=th_batch[first], n_particles=n_particles * L))
        state["it"] += L
        counters["filter_runs"] += int(live.sum())
        counters["launches"] += prep.launches + (0 if prep.launches == 1 else 1)
        ok = live & (r[:n, 1] == 0)                            # collapsed filters count as rejections (:365-369)
        return ok, r[:n, 0], r

    def finish_theta(theta_vec):
        if probs is None:
            out_t = np.array(theta_vec, dtype=float)
            out_t[-1] = max(min(out_t[-1], 1), 0)
            return out_t
        return theta_vec

    def proposals(mean, cov, n):
        """n draws of N(mean, cov) (pmcmc.py:330) -- one factorisation, one call for the standard normals"""
        w, v = np.linalg.eigh(np.asarray(cov, dtype=float))
        return mean + (np.random.standard_normal((n, n_par)) * np.sqrt(np.maximum(w, 0))) @ v.T

    start = np.array(parameters, dtype=float)
    done = False
    while not done:                                                               # pmcmc.py:276-310 initial draw
        props = proposals(start, h * std, L)
        ok, lz, r = run_batch(props, start)
        if ok.any():
            k = int(np.argmax(ok))
            thetas[0] = finish_theta(props[k]); loglik[0] = lz[k]
            sampled_trajs[:, 0, :] = r[k, engine.ITER_HEADER:].reshape(T, Cn)
            counters["discarded"] += L - 1 - k
            done = True

    # running sums of the chain for the adaptive covariance (:327-328) under the rejection assumption: iterations
    # i .. i+k-1 repeat thetas[i-1], so cov(thetas[:i+k]) follows from the sums over thetas[:i] in O(P^2)
    s1 = thetas[0].copy(); s2 = np.outer(thetas[0], thetas[0])
    i = 1
    while i < n_chains:                                                           # pmcmc.py:325-403
        n = min(L, n_chains - i)
        cur = thetas[i - 1]
        if adaptive and i + n - 1 > 1e3:
            props = np.empty((n, n_par))
            for k in range(n):                                                    # iteration i + k, if i .. i+k-1 all reject
                if i + k > 1e3:
                    m = (s1 + k * cur) / (i + k)
                    std = (s2 + k * np.outer(cur, cur)) / (i + k) - np.outer(m, m) + 1e-4 * np.eye(n_par)
                props[k] = proposals(cur, h * std, 1)[0]
        else:
            props = proposals(cur, h * std, n)
        us = np.random.uniform(size=n)
        ok, lz, r = run_batch(props, cur)
        acc = ok & (np.log(us) < lz - loglik[i - 1])                              # :376-395, see D4
        k_acc = int(np.argmax(acc)) if acc.any() else n
        nrej = min(k_acc, n)
        if nrej:                                                                  # rejections: the chain stays
            thetas[i:i + nrej] = cur; loglik[i:i + nrej] = loglik[i - 1]
            sampled_trajs[:, i:i + nrej, :] = sampled_trajs[:, i - 1:i, :]
            s1 += nrej * cur; s2 += nrej * np.outer(cur, cur)
            i += nrej
        if k_acc < n:
            counters["acceptances"] += 1
            thetas[i] = finish_theta(props[k_acc]); loglik[i] = lz[k_acc]
            sampled_trajs[:, i, :] = r[k_acc, engine.ITER_HEADER:].reshape(T, Cn)
            s1 += thetas[i]; s2 += np.outer(thetas[i], thetas[i])
            i += 1
            counters["discarded"] += n - 1 - k_acc
    if isinstance(stats, dict):
        stats.update(counters)
    return thetas, (loglik if return_log else np.exp(loglik)), sampled_trajs


def _particle_mcmc_sharded(Y, type_model, parameters, h, adaptive, sigma, n_chains, observations, probs, n_particles,
                           n_population, mu, *, seed, arith, exact_genealogy, return_log, stats, group, exchange):
    """particle_mcmc with the filter sharded over the ranks of `group` (see particle_mcmc)."""
    import torch.distributed as dist
    from . import sharded as sh
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("particle_mcmc(sharded=True) needs an initialised torch.distributed process group")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if n_particles % world:
        raise ValueError("n_particles must be divisible by the number of ranks")
    Y = np.asarray(Y, dtype=np.float64)
    model = _model_id(type_model)
    n_par = len(parameters)
    T = Y.shape[0]
    G = len(mu) if model >= 2 else 1
    Cn = engine.model_dims(model, G)[0]
    box = [engine.new_seed() if seed is None else int(seed), engine.new_seed()]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    seed, mh_seed = box
    rng = np.random.RandomState(mh_seed % 2**32)             # the same MH draws on every rank
    thetas = np.zeros((n_chains, n_par))
    loglik = np.zeros(n_chains)
    sampled_trajs = np.zeros((T, n_chains, Cn))
    std = np.eye(n_par) if sigma is None else sigma
    counters = dict(filter_runs=0, acceptances=1, launches=0, exchange=None)
    th0 = _flatten_theta(model, _split(model, G, np.asarray(parameters, dtype=float), probs)[0])
    kw = dict(G=G, observations=observations, probs=.5 if probs is None else probs, seed=seed, mu=np.atleast_1d(mu),
              n_population=np.atleast_1d(n_population))
    n_local = n_particles // world
    if exchange == "auto":
        exchange = "device" if sh.device_exchange_supported(model, n_local, T, world, arith=engine.resolve_arith(model, arith, theta=th0, n_particles=n_local),
                                                            **kw) else "host"
    counters["exchange"] = exchange
    pf = None
    if exchange == "device":
        pf = sh.cached_peer_filter(group, rank, world, model, Y, n_local, arith=arith, theta=th0, path_exact=exact_genealogy, **kw)

    def run_filter(theta_vec, it):
        theta2, probs2 = _split(model, G, theta_vec, probs)
        th = _flatten_theta(model, theta2)
        counters["filter_runs"] += 1
        if pf is not None:
            r = pf.iteration(th, it, probs=probs2, arith=engine.resolve_arith(model, arith, theta=th, n_particles=n_local))
            counters["launches"] += 1
            if int(r[1]) != 0:                               # same status on every rank: clean the arenas together
                if int(r[1]) == engine._lib.ERR_PEER:
                    raise engine._lib.SemError("sharded filter: a peer rank did not answer (SEM_ERR_PEER)")
                pf.reset()
                dist.barrier(group=group)
                return None, None
            return float(r[0]), r[engine.ITER_HEADER:].reshape(T, Cn).copy()
        out = sh.run_distributed(Y, model, th, n_particles, group=group, filter_id=it, arith=arith, **dict(kw, probs=probs2))
        counters["launches"] += out["shard"].launches
        if out["collapsed"]:
            return None, None
        traj = sh.path_sample_distributed(out["shard"], seed, it, exact=exact_genealogy, group=group)
        return float(out["log_zetas"][-1]), traj

    def finish_theta(theta_vec):
        if probs is None:
            out_t = np.array(theta_vec, dtype=float)
            out_t[-1] = max(min(out_t[-1], 1), 0)
            return out_t
        return theta_vec

    if True:
        it = 0
        mvn = _Mvn()
        while True:                                                               # pmcmc.py:276-310
            theta_proposal = mvn(np.array(parameters), h * std, rng)
            if np.sum(theta_proposal < 0) > 0:
                continue
            lz, traj = run_filter(theta_proposal, it)
            it += 1
            if lz is not None:
                break
        thetas[0] = finish_theta(theta_proposal); loglik[0] = lz; sampled_trajs[:, 0, :] = traj
        for i in range(1, n_chains):                                              # pmcmc.py:325-403
            if adaptive and i > 1e3:
                std = np.cov(thetas[:i].T, ddof=0) + 1e-4 * np.eye(n_par)
            theta_proposal = mvn(thetas[i - 1], h * std, rng)
            lz = None
            if not np.sum(theta_proposal < 0) > 0:
                lz, traj = run_filter(theta_proposal, it)
                it += 1
            if lz is not None and np.log(rng.uniform()) < lz - loglik[i - 1]:
                counters["acceptances"] += 1
                thetas[i] = finish_theta(theta_proposal); loglik[i] = lz; sampled_trajs[:, i, :] = traj
            else:
                thetas[i] = thetas[i - 1]; loglik[i] = loglik[i - 1]; sampled_trajs[:, i, :] = sampled_trajs[:, i - 1, :]
    if isinstance(stats, dict):
        stats.update(counters)
    return thetas, (loglik if return_log else np.exp(loglik)), sampled_trajs


def _split(model, G, theta_vec, probs):
    """pmcmc.py:283-296 / 339-352: optional trailing p_obs (probs=None), subgroup reshape to (beta[G,G], gamma)."""
    theta_vec = np.asarray(theta_vec, dtype=np.float64)
    probs2 = probs
    if probs is None:
        probs2 = max(min(theta_vec[-1], 1), 0)
        theta_vec = theta_vec[:-1]
    if model >= 2:
        beta = np.zeros((G, G))
        for p in range(G * G):
            beta[p // G, p % G] = theta_vec[p]
        return (beta, theta_vec[-1]), probs2
    return theta_vec, probs2
