"""Drop-in for the reference's abc_algo.py: ABC rejection sampling for SIR on the GPU (sem_abc_run).

abc_algo(observed_data, no_of_samples, threshold, priors) keeps the reference signature (abc_algo.py:17) and
return value: ({"beta": [...], "gamma": [...]}, trajectories[no_of_samples, T, 4]) with rows [day, S, I, R].

The reference draws trials one after another until no_of_samples are accepted.  Here trials are numbered
0,1,2,... (the trial id keys its Philox streams), run in large batches, and the FIRST no_of_samples accepted
trials in id order are returned -- the same sequential-acceptance semantics, independent of batch size and of the
number of GPUs.  With torch.distributed initialised, the ids of each batch are split across ranks (independent
trials shard trivially; the only exchange is the gather of the accepted few).
"""
import numpy as np
import torch

from . import engine

__all__ = ["abc_algo", "distance_function", "abc_rejection"]


def distance_function(I_1, I_2, R_1, R_2):
    """abc_algo.py:10-13 (host helper kept for API compatibility; the kernels accumulate it on the fly)."""
    return (np.mean(abs(I_1 - I_2)) + np.mean(abs(R_1 - R_2))) / 2


def shard_range(start, count, rank, world):
    """Contiguous split of trial ids [start, start+count) over `world` ranks -> (start_r, count_r)."""
    base, rem = divmod(count, world)
    lo = start + rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


def _device_batch(obs, n, threshold, prior4, seed, trial0, arith, early_reject):
    """Run trials [trial0, trial0+n) on this GPU; return (ids, theta, distance) of the accepted ones (host)."""
    out = engine.abc_trials(obs, n, threshold, prior4, seed=seed, trial0=trial0, arith=arith, early_reject=early_reject)
    acc = torch.nonzero(out["distance"] <= threshold).flatten()          # ascending = trial order
    ids = (acc + trial0).cpu().numpy().astype(np.int64)
    return ids, out["theta"][acc].cpu().numpy(), out["distance"][acc].cpu().numpy(), int(out["n_events"].cpu()[0])


def _device_trajectories(obs, ids, threshold, prior4, seed, arith):
    """Re-simulate accepted trials by id (counter-based RNG => identical epidemics) to emit their [T,3] rows."""
    out = engine.abc_trials(obs, len(ids), threshold, prior4, seed=seed, trial_ids=ids, arith=arith, early_reject=False,
                            want_traj=True)
    return out["traj"].cpu().numpy()


def abc_rejection(observed_data, no_of_samples, threshold, priors, *, seed=None, batch=None, arith="fast32",
                  early_reject=True, max_trials=None, run_batch=None, run_traj=None, stats=None):
    """Engine behind abc_algo.  run_batch / run_traj are injectable (tests drive the sharding logic with a CPU
    executor); by default they launch the CUDA kernels."""
    obs = np.asarray(observed_data, dtype=np.float64)
    T = obs.shape[0]
    prior4 = [priors["beta"][0], priors["beta"][1], priors["gamma"][0], priors["gamma"][1]]
    run_batch = run_batch or _device_batch
    run_traj = run_traj or _device_trajectories
    dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    seed = engine.new_seed() if seed is None else seed
    if dist and world > 1:                                            # ONE Philox key for all ranks: trial ids are global, and every
        box = [seed]                                                  # rank re-simulates the accepted ids (rank 0's seed wins)
        dist.broadcast_object_list(box, src=0)
        seed = box[0]
    batch = batch or max(1 << 16, 64 * no_of_samples)
    next_id, trials, events = 0, 0, 0
    acc_ids, acc_theta, acc_dist = [], [], []
    while sum(len(a) for a in acc_ids) < no_of_samples:
        if max_trials is not None and next_id >= max_trials:
            break
        lo, cnt = shard_range(next_id, batch, rank, world)
        ids, th, d, ev = run_batch(obs, cnt, threshold, prior4, seed, lo, arith, early_reject) if cnt else \
            (np.zeros(0, np.int64), np.zeros((0, 2)), np.zeros(0), 0)
        if dist:
            gathered = [None] * world
            dist.all_gather_object(gathered, (ids, th, d, ev))
        else:
            gathered = [(ids, th, d, ev)]
        for g_ids, g_th, g_d, g_ev in gathered:                       # rank order == id order (contiguous shards)
            acc_ids.append(g_ids); acc_theta.append(g_th); acc_dist.append(g_d); events += g_ev
        next_id += batch
        trials += batch
        got = sum(len(a) for a in acc_ids)
        if 0 < got < no_of_samples:                                   # size the next batch from the acceptance rate
            need = (no_of_samples - got) * trials / got
            batch = int(min(max(need * 1.2, 1 << 14), 1 << 24))
    ids = np.concatenate(acc_ids)[:no_of_samples]
    theta = np.concatenate(acc_theta)[:no_of_samples]
    dists = np.concatenate(acc_dist)[:no_of_samples]
    if len(ids):
        sir = run_traj(obs, ids, threshold, prior4, seed, arith)      # every rank re-simulates the same few ids
    else:
        sir = np.zeros((0, T, 3), dtype=np.int32)
    traj = np.zeros((len(ids), T, 4))
    traj[:, :, 0] = np.arange(T)[None, :]                              # abc_algo.py:58-64 column 0 = day
    traj[:, :, 1:] = sir
    if isinstance(stats, dict):
        stats.update(trials=trials, accepted_ids=ids, distances=dists, n_events=events, seed=seed)
    return theta, traj


def abc_algo(observed_data, no_of_samples, threshold, priors, *, seed=None, batch=None, arith="fast32",
             early_reject=True, max_trials=None, stats=None):
    """ABC rejection sampler (abc_algo.py:17-109).  observed_data (T,3) [S,I,R]; priors {'beta':[lo,hi],'gamma':[lo,hi]}."""
    theta, traj = abc_rejection(observed_data, no_of_samples, threshold, priors, seed=seed, batch=batch, arith=arith,
                                early_reject=early_reject, max_trials=max_trials, stats=stats)
    posterior_distr = {"beta": [float(b) for b in theta[:, 0]], "gamma": [float(g) for g in theta[:, 1]]}
    return posterior_distr, traj
