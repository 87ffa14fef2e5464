"""Device plumbing between the drop-in Python entry points and the C ABI (include/sem_b200.h).

PyTorch is used for device memory, streams and torch.distributed only; all arithmetic of the hot path
runs in the hand-written kernels of libsem_b200.so.  There is no CPU fallback: a missing library or a
missing CUDA device raises.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

MODEL_NAMES = {"sir": 0, "seir": 1, "sir_subgroups": 2, "sir_subgroups2": 3}
RESAMPLERS = {"multinomial": 0, "systematic": 1}
ARITH = {"reference": 0, "fast": 1, "uniformized": 2, "fast32": 3, "uniformized32": 4}
ARITH_NAMES = {v: k for k, v in ARITH.items()}


AUTO_MAX_GROWTH = 0.5    # per observation interval; measured crossover between the two interval simulations
AUTO_SMALL_FILTER = 20000   # particles (all filters of a launch) below which a step is latency-bound (profiles/r02f_arith_sweep.txt)


def resolve_arith(model, arith, theta=None, dt=1.0, n_particles=None):
    """'auto' (the filters' default) = the faster exact interval simulation for the model family and dynamics:
    uniformized intervals with 32-bit candidates for SIR / SEIR (2-3 reactions) -- unless the filter is small (a step is
    then bound by the latency of one particle's chain, n_particles when known) AND the epidemic's early growth per
    observation interval (from theta, when it is known on the host) is so fast that an interval needs many short
    batches, each with its own set-up latency -- and the direct method with 32-bit streams otherwise: for the subgroup models (their G^2+G propensities
    make a uniformized candidate as dear as a direct event), for ABC and for simulations that log event times."""
    if arith == "auto":            # (model None: simulations with event logs and ABC trials need event times -> direct method)
        if model not in (0, 1):
            return ARITH["fast32"]
        if theta is not None and not isinstance(theta, torch.Tensor):
            th = np.asarray(theta, dtype=np.float64).reshape(-1)
            if model == 0 and th.size >= 2:
                growth = (th[0] - th[1]) * dt                                    # SIR: beta - gamma
            elif model == 1 and th.size >= 3:                                    # SEIR: dominant eigenvalue of the (E, I) linearisation
                growth = 0.5 * (np.sqrt((th[1] - th[2]) ** 2 + 4 * th[1] * th[0]) - (th[1] + th[2])) * dt
            else:
                growth = 0.0
            if growth > AUTO_MAX_GROWTH and (n_particles is None or n_particles < AUTO_SMALL_FILTER):
                return ARITH["fast32"]
        return ARITH["uniformized32"]
    return ARITH.get(arith, arith)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.SemError("no CUDA device: the B200 engine has no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise _lib.SemError("device must be a CUDA device")
    return dev


def model_dims(model, G):
    C_ = 3 if model == 0 else 4 if model == 1 else 3 * G
    P = 2 if model == 0 else 3 if model == 1 else G * G + 1
    Cobs = 3 if model == 3 else C_
    return C_, P, Cobs


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f64(a, dev):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev, non_blocking=True)


def check_rates(theta, what="theta"):
    """Host-side guard of the simulators' inputs: a negative rate constant gives a negative propensity, which the
    reference refuses inside np.random.choice (gillespie_algo.py:63, "probabilities are not non-negative") -- and which
    would run the kernels' blocked event loops backwards in time.  Device-resident inputs are the caller's job."""
    if theta is None or isinstance(theta, torch.Tensor):
        return
    th = np.asarray(theta, dtype=np.float64)
    if (th < 0).any():
        raise ValueError(f"probabilities are not non-negative ({what} holds a negative rate)")


def new_seed():
    """A Philox seed drawn from numpy's global generator, so np.random.seed() makes runs reproducible
    the same way it does for the reference."""
    return int(np.random.randint(0, 2**31 - 1)) * 2**31 + int(np.random.randint(0, 2**31 - 1))


@dataclass
class PfResult:
    """Device-resident result of one sem_pf_run (n_filters filters)."""
    log_zetas: torch.Tensor      # [F, T] float64
    X_hist: torch.Tensor         # [F, rows, C, N] int32 (SoA)
    ancestry: torch.Tensor       # [F, rows, N] int32
    status: torch.Tensor         # [F] int32
    n_events: torch.Tensor       # [F] int64
    cfg: _lib.PfConfig
    launches: int

    def collapsed(self):
        return self.status.cpu().numpy()

    def hidden_process(self, f=0):
        """(T,N,C) float64 on the device, the layout pmcmc.py:151 returns."""
        F, T, Cn, N = self.X_hist.shape
        out = torch.empty((T, N, Cn), dtype=torch.float64, device=self.X_hist.device)
        _lib.check(_lib.load().sem_hist_to_f64(_ptr(self.X_hist[f]), T, N, Cn, _ptr(out), _stream()), "sem_hist_to_f64")
        return out

    def path_sample(self, f=0, chosen=-1, exact=False, seed=0, filter_id=0):
        F, T, Cn, N = self.X_hist.shape
        traj = torch.empty((T, Cn), dtype=torch.int32, device=self.X_hist.device)
        _lib.check(_lib.load().sem_path_sample(_ptr(self.X_hist[f]), _ptr(self.ancestry[f]), T, N, Cn, int(chosen),
                                               int(bool(exact)), C.c_uint64(seed), C.c_uint32(filter_id), _ptr(traj),
                                               _stream()), "sem_path_sample")
        return traj


def make_pf_config(model, N, T, G=1, n_filters=1, observations=False, probs=.1, resampler="systematic", arith="auto",
                   seed=0, filter_id0=0, mu=None, n_population=None, dt=1.0, store_history=True, block_particles=0,
                   launch_per_step=False, path_exact=False, grid_barrier=False, theta=None):
    Cn, P, Cobs = model_dims(model, G)
    cfg = _lib.PfConfig(model=model, obs_kind=int(bool(observations)), resampler=RESAMPLERS.get(resampler, resampler),
                        arith=resolve_arith(model, arith, theta=theta, n_particles=int(N) * int(n_filters)), n_particles=int(N), n_obs=int(T), n_groups=int(G),
                        n_obs_cols=Cobs, n_filters=int(n_filters), block_particles=int(block_particles),
                        store_history=int(bool(store_history)), reserved=1 if launch_per_step else 2 if grid_barrier else 0, probs=float(probs),
                        dt=float(dt), path_exact=int(bool(path_exact)),
                        seed=int(seed) & (2**64 - 1), filter_id0=int(filter_id0) & 0xFFFFFF)
    if mu is not None:
        mu = np.asarray(mu, dtype=np.float64).reshape(-1)
        npop = np.asarray(n_population, dtype=np.float64).reshape(-1)
        for g in range(len(mu)):
            cfg.mu[g] = float(mu[g]); cfg.n_population[g] = float(npop[g])
    return cfg


ITER_HEADER = 4      # SEM_ITER_HEADER: log_zetas[T-1], status, n_events, chosen final particle


def run_pf(cfg, Y, theta, X0=None, replay=None, device=None, out=None, iter_out=None):
    """Enqueue one particle-filter pass (all T steps) on the current stream.  Returns PfResult (device tensors).

    iter_out: optional CUDA float64 tensor [F, ITER_HEADER + T*C]; the pass then ends with the path sample of
    particle_path_sampler and packs the results one MH iteration reads (sem_b200.h: iteration_result).

    Y (T,Cobs), theta (F,P) or (P,), X0 (N,C) or None (Poisson init on device), replay = dict(res_u (T-1,N),
    ssa_u flat, ssa_off CSR) for the bit-exact replay mode.
    """
    L = _lib.load()
    dev = require_cuda(device)
    F, T, N = cfg.n_filters, cfg.n_obs, cfg.n_particles
    G = cfg.n_groups if cfg.model >= 2 else 1
    Cn, P, Cobs = model_dims(cfg.model, G)
    check_rates(theta)
    with torch.cuda.device(dev):
        Yd = _dev_f64(Y, dev).reshape(T, Cobs)
        thd = _dev_f64(theta, dev).reshape(F, P)
        X0d = None
        if X0 is not None:
            X0n = np.asarray(X0).reshape(N, Cn) if not isinstance(X0, torch.Tensor) else X0.reshape(N, Cn)
            X0d = torch.as_tensor(X0n).to(dev).to(torch.int32).t().contiguous()         # SoA [C][N]
        rows = T if cfg.store_history else min(T, 2)
        if out is None:
            X_hist = torch.empty((F, rows, Cn, N), dtype=torch.int32, device=dev)
            anc = torch.empty((F, rows, N), dtype=torch.int32, device=dev)
            logz = torch.empty((F, T), dtype=torch.float64, device=dev)
            status = torch.empty((F,), dtype=torch.int32, device=dev)
            nev = torch.empty((F,), dtype=torch.int64, device=dev)
            ws = torch.empty((L.sem_pf_workspace_bytes(C.byref(cfg)),), dtype=torch.uint8, device=dev)
        else:
            X_hist, anc, logz, status, nev, ws = out
        buf = _lib.PfBuffers(Y=_ptr(Yd), theta=_ptr(thd), X0=_ptr(X0d), X_hist=_ptr(X_hist), ancestry=_ptr(anc),
                             log_zetas=_ptr(logz), status=_ptr(status), n_events=_ptr(nev), workspace=_ptr(ws),
                             iteration_result=_ptr(iter_out))
        keep = [Yd, thd, X0d, ws]
        if replay is not None:
            ru = _dev_f64(replay["res_u"], dev); su = _dev_f64(replay["ssa_u"], dev)
            so = torch.from_numpy(np.ascontiguousarray(replay["ssa_off"], dtype=np.int64)).to(dev)
            buf.replay_resample_u = _ptr(ru); buf.replay_ssa_u = _ptr(su); buf.replay_ssa_off = _ptr(so)
            keep += [ru, su, so]
        _lib.check(L.sem_pf_run(C.byref(cfg), C.byref(buf), _stream()), "sem_pf_run")
        res = PfResult(logz, X_hist, anc, status, nev, cfg, L.sem_pf_launch_count(C.byref(cfg)))
        res._keep = keep      # inputs stay alive until the stream has consumed them
        return res


class PreparedIteration:
    """Everything one PMCMC iteration needs, set up once: device outputs + workspace, Y, a device theta fed from pinned
    host memory, the packed iteration result and its pinned mirror.  run(theta, filter_id, probs, arith) is then one H2D
    of theta, ONE sem_pf_run (filter + path sample + packing, sem_b200.h: iteration_result), one D2H and one stream
    synchronisation -- no allocation and no Python-side tensor work per iteration."""

    def __init__(self, cfg, Y, device=None):
        self.L = _lib.load()
        self.dev = require_cuda(device)
        self.cfg = cfg
        F, T = cfg.n_filters, cfg.n_obs
        G = cfg.n_groups if cfg.model >= 2 else 1
        self.Cn, self.P, _ = model_dims(cfg.model, G)
        self.T = T
        with torch.cuda.device(self.dev):
            self.out = alloc_pf_outputs(cfg, self.dev)
            self.Y = _dev_f64(Y, self.dev).reshape(T, -1)
            self.pin_th = torch.empty((F, self.P), dtype=torch.float64).pin_memory()
            self.dev_th = torch.empty((F, self.P), dtype=torch.float64, device=self.dev)
            self.dev_it = torch.empty((F, ITER_HEADER + T * self.Cn), dtype=torch.float64, device=self.dev)
            self.pin_it = torch.empty((F, ITER_HEADER + T * self.Cn), dtype=torch.float64).pin_memory()
            self.pin_pr = torch.empty((F,), dtype=torch.float64).pin_memory()          # observation parameter per filter
            self.dev_pr = torch.empty((F,), dtype=torch.float64, device=self.dev)
        X_hist, anc, logz, status, nev, ws = self.out
        self.buf = _lib.PfBuffers(Y=_ptr(self.Y), theta=_ptr(self.dev_th), X0=None, X_hist=_ptr(X_hist), ancestry=_ptr(anc),
                                  log_zetas=_ptr(logz), status=_ptr(status), n_events=_ptr(nev), workspace=_ptr(ws),
                                  iteration_result=_ptr(self.dev_it))
        self.th_host = self.pin_th.numpy()
        self.it_host = self.pin_it.numpy()
        self.pr_host = self.pin_pr.numpy()
        self.launches = self.L.sem_pf_launch_count(C.byref(cfg))

    def run(self, theta, filter_id, probs=None, arith=None):
        """Returns the pinned [F, ITER_HEADER + T*C] result (valid until the next run).  theta (F,P) or (P,); probs a
        scalar, or one value per filter (a batch of proposals that also differ in the observation parameter)."""
        cfg = self.cfg
        cfg.filter_id0 = int(filter_id) & 0xFFFFFF
        per_filter = probs is not None and np.ndim(probs) > 0
        if probs is not None and not per_filter:
            cfg.probs = float(probs)
        if arith is not None:
            cfg.arith = int(arith)
        check_rates(theta)
        self.th_host[...] = theta
        self.buf.probs_per_filter = _ptr(self.dev_pr) if per_filter else None
        if per_filter:
            self.pr_host[...] = probs
        stream = torch.cuda.current_stream(self.dev)
        _lib.check(self.L.sem_pf_iteration(C.byref(cfg), C.byref(self.buf), C.c_void_p(self.pin_th.data_ptr()),
                                           C.c_void_p(self.pin_pr.data_ptr()) if per_filter else None,
                                           C.c_void_p(self.pin_it.data_ptr()), C.c_void_p(stream.cuda_stream)), "sem_pf_iteration")
        stream.synchronize()
        return self.it_host


def alloc_pf_outputs(cfg, device=None):
    """Pre-allocated output + workspace tensors for repeated run_pf calls (e.g. the MH loop)."""
    L = _lib.load()
    dev = require_cuda(device)
    F, T, N = cfg.n_filters, cfg.n_obs, cfg.n_particles
    G = cfg.n_groups if cfg.model >= 2 else 1
    Cn, _, _ = model_dims(cfg.model, G)
    rows = T if cfg.store_history else min(T, 2)
    return (torch.empty((F, rows, Cn, N), dtype=torch.int32, device=dev),
            torch.empty((F, rows, N), dtype=torch.int32, device=dev),
            torch.empty((F, T), dtype=torch.float64, device=dev),
            torch.empty((F,), dtype=torch.int32, device=dev),
            torch.empty((F,), dtype=torch.int64, device=dev),
            torch.empty((L.sem_pf_workspace_bytes(C.byref(cfg)),), dtype=torch.uint8, device=dev))


def simulate(model, x0, theta, max_time, G=1, arith="fast32", seed=0, sim_index0=0, record_capacity=0, replay=None,
             n_sims=None, daily_rows=0, device=None):
    """Batch of independent SSA runs (sem_ssa_simulate).  x0 (n,C) or (C,), theta (n,P) or (P,).
    Returns dict(x (n,C) int32, n_rows (n,), times (n,cap), states (n,cap,C)) as device tensors."""
    L = _lib.load()
    dev = require_cuda(device)
    Cn, P, _ = model_dims(model, G)
    check_rates(theta)
    with torch.cuda.device(dev):
        x0t = torch.as_tensor(np.asarray(x0)).to(dev).to(torch.int32).contiguous()
        tht = _dev_f64(theta, dev)
        shared_x0 = x0t.dim() == 1
        shared_th = tht.dim() == 1
        n = 1
        if not shared_x0:
            n = x0t.shape[0]
        elif not shared_th:
            n = tht.shape[0]
        if replay is not None:
            n = len(replay["off"]) - 1
        elif n_sims is not None:
            n = int(n_sims)
        cfg = _lib.SimConfig(model=model, n_groups=G, arith=resolve_arith(None, arith), n_sims=n,
                             shared_theta=int(shared_th), shared_x0=int(shared_x0), record_capacity=int(record_capacity),
                             max_time=float(max_time), seed=int(seed) & (2**64 - 1), sim_index0=int(sim_index0),
                             daily_rows=int(daily_rows))
        x_out = torch.empty((n, Cn), dtype=torch.int32, device=dev)
        n_rows = torch.empty((n,), dtype=torch.int64, device=dev)
        times = states = ru = ro = None
        if record_capacity > 0:
            times = torch.empty((n, record_capacity), dtype=torch.float64, device=dev)
            states = torch.empty((n, record_capacity, Cn), dtype=torch.int32, device=dev)
        elif daily_rows > 0:
            states = torch.empty((n, daily_rows, Cn), dtype=torch.int32, device=dev)
        if replay is not None:
            ru = _dev_f64(replay["u"], dev)
            ro = torch.from_numpy(np.ascontiguousarray(replay["off"], dtype=np.int64)).to(dev)
        _lib.check(L.sem_ssa_simulate(C.byref(cfg), _ptr(x0t), _ptr(tht), _ptr(ru), _ptr(ro), _ptr(x_out), _ptr(n_rows),
                                      _ptr(times), _ptr(states), _stream()), "sem_ssa_simulate")
        return dict(x=x_out, n_rows=n_rows, times=times, states=states, _keep=[x0t, tht, ru, ro])


def daily_rows_of_grid(t):
    """pmcmc.py:68-74 (and :92-98, :107-113): per integer day d = 0 .. ceil(t[-1]) the reference keeps the LAST grid point
    with ceil(t) == d.  Returns (row_of_grid[n_grid] int32 with -1 = not kept, n_rows); raises ValueError like the
    reference's list.index when a day has no grid point."""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    days = np.ceil(t).astype(np.int64)
    n_rows = int(days[-1]) + 1
    row = np.full(t.size, -1, dtype=np.int32)
    for d in range(n_rows):
        hit = np.nonzero(days == d)[0]
        if hit.size == 0:
            raise ValueError(f"{d} is not in list")
        row[hit[-1]] = d
    return row, n_rows


def ode_daily(model, y0, theta, t, G=1, substeps=8, device=None):
    """Batched deterministic ODE synthesiser (sem_ode_daily; pmcmc.py:16-113): y0 (B,C) or (C,), theta (B,P) or (P,), time
    grid t.  Returns a CUDA float64 tensor (B, days+1, C): for every integer day the last grid point with ceil(t) == day."""
    L = _lib.load()
    dev = require_cuda(device)
    Cn, P, _ = model_dims(model, G)
    row, n_rows = daily_rows_of_grid(t)
    with torch.cuda.device(dev):
        y0t = _dev_f64(y0, dev); tht = _dev_f64(theta, dev)
        shared_y0, shared_th = y0t.dim() == 1, tht.dim() == 1
        B = 1 if (shared_y0 and shared_th) else (tht.shape[0] if shared_y0 else y0t.shape[0])
        tg = _dev_f64(t, dev).reshape(-1)
        rg = torch.from_numpy(row).to(dev)
        out = torch.empty((B, n_rows, Cn), dtype=torch.float64, device=dev)
        cfg = _lib.OdeConfig(model=model, n_groups=G, n_sets=B, n_grid=tg.numel(), n_rows=n_rows, substeps=int(substeps),
                             shared_y0=int(shared_y0), shared_theta=int(shared_th))
        _lib.check(L.sem_ode_daily(C.byref(cfg), _ptr(y0t.contiguous()), _ptr(tht.contiguous()), _ptr(tg), _ptr(rg), _ptr(out), _stream()),
                   "sem_ode_daily")
        out._keep = [y0t, tht, tg, rg]
        return out


def abc_trials(obs, n_trials, threshold, priors, seed=0, trial0=0, trial_ids=None, arith="fast32", early_reject=False,
               want_traj=False, replay=None, device=None):
    """Run n_trials ABC trials (sem_abc_run).  Returns dict(theta (n,2), distance (n,), traj (n,T,3)|None, n_events)."""
    L = _lib.load()
    dev = require_cuda(device)
    check_rates(priors, "the prior box")
    if replay is not None:
        check_rates(replay.get("theta"))
    with torch.cuda.device(dev):
        obsd = _dev_f64(obs, dev).reshape(-1, 3)
        T = obsd.shape[0]
        ids = None
        if trial_ids is not None:
            ids = torch.as_tensor(np.asarray(trial_ids, dtype=np.int64)).to(dev)
            n_trials = ids.numel()
        cfg = _lib.AbcConfig(n_days=T, arith=resolve_arith(None, arith), early_reject=int(bool(early_reject)),
                             n_trials=int(n_trials), trial0=int(trial0), threshold=float(threshold),
                             prior=(C.c_double * 4)(*[float(v) for v in priors]), seed=int(seed) & (2**64 - 1))
        theta = torch.empty((n_trials, 2), dtype=torch.float64, device=dev)
        dist = torch.empty((n_trials,), dtype=torch.float64, device=dev)
        traj = torch.empty((n_trials, T, 3), dtype=torch.int32, device=dev) if want_traj else None
        nev = torch.zeros((1,), dtype=torch.int64, device=dev)
        work = torch.zeros((1,), dtype=torch.int64, device=dev)
        th_in = ns_in = ru = ro = None
        if replay is not None:
            th_in = _dev_f64(replay["theta"], dev)
            ns_in = torch.from_numpy(np.ascontiguousarray(replay["n_start"], dtype=np.int64)).to(dev)
            ru = _dev_f64(replay["u"], dev)
            ro = torch.from_numpy(np.ascontiguousarray(replay["off"], dtype=np.int64)).to(dev)
        _lib.check(L.sem_abc_run(C.byref(cfg), _ptr(obsd), _ptr(ids), _ptr(th_in), _ptr(ns_in), _ptr(ru), _ptr(ro),
                                 _ptr(theta), _ptr(dist), _ptr(traj), _ptr(nev), _ptr(work), _stream()), "sem_abc_run")
        return dict(theta=theta, distance=dist, traj=traj, n_events=nev, _keep=[obsd, ids, th_in, ns_in, ru, ro, work])
