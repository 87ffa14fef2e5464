"""ctypes front-end to the C oracle (oracle/sem_oracle.c).  TEST INFRASTRUCTURE ONLY.

Builds oracle/libsem_oracle.so on demand (make -C oracle).  Used by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsem_oracle.so")

MODEL_IDS = {"SIR": 0, "SEIR": 1, "SIR_SUBGROUPS": 2, "SIR_SUBGROUPS2": 3}
DOM_SSA, DOM_RESAMPLE, DOM_INIT, DOM_PATH, DOM_ABC_PRIOR, DOM_ABC_SSA, DOM_SIM = 1, 2, 3, 4, 5, 6, 7


def build(force=False):
    src = os.path.join(HERE, "sem_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libsem_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class PfCfg(C.Structure):
    _fields_ = [("model", C.c_int32), ("obs_kind", C.c_int32), ("resampler", C.c_int32), ("arith", C.c_int32),
                ("philox", C.c_int32), ("N", C.c_int32), ("T", C.c_int32), ("G", C.c_int32), ("Cobs", C.c_int32),
                ("init_poisson", C.c_int32), ("n_threads", C.c_int32), ("pad", C.c_int32),
                ("probs", C.c_double), ("dt", C.c_double), ("seed", C.c_uint64), ("filter_id", C.c_uint32),
                ("pad2", C.c_uint32)]


class AbcCfg(C.Structure):
    _fields_ = [("philox", C.c_int32), ("T", C.c_int32), ("arith", C.c_int32), ("early_reject", C.c_int32),
                ("n_threads", C.c_int32), ("pad", C.c_int32), ("threshold", C.c_double), ("prior", C.c_double * 4),
                ("seed", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.so_binom_logpmf.restype = C.c_double
        L.so_binom_logpmf.argtypes = [C.c_double] * 3
        L.so_norm_logpdf.restype = C.c_double
        L.so_norm_logpdf.argtypes = [C.c_double] * 3
        L.so_poisson_philox.restype = C.c_double
        L.so_poisson_philox.argtypes = [C.c_double, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.so_poisson_u_philox.restype = C.c_double
        L.so_poisson_u_philox.argtypes = [C.c_double, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.so_ssa_replay.restype = C.c_int64
        L.so_ssa_philox.restype = C.c_int64
        L.so_pf_run.restype = C.c_int
        L.so_abc_trials.restype = C.c_int
        L.so_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def n_cols(model, G):
    return 3 if model == 0 else 4 if model == 1 else 3 * G


def philox4x32(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().so_philox4x32(c, k, o)
    return [int(x) for x in o]


def philox_uniform_pair(seed, k, c1, c2, domain, fid=0):
    u1, u2 = C.c_double(), C.c_double()
    lib().so_philox_uniform_pair(C.c_uint64(seed), C.c_uint32(k), C.c_uint32(c1), C.c_uint32(c2),
                                 C.c_uint32(domain), C.c_uint32(fid), C.byref(u1), C.byref(u2))
    return u1.value, u2.value


def binom_logpmf(k, n, p):
    f = lib().so_binom_logpmf
    return np.array([f(float(a), float(b), float(c)) for a, b, c in zip(*np.broadcast_arrays(k, n, p))])


def norm_logpdf(y, x, probs):
    f = lib().so_norm_logpdf
    return np.array([f(float(a), float(b), float(c)) for a, b, c in zip(*np.broadcast_arrays(y, x, probs))])


def poisson_philox(mu, seed, c1, c2, domain=DOM_INIT, fid=0):
    return lib().so_poisson_philox(float(mu), C.c_uint64(seed), C.c_uint32(c1), C.c_uint32(c2),
                                   C.c_uint32(domain), C.c_uint32(fid))


def poisson_u_philox(mu, seed, c1, c2, domain=8, fid=0):
    """Candidate count of a uniformized32 batch (sem_oracle.c poisson_draw_u; DOM_AUX stream)."""
    return lib().so_poisson_u_philox(float(mu), C.c_uint64(seed), C.c_uint32(c1), C.c_uint32(c2),
                                     C.c_uint32(domain), C.c_uint32(fid))


def ssa(model, G, x0, theta, max_time, arith=0, u=None, seed=None, sim_index=0, max_rec=0):
    """One SSA run. Replay (u given) or philox (seed). Returns dict(x, pairs, times, states)."""
    Cn = n_cols(model, G)
    x = np.array(x0, dtype=np.float64).reshape(-1).copy()
    assert x.size == Cn
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    times = np.zeros(max(max_rec, 1)); states = np.zeros((max(max_rec, 1), Cn))
    nrec = C.c_int64(0); over = C.c_int(0)
    if u is not None:
        u = np.ascontiguousarray(u, dtype=np.float64)
        pairs = lib().so_ssa_replay(C.c_int(model), C.c_int(G), _p(x), _p(theta), C.c_double(max_time), C.c_int(arith),
                                    _p(u), C.c_int64(u.size), _p(times), _p(states), C.c_int64(max_rec),
                                    C.byref(nrec), C.byref(over))
        if over.value:
            raise RuntimeError("replay buffer exhausted")
    else:
        # no event log asked for: null buffers, so that the uniformized orders run as such (they have no event times)
        pairs = lib().so_ssa_philox(C.c_int(model), C.c_int(G), _p(x), _p(theta), C.c_double(max_time), C.c_int(arith),
                                    C.c_uint64(seed), C.c_uint32(sim_index), _p(times if max_rec > 0 else None),
                                    _p(states if max_rec > 0 else None), C.c_int64(max_rec), C.byref(nrec))
    n = min(nrec.value, max_rec)
    return dict(x=x, pairs=int(pairs), n_rec=int(nrec.value), times=times[:n], states=states[:n])


def pf_run(model, Y, theta, observations, probs, N, *, G=1, resampler=0, arith=0, dt=1.0,
           X0=None, mu=None, npop=None, flat_u=None, res_u=None, ssa_u=None, ssa_off=None,
           seed=None, filter_id=0, n_threads=0, want_logw=False):
    """Particle filter on the CPU oracle.

    Replay: flat_u (sequential, reference order) or (res_u, ssa_u, ssa_off).  Philox: seed (+X0 or mu/npop).
    Returns dict with X_hist (T,N,C) int32, ancestry (T,N) int32, log_zetas (T,), collapsed (0 or step),
    and for flat replay the per-particle consumption map (res_u, ssa_off into flat_u).
    """
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    T, Cobs = Y.shape
    Cn = n_cols(model, G)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    philox = seed is not None
    cfg = PfCfg(model=model, obs_kind=int(bool(observations)), resampler=resampler, arith=arith, philox=int(philox),
                N=N, T=T, G=G, Cobs=Cobs, init_poisson=int(X0 is None), n_threads=n_threads, probs=float(probs),
                dt=float(dt), seed=int(seed or 0), filter_id=filter_id)
    if X0 is not None:
        X0 = np.ascontiguousarray(np.asarray(X0).reshape(N, Cn), dtype=np.int32)
    if mu is not None:
        mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float64).reshape(-1))
        npop = np.ascontiguousarray(np.asarray(npop, dtype=np.float64).reshape(-1))
    X_hist = np.zeros((T, N, Cn), dtype=np.int32)
    anc = np.zeros((T, N), dtype=np.int32)
    logz = np.zeros(T)
    logw = np.zeros((T, N)) if want_logw else None
    nev = C.c_int64(0); used = C.c_int64(0)
    res_u_out = ssa_off_out = None
    if flat_u is not None:
        flat_u = np.ascontiguousarray(flat_u, dtype=np.float64)
        res_u_out = np.zeros((T - 1, N)); ssa_off_out = np.zeros((T - 1) * N + 1, dtype=np.int64)
    if res_u is not None:
        res_u = np.ascontiguousarray(res_u, dtype=np.float64)
        ssa_u = np.ascontiguousarray(ssa_u, dtype=np.float64)
        ssa_off = np.ascontiguousarray(ssa_off, dtype=np.int64)
    rc = lib().so_pf_run(C.byref(cfg), _p(Y), _p(theta), _p(X0), _p(mu), _p(npop),
                         _p(flat_u), C.c_int64(0 if flat_u is None else flat_u.size), C.byref(used),
                         _p(res_u), _p(ssa_u), _p(ssa_off), _p(res_u_out), _p(ssa_off_out),
                         _p(X_hist), _p(anc), _p(logz), _p(logw), C.byref(nev))
    if rc == -1:
        raise RuntimeError("replay buffer exhausted")
    out = dict(X_hist=X_hist, ancestry=anc, log_zetas=logz, collapsed=rc, n_events=int(nev.value))
    if want_logw:
        out["log_w"] = logw
    if flat_u is not None:
        # in flat mode ssa_off_out holds the cursor AFTER each particle; the cursor before particle (p,0) is
        # the end of step p's N resample uniforms.
        ends = ssa_off_out[1:].reshape(T - 1, N)
        starts = np.empty_like(ends)
        starts[:, 1:] = ends[:, :-1]
        starts[0, 0] = N
        if T > 2:
            starts[1:, 0] = ends[:-1, -1] + N
        out.update(res_u=res_u_out, ssa_start=starts, ssa_end=ends, flat_used=int(used.value))
    return out


def flat_to_csr(flat_u, ssa_start, ssa_end):
    """Cut the flat reference-order stream into per-(step,particle) CSR buffers."""
    Tm1, N = ssa_start.shape
    lens = (ssa_end - ssa_start).reshape(-1)
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    out = np.empty(off[-1])
    s = ssa_start.reshape(-1); e = ssa_end.reshape(-1)
    for q in range(lens.size):
        out[off[q]:off[q + 1]] = flat_u[s[q]:e[q]]
    return out, off


def path_sample(X_hist, ancestry, chosen, exact=False):
    T, N, Cn = X_hist.shape
    X_hist = np.ascontiguousarray(X_hist, dtype=np.int32); ancestry = np.ascontiguousarray(ancestry, dtype=np.int32)
    traj = np.zeros((T, Cn), dtype=np.int32)
    lib().so_path_sample(_p(X_hist), _p(ancestry), C.c_int(T), C.c_int(N), C.c_int(Cn), C.c_int(chosen),
                         C.c_int(int(exact)), _p(traj))
    return traj


def abc_trials(obs, n_trials, threshold, priors=(0, 5, 0, 5), *, arith=0, seed=None, trial0=0, theta=None,
               n_start=None, ssa_u=None, ssa_off=None, early_reject=False, want_traj=True, n_threads=0):
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    T = obs.shape[0]
    cfg = AbcCfg(philox=int(seed is not None), T=T, arith=arith, early_reject=int(early_reject), n_threads=n_threads,
                 threshold=float(threshold), prior=(C.c_double * 4)(*[float(v) for v in priors]), seed=int(seed or 0))
    if theta is not None:
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n_start = np.ascontiguousarray(n_start, dtype=np.int64)
        ssa_u = np.ascontiguousarray(ssa_u, dtype=np.float64)
        ssa_off = np.ascontiguousarray(ssa_off, dtype=np.int64)
    th_out = np.zeros((n_trials, 2)); dist = np.zeros(n_trials)
    traj = np.zeros((n_trials, T, 3), dtype=np.int32) if want_traj else None
    nev = C.c_int64(0)
    rc = lib().so_abc_trials(C.byref(cfg), _p(obs), C.c_int64(n_trials), C.c_uint64(trial0), _p(theta), _p(n_start),
                             _p(ssa_u), _p(ssa_off), _p(th_out), _p(dist), _p(traj), C.byref(nev))
    if rc:
        raise RuntimeError("replay buffer exhausted")
    return dict(theta=th_out, distance=dist, traj=traj, n_events=int(nev.value))
