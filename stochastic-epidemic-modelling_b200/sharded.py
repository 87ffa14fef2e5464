"""One particle filter sharded over several GPUs (BASELINE config 5).

Every rank owns a contiguous slice of the N particles; resampling is global and systematic (one shared uniform).

Two implementations of the per-step exchange:

* DEVICE-SIDE (PeerFilter, default whenever a rank's slice is co-resident on its GPU): every rank makes ONE cooperative
  launch per filter pass (sem_pf_run_sharded); the resampling barrier and the particle migration run inside the kernels
  over peer memory (NVLink): each CTA stores its 16-byte weight partial into every rank's arena, each parent stores one
  record per child into the arena of the rank that owns the child's slot, readers poll for the data itself.  No host
  work, no collective call and no kernel launch per step.  The arenas are cudaMalloc'ed by the library and mapped into
  the peers with CUDA IPC (one handle per rank, exchanged once through torch.distributed).
* HOST-DRIVEN (Shard / run_distributed(exchange="host")): for slices too large to be co-resident (config 5 at full
  size, where a step is ~1 s of compute).  Per step:
  1. all-gather of the per-shard weight summaries (max log-weight M_r, sum exp(logw - M_r))      [2 doubles / rank]
  2. every rank derives the same global quantities on the host: M, the exclusive scan G_r of the rescaled totals,
     the global total, the likelihood increment and -- from the single systematic uniform u0 of the step -- the slot
     range [J(G_r), J(G_r+1)) whose ancestors live on rank r  (global systematic resampling, one shared uniform)
  3. kernel sem_shard_offspring writes one record (state, global ancestor index) per child, ordered by slot
  4. all-to-all-v of the records to the ranks that own the slots (particle migration; int32, C+1 words per particle)
  5. kernel sem_shard_propagate: exact SSA over the interval, store, weigh, local scan, new summary
Philox streams are keyed by the GLOBAL particle index, so the sharded filter reproduces the single-GPU filter of
the same size and seed (up to ulp-level ties in the cdf).

`Comm` hides the collectives: TorchComm = torch.distributed (NCCL on GPUs, gloo in the CPU tests);
run_local() drives W shards in lock-step inside one process (used to test the full path on a single GPU).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib, engine


# ------------------------------------------------------------------------------------------ host-side arithmetic
def philox4x32_10(ctr, key):
    """Philox4x32-10 on the host (same function as csrc/sem_common.cuh) -- used for the step's shared uniform."""
    c0, c1, c2, c3 = [int(v) & 0xFFFFFFFF for v in ctr]
    k0, k1 = [int(v) & 0xFFFFFFFF for v in key]
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def step_uniform(seed, p, filter_id):
    """u0 of step p: Philox(k=0, item=0, step=p, domain 2 | filter id) -> 52-bit uniform, as pf_step draws it."""
    w = philox4x32_10((0, 0, p, (2 << 24) | (filter_id & 0xFFFFFF)), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return (((w[1] << 32) | w[0]) >> 12) * 2.0 ** -52


def shard_bounds(n, world):
    """Contiguous slices of n particles: [(start, count)] per rank."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        out.append((lo, cnt)); lo += cnt
    return out


def first_slot_ge(c, u0, n, total):
    """J(c) = smallest slot j in [0,n] with v_j >= c, v_j = ((j+u0)/n)*total  (same fp64 operations as the kernel)."""
    if c >= total:
        return n
    nd = float(n)
    v = lambda j: ((float(j) + u0) / nd) * total
    g = math.ceil((c / total) * nd - u0)
    j = 0 if g < 0 else (n if g > n else int(g))
    while j > 0 and v(j - 1) >= c:
        j -= 1
    while j < n and v(j) < c:
        j += 1
    return j


def combine_summaries(summ):
    """summ [W,2] = (M_r, total_r).  Returns dict(M, s[W], G[W+1], total, ok): global max, rescale factors
    exp(M_r - M), exclusive scan of s_r*total_r (G[W] = total), and ok = not collapsed."""
    summ = np.asarray(summ, dtype=np.float64)
    M = float(np.max(summ[:, 0]))
    if not np.isfinite(M):
        return dict(ok=False)
    s = np.where(np.isfinite(summ[:, 0]), np.exp(summ[:, 0] - M), 0.0)
    G = np.zeros(len(summ) + 1)
    for r in range(len(summ)):
        G[r + 1] = G[r] + s[r] * summ[r, 1]
    total = float(G[-1])
    return dict(ok=bool(total > 0 and np.isfinite(total)), M=M, s=s, G=G, total=total)


def exchange_plan(G, total, u0, n_global, world):
    """Slot ranges produced by every rank and the all-to-all-v split sizes (in records).
    Returns (slot_lo[W+1], send_counts[W][W]) with send_counts[src][dst]."""
    slot = [first_slot_ge(float(G[r]), u0, n_global, total) for r in range(world)] + [n_global]
    slot[0] = 0
    for r in range(1, world + 1):           # monotone by construction; guard against ulp reversals
        slot[r] = max(slot[r], slot[r - 1])
    bounds = shard_bounds(n_global, world)
    counts = np.zeros((world, world), dtype=np.int64)
    for src in range(world):
        for dst, (lo, cnt) in enumerate(bounds):
            counts[src, dst] = max(0, min(slot[src + 1], lo + cnt) - max(slot[src], lo))
    return slot, counts


# ------------------------------------------------------------------------------------------ one shard on one device
class Shard:
    def __init__(self, rank, world, model, Y, theta, n_global, G=1, observations=False, probs=.1, arith="auto", seed=0,
                 filter_id=0, mu=None, n_population=None, store_history=True, device=None):
        self.rank, self.world, self.n_global = rank, world, n_global
        self.j0, self.n_local = shard_bounds(n_global, world)[rank]
        self.dev = engine.require_cuda(device)
        self.seed, self.filter_id = int(seed), int(filter_id)
        self.Cn = engine.model_dims(model, G)[0]
        self.cfg = engine.make_pf_config(model, self.n_local, len(Y), G=G, observations=observations, probs=probs,
                                         resampler="systematic", arith=arith, seed=seed, filter_id0=filter_id, mu=mu,
                                         n_population=n_population, store_history=store_history)
        with torch.cuda.device(self.dev):
            self.Y = torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float64)).to(self.dev)
            self.theta = torch.from_numpy(np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)).to(self.dev)
            self.out = engine.alloc_pf_outputs(self.cfg, self.dev)
            self.summary = torch.zeros(2, dtype=torch.float64, device=self.dev)
        X_hist, anc, logz, status, nev, ws = self.out
        self.buf = _lib.PfBuffers(Y=engine._ptr(self.Y), theta=engine._ptr(self.theta), X0=None, X_hist=engine._ptr(X_hist),
                                  ancestry=engine._ptr(anc), log_zetas=engine._ptr(logz), status=engine._ptr(status),
                                  n_events=engine._ptr(nev), workspace=engine._ptr(ws))
        self.launches = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def init(self):
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().sem_shard_init(C.byref(self.cfg), C.byref(self.buf), self.j0, engine._ptr(self.summary),
                                                  self._stream()), "sem_shard_init")
        self.launches += 1
        return self.summary

    def _step_struct(self, p, comb, summ, u0, slot):
        r = self.rank
        return _lib.ShardStep(step=p, particle_offset=self.j0, n_global=self.n_global, u0=u0, total=comb["total"],
                              total_local=float(summ[r, 1]), G=float(comb["G"][r]), G_next=float(comb["G"][r + 1]),
                              s=float(comb["s"][r]), slot0=int(slot[r]))

    def offspring(self, p, comb, summ, u0, slot):
        n_children = int(slot[self.rank + 1] - slot[self.rank])
        with torch.cuda.device(self.dev):
            send = torch.empty((max(n_children, 1), self.Cn + 1), dtype=torch.int32, device=self.dev)
            st = self._step_struct(p, comb, summ, u0, slot)
            _lib.check(_lib.load().sem_shard_offspring(C.byref(self.cfg), C.byref(self.buf), C.byref(st), engine._ptr(send),
                                                       self._stream()), "sem_shard_offspring")
        self.launches += 1
        return send[:n_children]

    def propagate(self, p, comb, summ, u0, slot, recv):
        assert recv.shape == (self.n_local, self.Cn + 1) and recv.is_contiguous()
        with torch.cuda.device(self.dev):
            st = self._step_struct(p, comb, summ, u0, slot)
            _lib.check(_lib.load().sem_shard_propagate(C.byref(self.cfg), C.byref(self.buf), C.byref(st), engine._ptr(recv),
                                                       engine._ptr(self.summary), self._stream()), "sem_shard_propagate")
        self.launches += 1
        self._keep = recv
        return self.summary

    # results of this shard
    @property
    def X_hist(self):       # [rows][C][n_local] int32
        return self.out[0][0]

    @property
    def ancestry(self):     # [rows][n_local] int32, GLOBAL parent indices
        return self.out[1][0]

    @property
    def n_events(self):
        return int(self.out[4].cpu()[0])


# ------------------------------------------------------------------------------------------ drivers
def _advance(log_zetas, p, comb, n_global):
    log_zetas[p] = log_zetas[p - 1] + comb["M"] + math.log(comb["total"]) - math.log(n_global)   # pmcmc.py:183, log domain


def run_distributed(Y, model, theta, n_global, *, group=None, shard_cls=None, **kw):
    """Run the sharded filter (host-driven exchange) with one shard per torch.distributed rank.  Returns
    dict(log_zetas, shard, collapsed).  seed=None draws a fresh Philox seed on rank 0 and broadcasts it.
    shard_cls lets the CPU (gloo) test drive the exchange logic with a stand-in for the kernels."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if kw.get("seed") is None:
        box = [engine.new_seed()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        kw["seed"] = box[0]
    sh = (shard_cls or Shard)(rank, world, model, Y, theta, n_global, **kw)
    T = len(Y)
    log_zetas = np.zeros(T)
    dev = sh.dev

    def gather(summary):
        out = torch.empty((world, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, summary.reshape(1, 2), group=group)
        return out.cpu().numpy()

    summ = gather(sh.init())
    for p in range(1, T):
        comb = combine_summaries(summ)
        if not comb["ok"]:
            return dict(log_zetas=None, shard=sh, collapsed=p)
        _advance(log_zetas, p, comb, n_global)
        u0 = step_uniform(sh.seed, p, sh.filter_id)
        slot, counts = exchange_plan(comb["G"], comb["total"], u0, n_global, world)
        send = sh.offspring(p, comb, summ, u0, slot)
        recv = torch.empty((sh.n_local, sh.Cn + 1), dtype=torch.int32, device=dev)
        w = sh.Cn + 1
        dist.all_to_all_single(recv.view(-1), send.reshape(-1), [int(c) * w for c in counts[:, rank]],
                               [int(c) * w for c in counts[rank, :]], group=group)
        summary = sh.propagate(p, comb, summ, u0, slot, recv)
        if p < T - 1:
            summ = gather(summary)
    return dict(log_zetas=log_zetas, shard=sh, collapsed=0)


def path_pick(seed, filter_id, n_global):
    """Final particle of the path sample: Philox(domain 4 | filter id) -> randint(0, N) (pmcmc.py:241), as the kernels draw it."""
    w = philox4x32_10((0, 0, 0, (4 << 24) | (filter_id & 0xFFFFFF)), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = (((w[1] << 32) | w[0]) >> 12) * 2.0 ** -52
    return min(int(u * float(n_global)), n_global - 1)


def path_sample_distributed(sh, seed, filter_id, exact=False, group=None):
    """particle_path_sampler (pmcmc.py:236-248) over the shards of the host-driven filter: the lineage is chased
    backwards through the GLOBAL ancestor indices; at every time the rank that owns the current particle broadcasts its
    state row and parent index (T small broadcasts; a step of these filters is ~1 s of compute).  Every rank returns the
    same (T, C) trajectory.  exact=False keeps the reference's indexing of ancestry row p."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = shard_bounds(sh.n_global, world)
    X, A = sh.X_hist, sh.ancestry
    T, Cn = X.shape[0], X.shape[1]
    traj = np.zeros((T, Cn))
    idx = path_pick(int(seed), int(filter_id), sh.n_global)
    for cur in range(T - 1, -1, -1):
        owner = next(r for r, (lo, cnt) in enumerate(bounds) if lo <= idx < lo + cnt)
        buf = torch.zeros(Cn + 1, dtype=torch.float64, device=X.device)
        if rank == owner:
            loc = idx - bounds[rank][0]
            buf[:Cn] = X[cur, :, loc].to(torch.float64)
            if cur > 0:
                buf[Cn] = A[cur if exact else cur - 1, loc].to(torch.float64)
        dist.broadcast(buf, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        v = buf.cpu().numpy()
        traj[cur] = v[:Cn]
        idx = int(v[Cn])
    return traj


def run_local(Y, model, theta, n_global, world, *, devices=None, **kw):
    """Lock-step emulation of `world` ranks in one process (all shards on one device unless `devices` is given):
    the same kernels and host logic as run_distributed, with the collectives replaced by tensor copies."""
    if kw.get("seed") is None:
        kw["seed"] = engine.new_seed()
    shards = [Shard(r, world, model, Y, theta, n_global, device=None if devices is None else devices[r], **kw)
              for r in range(world)]
    T = len(Y)
    log_zetas = np.zeros(T)
    summ = np.stack([sh.init().cpu().numpy() for sh in shards])
    for p in range(1, T):
        comb = combine_summaries(summ)
        if not comb["ok"]:
            return dict(log_zetas=None, shards=shards, collapsed=p)
        _advance(log_zetas, p, comb, n_global)
        u0 = step_uniform(shards[0].seed, p, shards[0].filter_id)
        slot, counts = exchange_plan(comb["G"], comb["total"], u0, n_global, world)
        sends = [sh.offspring(p, comb, summ, u0, slot) for sh in shards]
        new = []
        for dst, sh in enumerate(shards):
            parts = []
            for src in range(world):
                off = int(counts[src, :dst].sum())
                parts.append(sends[src][off:off + int(counts[src, dst])].to(sh.dev))
            recv = torch.cat(parts).contiguous()
            new.append(sh.propagate(p, comb, summ, u0, slot, recv).cpu().numpy())
        if p < T - 1:
            summ = np.stack(new)
    return dict(log_zetas=log_zetas, shards=shards, collapsed=0)


# ------------------------------------------------------------------------------------------ device-side exchange
class _DevMem:
    """Raw device memory as a __cuda_array_interface__ object (zero-copy torch view of a library-owned arena)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3)


class PeerFilter:
    """One rank's shard of a filter whose exchange runs on the device (include/sem_b200.h: sem_pf_run_sharded).

    Life cycle: construct on every rank with the same arguments -> exchange `handle` (64 bytes) -> connect(...) ->
    [host barrier] -> run(...) any number of times (one launch each; no host synchronisation between ranks) ->
    after a run whose status is not 0: reset() on every rank + host barrier.
    """

    def __init__(self, rank, world, model, Y, n_local, *, G=1, observations=False, probs=.1, arith="auto", theta=None,
                 seed=0, mu=None, n_population=None, store_history=True, device=None, block_particles=0, timeout_s=20.0,
                 path_exact=False, want_handle=True):
        self.L = _lib.load()
        self.rank, self.world, self.n_local, self.n_global = int(rank), int(world), int(n_local), int(n_local) * int(world)
        self.dev = engine.require_cuda(device)
        self.model, self.G = model, G
        self.Cn, self.P, _ = engine.model_dims(model, G)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        self.T = Y.shape[0]
        self.cfg = engine.make_pf_config(model, self.n_local, self.T, G=G, observations=observations, probs=probs,
                                         resampler="systematic", arith=engine.resolve_arith(model, arith, theta=theta, n_particles=self.n_local),
                                         seed=seed, mu=mu, n_population=n_population, store_history=store_history,
                                         block_particles=block_particles, path_exact=path_exact)
        with torch.cuda.device(self.dev):
            if not self.L.sem_pf_sharded_supported(C.byref(self.cfg), self.world):
                raise _lib.SemError("device-side exchange not available: " + self.L.sem_last_error().decode())
            self.Y = torch.from_numpy(Y).to(self.dev)
            self.out = engine.alloc_pf_outputs(self.cfg, self.dev)
            self.pin_th = torch.empty((self.P,), dtype=torch.float64).pin_memory()
            self.dev_th = torch.empty((self.P,), dtype=torch.float64, device=self.dev)
            self.bytes = self.L.sem_xchg_bytes(C.byref(self.cfg), self.world)
            arena = C.c_void_p()
            hbuf = (C.c_ubyte * 64)()
            _lib.check(self.L.sem_xchg_alloc(self.bytes, C.byref(arena), hbuf if want_handle else None), "sem_xchg_alloc")
            self.arena = arena.value
            self.handle = bytes(hbuf)
            self.desc = _lib.XchgDesc(world=self.world, rank=self.rank, generation=0, launch_tag=0, timeout_s=float(timeout_s))
            self.desc.arena[self.rank] = self.arena
            self._opened = []
            n_it = engine.ITER_HEADER + self.T * self.Cn
            self.iter_ptr = self.L.sem_xchg_iteration_result(C.byref(self.cfg), self.world, self.arena)
            self.dev_it = torch.as_tensor(_DevMem(self.iter_ptr, (n_it,), "<f8"), device=self.dev)
            self.pin_it = torch.empty((n_it,), dtype=torch.float64).pin_memory()
            self.reset()
        X_hist, anc, logz, status, nev, ws = self.out
        self.buf = _lib.PfBuffers(Y=engine._ptr(self.Y), theta=engine._ptr(self.dev_th), X0=None, X_hist=engine._ptr(X_hist),
                                  ancestry=engine._ptr(anc), log_zetas=engine._ptr(logz), status=engine._ptr(status),
                                  n_events=engine._ptr(nev), workspace=engine._ptr(ws))
        self.launches = 0

    # -- wiring
    def connect(self, arenas):
        """arenas[r] = rank r's arena as a device pointer valid in THIS process (own entry may be None)."""
        for r in range(self.world):
            if r != self.rank:
                self.desc.arena[r] = int(arenas[r])

    def open_handles(self, handles):
        """Map the peers' arenas from their CUDA IPC handles (other processes)."""
        arenas = [None] * self.world
        with torch.cuda.device(self.dev):
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                ptr = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(self.L.sem_xchg_open(hb, C.byref(ptr)), "sem_xchg_open")
                self._opened.append(ptr.value)
                arenas[r] = ptr.value
        self.connect(arenas)

    def reset(self):
        """Empty marks in the own arena; generation restarts.  All ranks must reset, then barrier, before the next run."""
        with torch.cuda.device(self.dev):
            _lib.check(self.L.sem_xchg_reset(C.byref(self.cfg), self.world, self.arena, engine._stream()), "sem_xchg_reset")
            torch.cuda.current_stream(self.dev).synchronize()
        self.desc.generation = 0

    def close(self):
        with torch.cuda.device(self.dev):
            torch.cuda.synchronize(self.dev)
            for ptr in self._opened:
                self.L.sem_xchg_close(ptr)
            self._opened = []
            if self.arena:
                self.L.sem_xchg_free(self.arena)
                self.arena = None

    # -- one filter pass
    def run(self, theta, filter_id=0, probs=None, arith=None, want_path=False, stream=None):
        """Enqueue this rank's launch of one filter pass (all T steps).  Returns nothing; read results after a stream
        synchronisation (log_zetas, status, X_hist, ancestry properties) or use iteration()."""
        cfg = self.cfg
        cfg.filter_id0 = int(filter_id) & 0xFFFFFF
        if probs is not None:
            cfg.probs = float(probs)
        if arith is not None:
            cfg.arith = int(arith)
        self.pin_th.numpy()[...] = np.asarray(theta, dtype=np.float64).reshape(-1)
        stream = stream or torch.cuda.current_stream(self.dev)
        self.buf.iteration_result = self.iter_ptr if want_path else None
        with torch.cuda.device(self.dev):
            if want_path:                                    # theta H2D + launch + result D2H in one C call
                _lib.check(self.L.sem_pf_iteration_sharded(C.byref(cfg), C.byref(self.buf), C.byref(self.desc), C.c_void_p(self.pin_th.data_ptr()),
                                                           C.c_void_p(self.pin_it.data_ptr()), C.c_void_p(stream.cuda_stream)),
                           "sem_pf_iteration_sharded")
            else:
                with torch.cuda.stream(stream):
                    self.dev_th.copy_(self.pin_th, non_blocking=True)
                    _lib.check(self.L.sem_pf_run_sharded(C.byref(cfg), C.byref(self.buf), C.byref(self.desc),
                                                         C.c_void_p(stream.cuda_stream)), "sem_pf_run_sharded")
        self.launches += 1

    def iteration(self, theta, filter_id=0, probs=None, arith=None):
        """One PMCMC iteration's worth: filter pass + path sample over all shards; returns the pinned packed result
        [log_zetas[T-1], status, this rank's n_events, chosen global particle, trajectory[T][C]] (valid until the next run)."""
        self.run(theta, filter_id, probs=probs, arith=arith, want_path=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.pin_it.numpy()

    # -- results of this shard (after synchronisation)
    @property
    def X_hist(self):       # [rows][C][n_local] int32
        return self.out[0][0]

    @property
    def ancestry(self):     # [rows][n_local] int32, GLOBAL parent indices
        return self.out[1][0]

    @property
    def log_zetas(self):
        return self.out[2][0]

    @property
    def status(self):
        return int(self.out[3].cpu()[0])

    @property
    def n_events(self):
        return int(self.out[4].cpu()[0])


_PEER_CACHE = {}


def cached_peer_filter(group, rank, world, model, Y, n_local, *, G=1, observations=False, probs=.1, arith="auto", theta=None, seed=0,
                       mu=None, n_population=None, path_exact=False):
    """A connected PeerFilter for this configuration, kept across calls (particle_mcmc(sharded=True) called repeatedly does
    not pay cudaMalloc + CUDA IPC mapping + a barrier each time).  Every rank must call with the same arguments."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    key = (id(group), rank, world, model, Y.shape, int(n_local), G, bool(observations), bool(path_exact),
           tuple(np.atleast_1d(mu).tolist()), tuple(np.atleast_1d(n_population).tolist()), torch.cuda.current_device())
    pf = _PEER_CACHE.get(key)
    if pf is None:
        pf = PeerFilter(rank, world, model, Y, n_local, G=G, observations=observations, probs=probs, arith=arith, theta=theta,
                        seed=seed, mu=mu, n_population=n_population, path_exact=path_exact)
        connect_distributed(pf, group)
        _PEER_CACHE[key] = pf
    else:
        pf.Y.copy_(torch.from_numpy(Y))
        pf.cfg.seed = int(seed) & (2**64 - 1)
        pf.cfg.probs = float(probs)
        pf.cfg.arith = engine.resolve_arith(model, arith, theta=theta, n_particles=n_local)
    return pf


def _close_cached():
    for pf in list(_PEER_CACHE.values()):
        try:
            pf.close()
        except Exception:
            pass
    _PEER_CACHE.clear()


import atexit  # noqa: E402
atexit.register(_close_cached)


def connect_distributed(pf, group=None):
    """Exchange the CUDA IPC handles of the arenas over torch.distributed (once) and map the peers."""
    import torch.distributed as dist
    handles = [None] * pf.world
    dist.all_gather_object(handles, pf.handle, group=group)
    pf.open_handles(handles)
    dist.barrier(group=group)           # every arena is reset and mapped before anyone launches


def connect_local(pfs):
    """Shards living in ONE process (one or several devices): plain pointers, peer access between the devices."""
    L = _lib.load()
    for a in pfs:
        for b in pfs:
            if a.dev != b.dev:
                _lib.check(L.sem_peer_enable(a.dev.index, b.dev.index), "sem_peer_enable")
    for a in pfs:
        a.connect([b.arena for b in pfs])


def device_exchange_supported(model, n_local, T, world, **kw):
    """True when a rank's slice can run the device-side exchange (co-resident CTAs, supported arithmetic)."""
    cfg = engine.make_pf_config(model, n_local, T, resampler="systematic", **kw)
    return bool(_lib.load().sem_pf_sharded_supported(C.byref(cfg), int(world)))


def run_peer_distributed(Y, model, theta, n_global, *, group=None, filter_id=0, want_path=False, pf=None, **kw):
    """One pass of the sharded filter with the device-side exchange, one shard per torch.distributed rank.
    Returns dict(log_zetas, shard, collapsed[, iteration]).  Pass `pf` to reuse a connected PeerFilter."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if n_global % world:
        raise ValueError("the device-side exchange needs n_global divisible by the number of ranks")
    if pf is None:
        pf = PeerFilter(rank, world, model, Y, n_global // world, theta=theta, **kw)
        connect_distributed(pf, group)
    pf.run(theta, filter_id, want_path=want_path)
    torch.cuda.current_stream(pf.dev).synchronize()
    st = pf.status
    out = dict(log_zetas=None if st else pf.log_zetas.cpu().numpy(), shard=pf, collapsed=st)
    if want_path:
        out["iteration"] = pf.pin_it.numpy()
    if st != 0:                         # leave the arenas clean for the next pass
        pf.reset()
        dist.barrier(group=group)
    return out


def run_peer_local(Y, model, theta, n_global, world, *, devices=None, filter_id=0, want_path=False, passes=1, **kw):
    """All W ranks of the device-side exchange inside one process: one stream (and kernel) per rank, running
    CONCURRENTLY on one device (small shards; used by the single-GPU tests) or on several devices."""
    if n_global % world:
        raise ValueError("n_global must be divisible by world")
    pfs = [PeerFilter(r, world, model, Y, n_global // world, theta=theta, want_handle=False,
                      device=None if devices is None else devices[r], **kw) for r in range(world)]
    try:
        connect_local(pfs)
        streams = [torch.cuda.Stream(device=pf.dev) for pf in pfs]
        outs = []
        for k in range(passes):
            for pf, st in zip(pfs, streams):
                pf.run(theta, filter_id + k, want_path=want_path, stream=st)
            for st in streams:
                st.synchronize()
            status = [pf.status for pf in pfs]
            outs.append(dict(status=status, log_zetas=[pf.log_zetas.cpu().numpy() for pf in pfs],
                             X_hist=torch.cat([pf.X_hist.to(pfs[0].dev) for pf in pfs], dim=2).cpu().numpy(),
                             ancestry=torch.cat([pf.ancestry.to(pfs[0].dev) for pf in pfs], dim=1).cpu().numpy(),
                             n_events=sum(pf.n_events for pf in pfs),
                             iteration=[pf.pin_it.numpy().copy() for pf in pfs] if want_path else None))
            if any(status):
                for pf in pfs:
                    pf.reset()
        return outs
    finally:
        for pf in pfs:
            pf.close()
