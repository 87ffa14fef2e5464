"""One particle filter sharded over several GPUs (BASELINE config 5).

Every rank owns a contiguous slice of the N particles.  Per observation step there is ONE exchange:
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
    """Run the sharded filter with one shard per torch.distributed rank.  Returns dict(log_zetas, shard, collapsed).
    shard_cls lets the CPU (gloo) test drive the exchange logic with a stand-in for the kernels."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
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


def run_local(Y, model, theta, n_global, world, *, devices=None, **kw):
    """Lock-step emulation of `world` ranks in one process (all shards on one device unless `devices` is given):
    the same kernels and host logic as run_distributed, with the collectives replaced by tensor copies."""
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
