"""CPU-only checks: the C-ABI library loads and exports every symbol include/sem_b200.h declares (no compute
calls without a GPU), the ctypes structs match the header's field lists, the product path fails loudly without a
CUDA device, and the host-side sharding logic (world_size 2 over gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

HEADER = os.path.join(ROOT, "include", "sem_b200.h")


def _built_lib():
    import sem_b200
    from sem_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        sys.path.insert(0, os.path.join(ROOT, "stochastic-epidemic-modelling_b200"))
        import importlib.util
        spec = importlib.util.spec_from_file_location("sem_build", os.path.join(ROOT, "stochastic-epidemic-modelling_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        mod.build()
    return _lib


def test_library_exports_every_declared_symbol():
    _lib = _built_lib()
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = set(re.findall(r"\b(sem_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 15
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/sem_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert _lib.load().sem_abi_version() == 2


def test_struct_layouts_match_header():
    _lib = _built_lib()
    src = open(HEADER).read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(u?int\d+_t|double|void|size_t)\s*\**", "", decl)
            for part in decl.split(","):
                names.append(re.sub(r"\[.*\]", "", part).replace("*", "").strip())
        return names

    for cname, cls in [("sem_pf_config", _lib.PfConfig), ("sem_pf_buffers", _lib.PfBuffers),
                       ("sem_sim_config", _lib.SimConfig), ("sem_abc_config", _lib.AbcConfig)]:
        assert fields(cname) == [f[0] for f in cls._fields_], cname
    assert C.sizeof(_lib.PfConfig) == 12 * 4 + 8 + 8 + 8 + 4 + 4 + 64
    assert C.sizeof(_lib.AbcConfig) == 16 + 8 + 8 + 8 + 32 + 8


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sem_b200
    with pytest.raises(sem_b200._lib.SemError):
        sem_b200.particle_filter(np.zeros((3, 3)), sem_b200.ModelType.SIR, np.array([2., 1.]))
    with pytest.raises(sem_b200._lib.SemError):
        sem_b200.sir_simulate([10, 1, 0], np.array([2., 1.]), 1, True)
    with pytest.raises(sem_b200._lib.SemError):
        sem_b200.abc_algo.abc_algo(np.ones((3, 3)), 1, 10., {"beta": [0, 5], "gamma": [0, 5]})


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import, include, load or execute it."""
    pkg = os.path.join(ROOT, "stochastic-epidemic-modelling_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle)|(#include.*oracle)|c_oracle|sem_oracle|libsem_oracle|oracle/", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not bad.search(open(os.path.join(dirpath, f)).read()), f
    assert not bad.search(open(os.path.join(ROOT, "sem_b200.py")).read())


def test_model_type_and_split():
    import sem_b200
    pm = sem_b200.pmcmc
    assert [m.value for m in pm.ModelType] == ["sir", "seir", "sir_subgroups", "sir_subgroups2"]
    assert pm._model_id(pm.ModelType.SIR) == 0 and pm._model_id(pm.ModelType.SEIR) == 1
    assert pm._model_id(pm.ModelType.SIR_SUBGROUPS2) == 3 and pm._model_id("sir") == 0
    assert pm._model_id("anything else") == 2                        # pmcmc.py:138-139 falls to the subgroup branch
    (beta, gamma), p2 = pm._split(2, 2, np.array([5., 2., 1., 3., .5, 1.7]), None)
    assert np.array_equal(beta, [[5, 2], [1, 3]]) and gamma == .5 and p2 == 1.0      # p_obs clipped to [0,1]
    th, p2 = pm._split(0, 1, np.array([2., 1.]), .1)
    assert np.array_equal(th, [2., 1.]) and p2 == .1


def test_shard_range_partitions():
    from sem_b200.abc_algo import shard_range
    for start, count, world in [(0, 10, 3), (100, 7, 8), (5, 64, 4), (0, 0, 2)]:
        got = []
        for r in range(world):
            lo, n = shard_range(start, count, r, world)
            got += list(range(lo, lo + n))
        assert got == list(range(start, start + count))


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import sem_b200
from oracle import c_oracle as co
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "abc_sir_small.npz"))
obs = g["observed"]
def run_batch(obs, n, thr, prior4, seed, trial0, arith, er):      # CPU executor standing in for the kernel
    o = co.abc_trials(obs, n, thr, prior4, arith=1, seed=seed, trial0=trial0, want_traj=False, n_threads=1)
    acc = np.nonzero(o["distance"] <= thr)[0]
    return acc + trial0, o["theta"][acc], o["distance"][acc], o["n_events"]
def run_traj(obs, ids, thr, prior4, seed, arith):
    return np.stack([co.abc_trials(obs, 1, thr, prior4, arith=1, seed=seed, trial0=int(i), n_threads=1)["traj"][0] for i in ids])
st = {}
theta, traj = sem_b200.abc_algo.abc_rejection(obs, 4, 45.0, {"beta": [0, 5], "gamma": [0, 5]}, seed=7, batch=600,
                                              run_batch=run_batch, run_traj=run_traj, stats=st)
np.savez(sys.argv[4] + ".%s.npz" % sys.argv[3], theta=theta, traj=traj, ids=st["accepted_ids"])
dist.destroy_process_group()
"""


def test_abc_sharding_world2_gloo(tmp_path, c_oracle):
    """Two ranks shard the trial ids; both end with the accepted set a single process finds."""
    port = str(29500 + os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    assert np.array_equal(r0["ids"], r1["ids"]) and np.array_equal(r0["theta"], r1["theta"]) and np.array_equal(r0["traj"], r1["traj"])
    obs = golden("abc_sir_small")["observed"]
    ref = c_oracle.abc_trials(obs, 6000, 45.0, (0, 5, 0, 5), arith=1, seed=7, trial0=0)
    acc = np.nonzero(ref["distance"] <= 45.0)[0][:4]
    assert np.array_equal(r0["ids"], acc)
    assert np.array_equal(r0["theta"], ref["theta"][acc])
    assert np.array_equal(r0["traj"][:, :, 1:], ref["traj"][acc].astype(float))


WORKER_NOSEED = WORKER.replace("seed=7, batch=600", "seed=None, batch=600").replace(
    'ids=st["accepted_ids"])', 'ids=st["accepted_ids"], seed=np.array([st["seed"]], dtype=np.uint64))').replace(
    "import sem_b200", "np.random.seed(1000 + int(sys.argv[3]))\nimport sem_b200")


def test_abc_sharding_world2_gloo_default_seed(tmp_path, c_oracle):
    """seed=None under torch.distributed: the ranks' numpy generators differ, rank 0's Philox seed is broadcast, so both
    ranks key the global trial ids with ONE seed and return the same accepted (theta, trajectory) set, which is the
    single-process answer for that seed."""
    port = str(33500 + os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER_NOSEED)
    out = str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    assert int(r0["seed"][0]) == int(r1["seed"][0])
    assert np.array_equal(r0["ids"], r1["ids"]) and np.array_equal(r0["theta"], r1["theta"]) and np.array_equal(r0["traj"], r1["traj"])
    obs = golden("abc_sir_small")["observed"]
    ref = c_oracle.abc_trials(obs, 6000, 45.0, (0, 5, 0, 5), arith=1, seed=int(r0["seed"][0]), trial0=0)
    acc = np.nonzero(ref["distance"] <= 45.0)[0][:4]
    assert np.array_equal(r0["ids"], acc) and np.array_equal(r0["theta"], ref["theta"][acc])
    assert np.array_equal(r0["traj"][:, :, 1:], ref["traj"][acc].astype(float))


PATH_WORKER = r"""
import os, sys, types, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import sem_b200
from sem_b200 import sharded
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
g = np.load(sys.argv[4] + ".in.npz")
X, A = g["X"], g["A"]                                   # (T,C,N) states, (T,N) GLOBAL parent indices
N = X.shape[2]
lo, cnt = sharded.shard_bounds(N, 2)[rank]
sh = types.SimpleNamespace(n_global=N, X_hist=torch.from_numpy(X[:, :, lo:lo + cnt].copy()), ancestry=torch.from_numpy(A[:, lo:lo + cnt].copy()))
out = {}
for exact in (False, True):
    for fid in (0, 3, 11):
        out["t_%d_%d" % (exact, fid)] = sharded.path_sample_distributed(sh, 0xBEEF1234, fid, exact=exact)
np.savez(sys.argv[4] + ".%d.npz" % rank, **out)
dist.destroy_process_group()
"""


def test_path_sampler_over_shards_world2_gloo(tmp_path, c_oracle):
    """particle_path_sampler over a sharded history (pmcmc.py:236-248 globalised): the lineage chase with one broadcast
    per time equals the oracle's path sample on the concatenated history, with the reference's off-by-one indexing and
    with the exact genealogy; both ranks return the same trajectory; the final particle is the kernels' Philox pick."""
    from sem_b200 import sharded
    rng = np.random.RandomState(4)
    T, C, N = 9, 3, 37
    X = rng.randint(0, 1000, (T, C, N)).astype(np.int32)
    A = rng.randint(0, N, (T, N)).astype(np.int32)
    A[0] = 0
    out = str(tmp_path / "res")
    np.savez(out + ".in.npz", X=X, A=A)
    port = str(35500 + os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(PATH_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    for exact in (False, True):
        for fid in (0, 3, 11):
            k = "t_%d_%d" % (exact, fid)
            assert np.array_equal(r0[k], r1[k])
            chosen = sharded.path_pick(0xBEEF1234, fid, N)
            u, _ = c_oracle.philox_uniform_pair(0xBEEF1234, 0, 0, 0, 4, fid)
            assert chosen == min(int(u * N), N - 1)
            ref = c_oracle.path_sample(np.ascontiguousarray(np.transpose(X, (0, 2, 1))), A, chosen, exact=exact)
            assert np.array_equal(r0[k], ref.astype(float)), (exact, fid)


# ------------------------------------------------------------------ sharded filter: host-side logic
def test_sharded_host_arithmetic(c_oracle):
    from sem_b200 import sharded
    # host Philox == oracle Philox; the step uniform is the first word pair mapped to [0,1)
    for ctr, key in [((0, 0, 0, 0), (0, 0)), ((7, 1, 2, 3), (5, 9)), ((0, 0, 9, (2 << 24) | 5), (0x1234, 0xC0FFEE))]:
        assert list(sharded.philox4x32_10(ctr, key)) == c_oracle.philox4x32(list(ctr), list(key))
    u_ref, _ = c_oracle.philox_uniform_pair(0xC0FFEE00001234, 0, 0, 9, 2, 5)
    assert sharded.step_uniform(0xC0FFEE00001234, 9, 5) == u_ref
    # J(c) against brute force, and the exchange plan partitions every slot exactly once
    rng = np.random.RandomState(0)
    for n, world in [(10, 2), (101, 3), (64, 8), (7, 4)]:
        w = rng.gamma(.5, size=n) * (rng.random_sample(n) > .3)
        w[rng.randint(n)] += 1.0
        u0 = float(rng.random_sample())
        cdf = np.cumsum(w); total = float(cdf[-1])
        v = ((np.arange(n) + u0) / float(n)) * total
        for c in list(cdf[:-1]) + [0.0, total, total * .5]:
            brute = int(np.argmax(v >= c)) if np.any(v >= c) else n
            assert sharded.first_slot_ge(float(c), u0, n, total) == (n if c >= total else brute)
        bounds = sharded.shard_bounds(n, world)
        assert sum(c for _, c in bounds) == n and bounds[0][0] == 0
        G = np.array([0.0] + [float(cdf[lo + cnt - 1]) for lo, cnt in bounds])
        slot, counts = sharded.exchange_plan(G, total, u0, n, world)
        assert slot[0] == 0 and slot[-1] == n and all(a <= b for a, b in zip(slot, slot[1:]))
        assert np.array_equal(counts.sum(0), [c for _, c in bounds])             # every rank receives exactly its slots
        assert np.array_equal(counts.sum(1), np.diff(slot))                      # every rank sends exactly its children
        # ancestors implied by the plan == searchsorted on the global cdf
        anc = np.searchsorted(cdf, v, side="right")
        for r, (lo, cnt) in enumerate(bounds):
            mine = anc[slot[r]:slot[r + 1]]
            assert np.all((mine >= lo) & (mine < lo + cnt))
    comb = sharded.combine_summaries(np.array([[-3.0, 2.0], [-1.0, 4.0], [-np.inf, 0.0]]))
    assert comb["ok"] and comb["M"] == -1.0 and np.allclose(comb["G"], [0, 2 * np.exp(-2.0), 2 * np.exp(-2.0) + 4, 2 * np.exp(-2.0) + 4])
    assert not sharded.combine_summaries(np.array([[-np.inf, 0.0], [-np.inf, 0.0]]))["ok"]


SHARD_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import sem_b200
from sem_b200 import sharded
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)

class FakeShard:
    '''CPU stand-in for the three shard kernels: state = (id, generation, weight key); propagation is a
    deterministic map; weights depend on the state.  Exercises the host logic + collectives of run_distributed.'''
    def __init__(self, rank, world, model, Y, theta, n_global, **kw):
        self.rank, self.world, self.n_global = rank, world, n_global
        self.j0, self.n_local = sharded.shard_bounds(n_global, world)[rank]
        self.dev = torch.device("cpu"); self.seed, self.filter_id, self.Cn = 99, 0, 3
        self.x = None; self.hist = []; self.anc = []
    def _summary(self):
        lw = -((self.x[:, 0] * 7 + self.x[:, 1] * 3) % 11) / 3.0
        lw[self.x[:, 0] % 5 == 0] = -np.inf
        self.lw = lw
        M = lw.max(); self.e = np.where(np.isfinite(lw), np.exp(lw - M), 0.0)
        return torch.tensor([M, self.e.sum()], dtype=torch.float64)
    def init(self):
        j = np.arange(self.j0, self.j0 + self.n_local)
        self.x = np.stack([j, 0 * j, (j * 13) % 17], 1).astype(np.int64)
        self.hist.append(self.x.copy()); self.anc.append(np.zeros(self.n_local, np.int64))
        return self._summary()
    def offspring(self, p, comb, summ, u0, slot):
        r = self.rank
        cdf = comb["G"][r] + comb["s"][r] * np.cumsum(self.e)
        cdf[-1] = comb["G"][r + 1]
        lo = np.concatenate([[comb["G"][r]], cdf[:-1]])
        recs = []
        for i in range(self.n_local):
            a, b = sharded.first_slot_ge(float(lo[i]), u0, self.n_global, comb["total"]), sharded.first_slot_ge(float(cdf[i]), u0, self.n_global, comb["total"])
            for _ in range(a, b):
                recs.append(list(self.x[i]) + [self.j0 + i])
        assert len(recs) == slot[r + 1] - slot[r]
        return torch.tensor(np.array(recs, dtype=np.int32).reshape(-1, 4))
    def propagate(self, p, comb, summ, u0, slot, recv):
        rec = recv.numpy().astype(np.int64)
        self.x = np.stack([rec[:, 0], rec[:, 1] + 1, (rec[:, 2] * 5 + p + np.arange(self.j0, self.j0 + self.n_local)) % 17], 1)
        self.hist.append(self.x.copy()); self.anc.append(rec[:, 3].copy())
        return self._summary()

Y = np.zeros((6, 3))
out = sharded.run_distributed(Y, 0, np.zeros(2), 53, shard_cls=FakeShard)
sh = out["shard"]
np.savez(sys.argv[4] + ".%d.npz" % rank, logz=out["log_zetas"], hist=np.array(sh.hist), anc=np.array(sh.anc))
dist.destroy_process_group()
"""


def test_sharded_exchange_world2_gloo(tmp_path):
    """run_distributed over gloo (2 ranks) == the same toy filter run unsharded in numpy: global systematic
    resampling, all-gather of summaries and all-to-all-v particle migration."""
    from sem_b200 import sharded
    port = str(31500 + os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(SHARD_WORKER)
    out = str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    assert np.array_equal(r0["logz"], r1["logz"])
    hist = np.concatenate([r0["hist"], r1["hist"]], axis=1); anc = np.concatenate([r0["anc"], r1["anc"]], axis=1)
    # unsharded numpy reference of the same toy filter
    n, T = 53, 6
    j = np.arange(n)
    x = np.stack([j, 0 * j, (j * 13) % 17], 1).astype(np.int64)
    logz = np.zeros(T)
    for p in range(1, T):
        assert np.array_equal(hist[p - 1], x)
        lw = -((x[:, 0] * 7 + x[:, 1] * 3) % 11) / 3.0
        lw[x[:, 0] % 5 == 0] = -np.inf
        M = lw.max(); e = np.where(np.isfinite(lw), np.exp(lw - M), 0.0)
        logz[p] = logz[p - 1] + M + np.log(e.sum()) - np.log(n)
        u0 = sharded.step_uniform(99, p, 0)
        cdf = np.cumsum(e)
        a = np.minimum(np.searchsorted(cdf, ((j + u0) / float(n)) * cdf[-1], side="right"), n - 1)
        assert np.array_equal(anc[p], a)
        x = np.stack([x[a, 0], x[a, 1] + 1, (x[a, 2] * 5 + p + j) % 17], 1)
    assert np.array_equal(hist[T - 1], x)
    np.testing.assert_allclose(r0["logz"], logz, rtol=1e-12)


# ------------------------------------------------------------------ "next" rows N2 / N3 (host utilities)
def test_helpers_match_reference_formulas(tmp_path):
    from sem_b200 import helpers, results_io
    rng = np.random.RandomState(0)
    chains = [rng.normal(size=(400, 3)) + .1 * i for i in range(3)]
    # helpers.py:15-43 restated literally
    M, N = 3, 400
    means = np.array([c.mean(0) for c in chains]); var = np.array([((c - c.mean(0)) ** 2).sum(0) / (N - 1) for c in chains])
    W = var.mean(0); B = N / (M - 1) * ((means - means.mean(0)) ** 2).sum(0); V = (N - 1) / N * W + (M + 1) / (M * N) * B
    np.testing.assert_allclose(helpers.gelman_rubin_test(chains), np.sqrt(V / W), rtol=1e-13)
    x = rng.normal(size=2000)
    cs = np.cumsum(np.insert(x, 0, 0))
    assert np.array_equal(helpers.running_mean(x, 7), (cs[7:] - cs[:-7]) / 7.0)                       # helpers.py:46-48
    assert helpers.posterior_mse(np.array([.1, .2, .3]), chains[0]) == np.mean((chains[0] - np.array([.1, .2, .3])) ** 2)
    lo, hi = helpers.hdi(x, .95)
    assert abs(lo + 1.96) < .15 and abs(hi - 1.96) < .15 and np.mean((x >= lo) & (x <= hi)) >= .95 - 1e-9
    m, lo2, hi2 = helpers.mean_credible_interval(x)
    assert m == x.mean() and (lo2, hi2) == (lo, hi)
    ar = np.zeros(5000); e = rng.normal(size=5000)
    for i in range(1, 5000):
        ar[i] = .9 * ar[i - 1] + e[i]
    assert 150 < helpers.effective_sample_size(ar) < 450 and helpers.effective_sample_size(e) > 3500        # (1-.9)/(1+.9)*5000 = 263
    th = np.repeat(rng.normal(size=(50, 2)), 4, axis=0)
    assert abs(helpers.acceptance_rate(th) - 49 / 199) < 1e-12
    # results I/O + warm start (tests/experiments/noise/noise_.1.py:18-26,45-56)
    thetas = np.repeat(rng.normal(size=(60, 2)) + 2, 5, axis=0); lik = rng.random_sample(300); trajs = rng.randint(0, 500, (15, 300, 3)).astype(float)
    d = str(tmp_path / "run1")
    results_io.save_results(d, thetas, lik, trajs)
    assert sorted(os.listdir(d)) == ["likelihoods.csv", "sampled_trajs_infected.csv", "sampled_trajs_recovered.csv",
                                     "sampled_trajs_susceptible.csv", "thetas.csv"]
    t2, l2, tr2 = results_io.load_results(d)
    assert np.array_equal(t2, thetas) and np.array_equal(l2, lik) and np.array_equal(tr2, trajs)
    theta0, sigma = results_io.warm_start(d)
    uniq = np.unique(thetas[100:][::20], axis=0)
    assert theta0 == thetas[-1].tolist() and np.array_equal(sigma, np.cov(uniq.T, ddof=0))


def test_hot_loops_fit_the_instruction_cache():
    """Pins the size of the event loops of the two headline kernels (<= 256 SASS instructions = 4 KB): a build whose
    direct-method block grew to 264 instructions ran the filter 23 % slower ('Instruction cache /
    code layout').  Parsed from the built library with cuobjdump; skipped when cuobjdump is not installed."""
    import re, shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    import sem_b200
    sem_b200._lib.load()
    sass = subprocess.run([cuobjdump, "-sass", sem_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    for arith, min_wide in ((3, 30), (4, 15)):                  # fast32: two Philox calls per block; uniformized32: one per group
        name = f"_ZN3sem13pf_persistentINS_8SirModelELi{arith}EEEvNS_5PfDevE"
        body = sass[sass.index("Function : " + name):]
        body = body[:body.index("Function : ", 20)] if "Function : " in body[20:] else body
        ins = {int(m.group(1), 16): m.group(2) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+([^;]+);", body)}
        sizes = []
        for a, t in ins.items():
            m = re.search(r"BRA (?:P\d, )?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loop = [ins[x] for x in range(int(m.group(1), 16), a + 16, 16) if x in ins]
                if sum("IMAD.WIDE" in i for i in loop) >= min_wide:
                    sizes.append(len(loop))
        assert sizes, f"event loop of arith {arith} not found"
        assert min(sizes) <= 256, (arith, min(sizes))


def test_auto_interval_simulation_choice():
    """arith='auto': uniformized32 for SIR / SEIR filters with slow dynamics (or unknown theta), the direct method with
    32-bit streams for fast growth, subgroup models, ABC and logged simulations; explicit names pass through."""
    from sem_b200 import engine
    A = engine.ARITH
    assert engine.resolve_arith(0, "auto") == A["uniformized32"] and engine.resolve_arith(1, "auto") == A["uniformized32"]
    assert engine.resolve_arith(0, "auto", theta=[.4, .2]) == A["uniformized32"]          # the BASELINE workload
    assert engine.resolve_arith(0, "auto", theta=[2.0, 1.0]) == A["fast32"]               # growth 1 per interval
    assert engine.resolve_arith(0, "auto", theta=[2.0, 1.0], n_particles=1000) == A["fast32"]       # ... and a latency-bound filter
    assert engine.resolve_arith(0, "auto", theta=[2.0, 1.0], n_particles=10**5) == A["uniformized32"]   # a full GPU: throughput decides
    assert engine.resolve_arith(1, "auto", theta=[.4, .1, .1]) == A["uniformized32"]
    assert engine.resolve_arith(1, "auto", theta=[4.0, 1.0, 1.0]) == A["fast32"]
    assert engine.resolve_arith(2, "auto") == engine.resolve_arith(3, "auto", theta=[5, 2, 1, 3, .5]) == A["fast32"]
    assert engine.resolve_arith(None, "auto") == A["fast32"]
    assert engine.resolve_arith(0, "fast") == 1 and engine.resolve_arith(0, 2) == 2
    cfg = engine.make_pf_config(0, 100, 5, mu=[20], n_population=[1000])
    assert cfg.arith == A["uniformized32"]


def test_negative_rates_are_refused_on_the_host():
    """A negative rate constant is refused before anything is launched, with numpy's own words for it
    (np.random.choice inside gillespie_algo.py:63); device-resident parameters are the caller's job."""
    import torch
    from sem_b200 import engine
    engine.check_rates([.4, .2]); engine.check_rates(None); engine.check_rates(torch.tensor([-1.0]))
    for bad in ([-.4, .2], np.array([[.4, .2], [.4, -1e-9]])):
        with pytest.raises(ValueError, match="not non-negative"):
            engine.check_rates(bad)
    # the drop-in simulators: a negative count next to a live infection is the reference's ValueError (gillespie_algo.py:63;
    # tests/golden/negative_s0_reference.npz holds the reference's own message), without one the loop never runs
    import sem_b200
    chk = sem_b200.gillespie_algo._check_propensities
    chk(0, 1, [10, 5, 0], [1.2, .4]); chk(0, 1, [-1, 0, 0], [1.2, .4]); chk(2, 2, [[3, 1, 0], [2, 5, 0]], [2.7, 1.7, 1.8, 1.2, .9])
    for model, G, x0, th in [(0, 1, [-1, 5, 0], [1.2, .4]), (1, 1, [10, -1, 5, 0], [1, 1, 1]), (2, 2, [[3, 0, 0], [-1, 5, 0]], [2.7, 1.7, 1.8, 1.2, .9])]:
        with pytest.raises(ValueError, match="not non-negative"):
            chk(model, G, x0, th)
    assert "non-negative" in str(golden("negative_s0_reference")["simulator_error"])


def test_lookahead_mh_has_the_law_of_the_sequential_loop(monkeypatch):
    """Host logic of particle_mcmc(lookahead=L) with the filter replaced by a stub that returns an exact
    log-density: the speculative batches (proposals drawn under the rejection assumption, first acceptance ends the batch,
    later evaluations discarded, adaptive covariance from running sums) must leave the chain's stationary law and
    acceptance rate unchanged -- an off-by-one in the bookkeeping would bias both."""
    import sem_b200
    from sem_b200 import engine, pmcmc
    mu, sd = np.array([2.0, 1.0]), np.array([.3, .2])
    T = 5
    launches = {"n": 0}

    class StubIteration:
        def __init__(self, cfg, Y, device=None):
            self.F = cfg.n_filters
            self.launches = 1
            self.buf = np.zeros((self.F, engine.ITER_HEADER + T * 3))

        def run(self, theta, filter_id, probs=None, arith=None):
            th = np.atleast_2d(np.asarray(theta, dtype=float))
            launches["n"] += 1
            self.buf[:, 0] = -0.5 * (((th - mu) / sd) ** 2).sum(1)
            self.buf[:, 1] = 0
            self.buf[:, engine.ITER_HEADER:] = np.repeat(th[:, :1], T * 3, axis=1)     # "trajectory" = beta of the evaluated proposal
            return self.buf

    class StubLib:
        def sem_pf_launch_count(self, cfg):
            return 1

    monkeypatch.setattr(engine, "PreparedIteration", StubIteration)
    monkeypatch.setattr(engine, "require_cuda", lambda device=None: "cpu")
    monkeypatch.setattr(engine._lib, "load", lambda: StubLib())
    Y = np.zeros((T, 3))
    n = 40000
    out = {}
    for L in (1, 8):
        np.random.seed(12)
        launches["n"] = 0
        st = {}
        th, lik, traj = pmcmc.particle_mcmc(Y, pmcmc.ModelType.SIR, [2.0, 1.0], .12, adaptive=(L == 8), n_chains=n, probs=.1, n_particles=100,
                                            n_population=1000, mu=20, seed=1, lookahead=L, return_log=True, stats=st)
        assert np.all(th > 0)
        np.testing.assert_allclose(lik, -0.5 * (((th - mu) / sd) ** 2).sum(1), rtol=1e-12, atol=1e-12)   # stored likelihood belongs to the stored theta
        assert np.array_equal(traj[0, :, 0], th[:, 0])                                                 # ... and so does the stored trajectory
        acc = np.mean(np.any(th[1:] != th[:-1], axis=1))
        out[L] = (th[2000:].mean(0), th[2000:].std(0), acc, launches["n"])
        ess = n / 12.0
        assert np.all(np.abs(th[2000:].mean(0) - mu) < 5 * sd / np.sqrt(ess)), out[L]
        assert np.all(np.abs(th[2000:].std(0) / sd - 1) < .08), out[L]
    assert out[1][3] > 0.95 * n and out[8][3] < 0.9 * n               # (negative proposals need no launch; the adaptive chain accepts ~80 %: 1.2 iterations per launch)
    assert 0.15 < out[1][2] < 0.6
