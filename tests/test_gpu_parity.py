"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libsem_b200.so), against the oracle and
the reference's golden vectors.  Bit-exact for integer states / ancestor indices; fp64 log-likelihoods to 1e-9
relative (north_star tolerance), in practice ~1e-13."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden, golden_names, mt_doubles

pytestmark = pytest.mark.gpu

MODEL_OF = {"sir": 0, "seir": 1, "subgroups": 2}
MODEL_IDS = {"SIR": 0, "SEIR": 1, "SIR_SUBGROUPS": 2, "SIR_SUBGROUPS2": 3}


@pytest.fixture(scope="module")
def sem():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sem_b200
    sem_b200._lib.load()
    return sem_b200


# ------------------------------------------------------------------ small device functions
def test_philox_known_answers(sem, c_oracle):
    L = sem._lib.load()
    for ctr, key in [([0, 0, 0, 0], [0, 0]), ([0xffffffff] * 4, [0xffffffff] * 2),
                     ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]), ([7, 1, 2, 3], [5, 9])]:
        out = (C.c_uint32 * 4)()
        assert L.sem_test_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out) == 0
        assert [int(v) for v in out] == c_oracle.philox4x32(ctr, key)
    out = (C.c_uint32 * 4)()
    L.sem_test_philox((C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), out)
    assert [int(v) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]      # Random123 KAT


def _map3(fn, a, b, c):
    a, b, c = [np.ascontiguousarray(v, dtype=np.float64) for v in np.broadcast_arrays(a, b, c)]
    out = np.empty_like(a)
    p = lambda v: v.ctypes.data_as(C.c_void_p)
    assert fn(p(a), p(b), p(c), p(out), C.c_int64(a.size)) == 0
    return out


def test_weights_vs_scipy_golden(sem, c_oracle):
    L = sem._lib.load()
    g = golden("weights_known_answers")
    lp = _map3(L.sem_test_binom_logpmf, g["binom_k"], g["binom_n"], g["binom_p"])
    ref = g["binom_pmf"]
    zero = ref == 0
    assert np.all(np.exp(lp[zero]) == 0)
    np.testing.assert_allclose(np.exp(lp[~zero]), ref[~zero], rtol=4e-12)                   # scipy / Boost values
    lo = c_oracle.binom_logpmf(g["binom_k"], g["binom_n"], g["binom_p"])
    fin = np.isfinite(lo)
    assert np.array_equal(np.isfinite(lp), fin)
    np.testing.assert_allclose(lp[fin], lo[fin], rtol=1e-12, atol=2e-12)                    # C oracle (libm / IEEE division)
    ln = _map3(L.sem_test_norm_logpdf, g["norm_y"], g["norm_x"], g["norm_probs"])
    ok = g["norm_pdf"] > 1e-300
    np.testing.assert_allclose(ln[ok], np.log(g["norm_pdf"][ok]), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("mu", [0.0, 0.5, 3.0, 9.99, 10.0, 20.0, 4800.0, 8000.0])
def test_poisson_matches_oracle(sem, c_oracle, mu):
    L = sem._lib.load()
    n = 1500
    out = np.empty(n)
    assert L.sem_test_poisson(mu, 1234, 3, 7, out.ctypes.data_as(C.c_void_p), n) == 0
    ref = np.array([c_oracle.poisson_philox(mu, 1234, i, 7, 3) for i in range(n)])
    assert np.array_equal(out, ref)
    if mu > 0:
        assert abs(out.mean() - mu) < 5 * np.sqrt(mu / n) + 1e-9


def test_fast_math(sem):
    """-log(x) (table + polynomial) and the Newton reciprocal of the FAST arithmetic against libm / IEEE."""
    L = sem._lib.load()
    rng = np.random.RandomState(0)
    u = np.floor(rng.random_sample(200000) * 2.0**52) / 2.0**52                     # what the Philox mapping produces
    x = np.concatenate([1.0 - u, [1.0, 2.0**-52, 0.5, 0.6875, 0.687499999, 0.75, 1 - 2.0**-52], 2.0**-rng.randint(0, 53, 1000)])
    a = np.concatenate([10 ** rng.uniform(-6, 9, x.size - 3), [1.0, 3.0, 1e-300]])
    nl = np.empty_like(x); rc = np.empty_like(x)
    p = lambda v: v.ctypes.data_as(C.c_void_p)
    assert L.sem_test_fast_math(p(x), p(a), p(nl), p(rc), C.c_int64(x.size)) == 0
    ref = -np.log(x)
    err = np.abs(nl - ref)
    assert np.all(err <= 4.5e-16 * np.maximum(ref, 1.0)), (err / np.maximum(ref, 1.0)).max()   # <= 2 ulp of max(E, 1)
    assert abs(nl[x == 1.0][0]) < 1e-16
    np.testing.assert_allclose(rc[:-1], 1.0 / a[:-1], rtol=2.3e-16)                   # <= 1 ulp


# ------------------------------------------------------------------ SSA replay vs the reference (gillespie_algo.py)
@pytest.mark.parametrize("name", golden_names("ssa_"))
def test_ssa_replay_vs_reference(sem, name):
    g = golden(name)
    model = str(g["model"])
    n_ev = g["times"].size - 1
    u = mt_doubles(g["mt_key"], g["mt_pos"], 2 * (n_ev + 2))
    ga = sem.gillespie_algo
    tmax = float(g["max_time"])
    if model == "sir":
        out = ga.sir_simulate(list(g["population"]), g["theta"], tmax, False, replay_u=u)
        last = ga.sir_simulate(list(g["population"]), g["theta"], tmax, True, replay_u=u)
        cols = ["s", "i", "r"]
    elif model == "seir":
        out = ga.seir_simulate(list(g["population"]), g["theta"], tmax, False, replay_u=u)
        last = ga.seir_simulate(list(g["population"]), g["theta"], tmax, True, replay_u=u)
        cols = ["s", "e", "i", "r"]
    else:
        G = g["population"].shape[0]
        th = g["theta"]
        out = ga.sir_subgroups_simulate(g["population"], th[:G * G].reshape(G, G), th[-1], tmax, False, replay_u=u)
        last = ga.sir_subgroups_simulate(g["population"], th[:G * G].reshape(G, G), th[-1], tmax, True, replay_u=u)
        last = [v for grp in last for v in grp]
        cols = [f"{c}_{grp}" for grp in range(G) for c in ("s", "i", "r")]
    states = np.array([out[c] for c in cols], dtype=float).T
    assert np.array_equal(states, g["states"])                           # integer trajectory bit-exact
    np.testing.assert_allclose(out["time"], g["times"], rtol=1e-13, atol=0)   # event times to a few ulp
    assert list(last) == list(g["states"][-1].astype(int))


# ------------------------------------------------------------------ particle filter replay vs the reference
PF_CASES = [n for n in golden_names("pf_") if n != "pf_sir_collapse"]


def _replay_inputs(c_oracle, g):
    model = MODEL_IDS[str(g["model"])]
    G = g["mu"].size if model >= 2 else 1
    n = 4096
    while True:
        flat = mt_doubles(g["mt_key"], g["mt_pos"], n)
        try:
            o = c_oracle.pf_run(model, g["Y"], g["theta"], bool(g["observations"]), float(g["probs"]),
                                int(g["n_particles"]), G=G, X0=g["hidden_process"][0], flat_u=flat)
            break
        except RuntimeError:
            n *= 4
    ssa_u, off = c_oracle.flat_to_csr(flat, o["ssa_start"], o["ssa_end"])
    return dict(res_u=o["res_u"], ssa_u=ssa_u, ssa_off=off), o


def _theta_arg(g):
    model = MODEL_IDS[str(g["model"])]
    th = g["theta"]
    if model >= 2:
        G = g["mu"].size
        return (th[:G * G].reshape(G, G), float(th[-1]))
    return th


@pytest.mark.parametrize("name", PF_CASES)
@pytest.mark.parametrize("block", [0, 32])
def test_pf_replay_vs_reference(sem, c_oracle, name, block):
    g = golden(name)
    replay, _ = _replay_inputs(c_oracle, g)
    mt = getattr(sem.ModelType, str(g["model"]))
    model = MODEL_IDS[str(g["model"])]
    mu = g["mu"] if model >= 2 else float(g["mu"][0])
    npop = g["n_population"] if model >= 2 else int(g["n_population"][0])
    zetas, hidden, anc = sem.particle_filter(g["Y"], mt, _theta_arg(g), bool(g["observations"]), float(g["probs"]),
                                             int(g["n_particles"]), npop, mu, 1, X0=g["hidden_process"][0],
                                             replay=replay, block_particles=block)
    assert hidden.dtype == np.float64 and hidden.shape == g["hidden_process"].shape
    assert np.array_equal(hidden, g["hidden_process"])                   # SSA trajectories bit-exact
    assert np.array_equal(anc, g["ancestry_matrix"])                     # resampling indices bit-exact
    np.testing.assert_allclose(np.log(zetas), np.log(g["zetas"]), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(zetas, g["zetas"], rtol=1e-9)


def test_pf_collapse_returns_none(sem):
    g = golden("pf_sir_collapse")
    out = sem.particle_filter(g["Y"], sem.ModelType.SIR, g["theta"], False, float(g["probs"]), int(g["n_particles"]),
                              int(g["n_population"][0]), float(g["mu"][0]), 1, seed=3)
    assert out == (None, None, None)                                     # pmcmc.py:191-192


@pytest.mark.parametrize("name", golden_names("path_"))
def test_path_sampler_vs_reference(sem, name):
    g = golden(name)
    src = golden(str(g["source"]))
    tr = sem.particle_path_sampler(src["hidden_process"], src["ancestry_matrix"], chosen_path=int(g["chosen"]))
    assert np.array_equal(tr, g["trajectory"])


# ------------------------------------------------------------------ production (Philox) path vs the C oracle
def _truth_Y(model, T, seed, probs, normal, dying=False, G=2):
    rng = np.random.RandomState(seed)
    t = np.arange(T)
    if dying:       # sub-critical outbreak: I dies out, a few recoveries; many particles get weight exactly 0 (k > n)
        base = np.stack([290 - 0 * t, 0 * t, np.minimum(t, 2)], 1)
    elif model == 1:
        base = np.stack([900 - 40 * t, 10 * t, 20 + 15 * t, 15 * t], 1)       # E_0 = 0 (pmcmc.py:163)
    elif model == 2 and G == 3:
        base = np.stack([190 - 8 * t, 5 + 3 * t, 5 * t, 280 - 12 * t, 8 + 5 * t, 7 * t, 140 - 6 * t, 4 + 3 * t, 3 * t], 1)
    elif model == 2 and G == 4:
        base = np.stack([190 - 8 * t, 5 + 3 * t, 5 * t, 280 - 12 * t, 8 + 5 * t, 7 * t, 140 - 6 * t, 4 + 3 * t, 3 * t,
                         240 - 9 * t, 6 + 4 * t, 5 * t], 1)
    elif model == 2:
        base = np.stack([400 - 20 * t, 15 + 8 * t, 12 * t, 600 - 30 * t, 20 + 12 * t, 18 * t], 1)
    elif model == 3:
        base = np.stack([1000 - 50 * t, 35 + 20 * t, 30 * t], 1)
    else:
        base = np.stack([980 - 45 * t, 20 + 25 * t, 20 * t], 1)
    base = np.maximum(base, 0)
    if normal:
        return np.floor(base * (1 + probs * rng.normal(size=base.shape) * .3))
    return rng.binomial(base.astype(int), probs).astype(float)


PHILOX_CASES = [
    # model, G, theta, npop, mu, N, T, normal, probs, resampler, arith, block
    (0, 1, [2.0, 1.0], [1000], [20], 2000, 10, False, .1, 1, 1, 0),
    (0, 1, [2.0, 1.0], [1000], [20], 2000, 10, False, .1, 0, 1, 64),
    (0, 1, [2.0, 1.0], [1000], [20], 1999, 8, True, .1, 1, 0, 96),
    (0, 1, [0.5, 3.0], [300], [3], 777, 8, False, .5, 0, 0, 0),          # many extinctions
    (1, 1, [4.0, 1.0, 1.0], [1000], [20], 1500, 8, False, .1, 1, 1, 0),
    (1, 1, [4.0, 1.0, 1.0], [1000], [20], 1500, 8, True, .2, 0, 0, 128),
    (2, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1200, 7, False, .1, 1, 1, 0),
    (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1200, 7, False, .1, 0, 1, 32),
    (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1000, 7, True, .1, 1, 0, 0),
    (2, 3, [3, 1, .5, .2, 2, 1, 1, .7, 4, .8], [200, 300, 150], [5, 8, 4], 600, 6, False, .2, 1, 1, 0),
    # 32-bit streams (arith 3, the production default): two events per Philox call, same fp64 arithmetic
    (0, 1, [2.0, 1.0], [1000], [20], 2000, 10, False, .1, 1, 3, 0),
    (0, 1, [2.0, 1.0], [1000], [20], 1999, 8, True, .1, 0, 3, 96),
    (0, 1, [0.5, 3.0], [300], [3], 777, 8, False, .5, 0, 3, 0),          # many extinctions
    (1, 1, [4.0, 1.0, 1.0], [1000], [20], 1500, 8, False, .1, 1, 3, 0),
    (1, 1, [4.0, 1.0, 1.0], [1000], [20], 1500, 8, True, .2, 0, 3, 128),
    (2, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1200, 7, False, .1, 1, 3, 0),
    (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1200, 7, False, .1, 0, 3, 32),
    (2, 3, [3, 1, .5, .2, 2, 1, 1, .7, 4, .8], [200, 300, 150], [5, 8, 4], 600, 6, False, .2, 1, 3, 0),
    (2, 4, [3, 1, .5, .2, .3, 2, 1, 1, .7, .2, 4, .8, .1, .6, .9, 2.5, .8], [200, 300, 150, 250], [5, 8, 4, 6], 500, 6, False, .2, 1, 3, 0),
    (2, 4, [3, 1, .5, .2, .3, 2, 1, 1, .7, .2, 4, .8, .1, .6, .9, 2.5, .8], [200, 300, 150, 250], [5, 8, 4, 6], 500, 6, False, .2, 1, 1, 0),
    # uniformized intervals with 32-bit candidates and the drift-anticipating bound (arith 4)
    (0, 1, [2.0, 1.0], [1000], [20], 2000, 10, False, .1, 1, 4, 0),
    (0, 1, [2.0, 1.0], [1000], [20], 1999, 8, True, .1, 0, 4, 96),
    (0, 1, [0.5, 3.0], [300], [3], 777, 8, False, .5, 0, 4, 0),          # many extinctions
    (1, 1, [4.0, 1.0, 1.0], [1000], [20], 1500, 8, False, .1, 1, 4, 0),
    (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1200, 6, False, .1, 1, 4, 32),   # (T = 7 collapses at step 6 under this stream, on the oracle too)
    (2, 3, [3, 1, .5, .2, 2, 1, 1, .7, 4, .8], [200, 300, 150], [5, 8, 4], 600, 6, False, .2, 1, 4, 0),
    # uniformized intervals (arith 2): exact law without waiting times, bit-checked against the oracle's statement
    (0, 1, [2.0, 1.0], [1000], [20], 2000, 10, False, .1, 1, 2, 0),
    (0, 1, [2.0, 1.0], [1000], [20], 1999, 8, True, .1, 0, 2, 96),
    (0, 1, [0.5, 3.0], [300], [3], 777, 8, False, .5, 0, 2, 0),
    (1, 1, [4.0, 1.0, 1.0], [1000], [20], 1500, 8, False, .1, 1, 2, 0),
    (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1200, 7, False, .1, 1, 2, 0),
]


@pytest.mark.parametrize("case", PHILOX_CASES, ids=lambda c: f"m{c[0]}G{c[1]}N{c[5]}r{c[9]}a{c[10]}b{c[11]}{'n' if c[7] else 'b'}")
def test_pf_philox_vs_oracle(sem, c_oracle, case):
    import torch
    model, G, theta, npop, mu, N, T, normal, probs, resampler, arith, block = case
    Y = _truth_Y(model, T, 5, probs, normal, dying=(model == 0 and theta[1] > theta[0]), G=G)
    seed = 0xC0FFEE1234
    cfg = sem.engine.make_pf_config(model, N, T, G=G, observations=normal, probs=probs, resampler=resampler, arith=arith,
                                    seed=seed, filter_id0=3, mu=mu, n_population=npop, block_particles=block)
    res = sem.engine.run_pf(cfg, Y, np.array(theta, float))
    torch.cuda.synchronize()
    ref = c_oracle.pf_run(model, Y, theta, normal, probs, N, G=G, resampler=resampler, arith=arith, seed=seed,
                          filter_id=3, mu=mu, npop=npop)
    assert int(res.status.cpu()[0]) == ref["collapsed"] == 0
    X = res.X_hist[0].permute(0, 2, 1).cpu().numpy()
    assert np.array_equal(X[0], ref["X_hist"][0])                        # Poisson X_0
    assert np.array_equal(res.ancestry[0].cpu().numpy(), ref["ancestry"])
    assert np.array_equal(X, ref["X_hist"])
    np.testing.assert_allclose(res.log_zetas[0].cpu().numpy(), ref["log_zetas"], rtol=1e-11, atol=1e-11)
    assert int(res.n_events.cpu()[0]) == ref["n_events"]


def test_pf_fast_equals_reference_arithmetic(sem):
    """FAST and REF fp64 operation orders walk the same integer trajectories from the same Philox streams."""
    import torch
    Y = _truth_Y(0, 12, 5, .1, False)
    outs = []
    for arith in (0, 1):
        cfg = sem.engine.make_pf_config(0, 5000, 12, probs=.1, resampler=1, arith=arith, seed=99, mu=[20], n_population=[1000])
        r = sem.engine.run_pf(cfg, Y, np.array([2.0, 1.0]))
        torch.cuda.synchronize()
        outs.append((r.X_hist.cpu().numpy(), r.ancestry.cpu().numpy(), r.log_zetas.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-12)


def test_pf_multi_filter_batch(sem, c_oracle):
    """n_filters side by side == the same filters run one at a time (independent chains on one GPU)."""
    import torch
    Y = _truth_Y(0, 9, 5, .1, False)
    thetas = np.array([[2.0, 1.0], [1.5, 0.8], [2.5, 1.2]])
    cfg = sem.engine.make_pf_config(0, 900, 9, n_filters=3, probs=.1, resampler=1, arith=1, seed=5, filter_id0=10,
                                    mu=[20], n_population=[1000])
    res = sem.engine.run_pf(cfg, Y, thetas)
    torch.cuda.synchronize()
    for f in range(3):
        ref = c_oracle.pf_run(0, Y, thetas[f], False, .1, 900, resampler=1, arith=1, seed=5, filter_id=10 + f,
                              mu=[20], npop=[1000])
        assert np.array_equal(res.X_hist[f].permute(0, 2, 1).cpu().numpy(), ref["X_hist"])
        np.testing.assert_allclose(res.log_zetas[f].cpu().numpy(), ref["log_zetas"], rtol=1e-11)


@pytest.mark.parametrize("arith,normal", [(4, False), (3, False), (4, True)])
def test_batch_of_proposals_equals_single_filters(sem, c_oracle, arith, normal):
    """A batch of MH proposals in ONE launch (the offspring-form kernel with several filters side by side, each with its
    own theta AND observation parameter, sem_pf_iteration / probs_per_filter): every filter equals the oracle's run of
    that proposal alone -- states, ancestors, log-likelihoods, and the packed iteration result."""
    import torch
    T, N = 12, 1300
    Y = _truth_Y(0, T, 5, .1, normal)
    thetas = np.array([[.45, .2], [.4, .25], [.6, .2], [.5, .3], [.42, .18]])
    probs = np.array([.1, .12, .08, .15, .1])
    F = len(thetas)
    cfg = sem.engine.make_pf_config(0, N, T, n_filters=F, observations=normal, probs=.5, resampler=1, arith=arith, seed=77,
                                    mu=[20], n_population=[1000])
    assert sem._lib.load().sem_pf_launch_count(sem.engine.C.byref(cfg)) == 1
    prep = sem.engine.PreparedIteration(cfg, Y)
    r = prep.run(thetas, 40, probs=probs, arith=arith).copy()
    torch.cuda.synchronize()
    X_hist, anc, logz, status, nev, _ = prep.out
    for f in range(F):
        ref = c_oracle.pf_run(0, Y, thetas[f], normal, probs[f], N, resampler=1, arith=arith, seed=77, filter_id=40 + f, mu=[20], npop=[1000])
        assert int(status[f]) == ref["collapsed"] == 0
        assert np.array_equal(X_hist[f].permute(0, 2, 1).cpu().numpy(), ref["X_hist"]), f
        assert np.array_equal(anc[f].cpu().numpy(), ref["ancestry"]), f
        np.testing.assert_allclose(logz[f].cpu().numpy(), ref["log_zetas"], rtol=1e-11, atol=1e-11)
        assert r[f, 0] == float(logz[f, -1]) and r[f, 1] == 0 and int(r[f, 2]) == ref["n_events"]
        traj = r[f, sem.engine.ITER_HEADER:].reshape(T, 3)
        chosen = int(r[f, 3])
        assert np.array_equal(traj[-1], ref["X_hist"][-1, chosen])          # the path ends in the chosen particle ...
        k = chosen
        for p in range(T - 2, -1, -1):                                      # ... and follows the reference's ancestry indexing (D8)
            k = ref["ancestry"][p, k]
            assert np.array_equal(traj[p], ref["X_hist"][p, k])


# ------------------------------------------------------------------ headline size: size-independent properties
def test_pf_full_size_properties(sem, c_oracle):
    """BASELINE config 4 size (SIR, pop 1e4, 1e5 particles, 101 observations): invariants of every particle."""
    import torch
    N, T, pop = 100_000, 101, 10_000
    Y = bench_truth(T, pop)
    cfg = sem.engine.make_pf_config(0, N, T, probs=.1, resampler=1, arith=1, seed=2024, mu=[20], n_population=[pop])
    res = sem.engine.run_pf(cfg, Y, np.array([.4, .2]))
    torch.cuda.synchronize()
    assert int(res.status.cpu()[0]) == 0
    X = res.X_hist[0]                                                     # [T][3][N]
    assert bool((X.sum(1) == pop).all())                                  # S+I+R conserved (closed population)
    assert bool((X >= 0).all())
    A = res.ancestry[0].long()
    assert bool((A[1:, 1:] >= A[1:, :-1]).all())                          # systematic resampling => sorted ancestors
    assert bool(((A >= 0) & (A < N)).all())
    for p in (1, 50, 100):                                                # along every lineage S only falls, R only rises
        par = X[p - 1][:, A[p]]
        assert bool((X[p][0] <= par[0]).all()) and bool((X[p][2] >= par[2]).all())
    lz = res.log_zetas[0].cpu().numpy()
    assert np.all(np.isfinite(lz)) and np.all(np.diff(lz) < 0)
    # likelihood estimate agrees with the oracle's at a size it finishes in seconds (Monte-Carlo error only)
    ref = [c_oracle.pf_run(0, Y, [.4, .2], False, .1, 4000, resampler=1, arith=1, seed=s, mu=[20], npop=[pop])["log_zetas"][-1]
           for s in range(4)]
    assert abs(lz[-1] - np.mean(ref)) < 6 * (np.std(ref) + 0.05) + 0.5
    # and matches the oracle exactly on a 20k-particle filter of the same problem
    cfg2 = sem.engine.make_pf_config(0, 20_000, T, probs=.1, resampler=1, arith=1, seed=77, mu=[20], n_population=[pop])
    r2 = sem.engine.run_pf(cfg2, Y, np.array([.4, .2]))
    o2 = c_oracle.pf_run(0, Y, [.4, .2], False, .1, 20_000, resampler=1, arith=1, seed=77, mu=[20], npop=[pop])
    assert np.array_equal(r2.X_hist[0].permute(0, 2, 1).cpu().numpy(), o2["X_hist"])
    assert np.array_equal(r2.ancestry[0].cpu().numpy(), o2["ancestry"])
    np.testing.assert_allclose(r2.log_zetas[0].cpu().numpy(), o2["log_zetas"], rtol=1e-11)


def bench_truth(T, pop, beta=.4, gamma=.2, p=.1, seed=0):
    """Deterministic SIR ODE truth (RK4, one row per day) observed with binomial(p): the bench workload's Y."""
    y = np.array([pop - 20.0, 20.0, 0.0])
    rows = [y.copy()]
    f = lambda v: np.array([-beta * v[0] * v[1] / pop, beta * v[0] * v[1] / pop - gamma * v[1], gamma * v[1]])
    for _ in range(T - 1):
        for _ in range(20):
            h = 1 / 20
            k1 = f(y); k2 = f(y + h / 2 * k1); k3 = f(y + h / 2 * k2); k4 = f(y + h * k3)
            y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        rows.append(y.copy())
    truth = np.array(rows)
    return np.random.RandomState(seed).binomial(truth.astype(int), p).astype(float)


# ------------------------------------------------------------------ ABC (abc_algo.py)
def test_abc_replay_vs_reference(sem):
    from oracle import sem_oracle as so
    g = golden("abc_sir_small")
    obs = g["observed"]
    n = g["distance"].size
    streams, offs = [], [0]
    for i in range(n):
        st = so.MTStream(g["mt_key"][i], g["mt_pos"][i])
        so.abc_trial(obs, g["theta"][i, 0], g["theta"][i, 1], g["n_start"][i], st)
        streams.append(np.array(st.log)); offs.append(offs[-1] + len(st.log))
    out = sem.engine.abc_trials(obs, n, float(g["threshold"]), [0, 5, 0, 5], want_traj=True,
                                replay=dict(theta=g["theta"], n_start=g["n_start"], u=np.concatenate(streams), off=np.array(offs)))
    d = out["distance"].cpu().numpy()
    tr = out["traj"].cpu().numpy()
    np.testing.assert_allclose(d, g["distance"], rtol=1e-13)
    assert np.array_equal(tr[:, :, 1], g["I_sim"].astype(np.int32)) and np.array_equal(tr[:, :, 2], g["R_sim"].astype(np.int32))
    acc = d <= float(g["threshold"])
    assert np.array_equal(tr[acc], g["trajectories"][:, :, 1:].astype(np.int32))


@pytest.mark.parametrize("arith", [0, 1, 3])
def test_abc_philox_vs_oracle(sem, c_oracle, arith):
    g = golden("abc_sir_small")
    obs = g["observed"]
    n = 3000
    out = sem.engine.abc_trials(obs, n, 45.0, [0, 5, 0, 5], seed=42, trial0=1000, arith=arith, want_traj=True)
    ref = c_oracle.abc_trials(obs, n, 45.0, (0, 5, 0, 5), arith=arith, seed=42, trial0=1000)
    assert np.array_equal(out["theta"].cpu().numpy(), ref["theta"])
    assert np.array_equal(out["traj"].cpu().numpy(), ref["traj"])
    np.testing.assert_allclose(out["distance"].cpu().numpy(), ref["distance"], rtol=1e-13)
    assert int(out["n_events"].cpu()[0]) == ref["n_events"]
    # early rejection never changes the accepted set or its distances
    er = sem.engine.abc_trials(obs, n, 45.0, [0, 5, 0, 5], seed=42, trial0=1000, arith=arith, early_reject=True)
    d0, d1 = out["distance"].cpu().numpy(), er["distance"].cpu().numpy()
    assert np.array_equal(d0 <= 45.0, d1 <= 45.0) and np.array_equal(d0[d0 <= 45.0], d1[d1 <= 45.0])
    assert int(er["n_events"].cpu()[0]) < ref["n_events"]


@pytest.mark.parametrize("T", [200, 2500])
def test_abc_long_observed_series_vs_oracle(sem, c_oracle, T):
    """The reference puts no limit on the length of the observed series (abc_algo.py:58-99): 200 days (staged in shared
    memory) and 2500 days (read from L2) against the oracle."""
    import workloads
    truth = workloads.sir_truth((990, 10, 0), T, .12, .05)
    obs = truth.astype(int).astype(float)                         # rows (S, I, R); the trial starts from Poisson(obs[0]) (abc_algo.py:39-40)
    n = 400
    thr = 60.0
    out = sem.engine.abc_trials(obs, n, thr, [0, .5, 0, .5], seed=5, trial0=10, arith=3, want_traj=True)
    ref = c_oracle.abc_trials(obs, n, thr, (0, .5, 0, .5), arith=3, seed=5, trial0=10)
    assert np.array_equal(out["theta"].cpu().numpy(), ref["theta"])
    assert np.array_equal(out["traj"].cpu().numpy(), ref["traj"])
    np.testing.assert_allclose(out["distance"].cpu().numpy(), ref["distance"], rtol=1e-13)


def test_abc_algo_dropin(sem, c_oracle):
    g = golden("abc_sir_small")
    obs = g["observed"]
    st = {}
    post, traj = sem.abc_algo.abc_algo(obs, 5, 45.0, {"beta": [0, 5], "gamma": [0, 5]}, seed=7, batch=4096, stats=st)
    assert sorted(post) == ["beta", "gamma"] and len(post["beta"]) == 5 and traj.shape == (5, obs.shape[0], 4)
    assert np.array_equal(traj[0, :, 0], np.arange(obs.shape[0]))
    # same accepted trials as the oracle's sequential acceptance over trial ids 0,1,2,...
    ref = c_oracle.abc_trials(obs, 4096, 45.0, (0, 5, 0, 5), arith=3, seed=7, trial0=0)     # fast32 is the default
    acc = np.nonzero(ref["distance"] <= 45.0)[0][:5]
    assert np.array_equal(st["accepted_ids"], acc)
    assert np.array_equal(np.array(post["beta"]), ref["theta"][acc, 0])
    assert np.array_equal(traj[:, :, 1:], ref["traj"][acc].astype(float))
    for k in range(5):
        d = sem.abc_algo.distance_function(traj[k, :, 2], obs[:, 1], traj[k, :, 3], obs[:, 2])
        assert d <= 45.0 and abs(d - st["distances"][k]) < 1e-9


# ------------------------------------------------------------------ PMCMC drop-in
def test_particle_mcmc_dropin(sem):
    np.random.seed(3)
    Y = _truth_Y(0, 10, 5, .1, False)
    st = {}
    thetas, lik, trajs = sem.particle_mcmc(Y, sem.ModelType.SIR, [2.0, 1.0], .01, n_chains=40, probs=.1, n_particles=2000,
                                           n_population=1000, mu=20, seed=11, stats=st)
    assert thetas.shape == (40, 2) and lik.shape == (40,) and trajs.shape == (10, 40, 3)
    assert np.all(thetas > 0) and np.all(lik > 0)
    assert np.all(trajs.sum(2) == 1000)
    assert 1 <= st["acceptances"] <= 40 and st["filter_runs"] >= 1
    # p_obs estimated (probs=None): last theta component is the clipped observation probability (pmcmc.py:283-287)
    th2, _, _ = sem.particle_mcmc(Y, sem.ModelType.SIR, [2.0, 1.0, .1], .001, n_chains=5, probs=None, n_particles=500,
                                  n_population=1000, mu=20, seed=12)
    assert th2.shape == (5, 3) and np.all((th2[:, 2] >= 0) & (th2[:, 2] <= 1))


def test_host_buffer_entry_point(sem, c_oracle):
    """sem_pf_run_host: the C-ABI call with HOST buffers returns the reference's (T,N,C)/(T,N) float64 layout."""
    L = sem._lib.load()
    N, T = 640, 9
    Y = _truth_Y(0, T, 5, .1, False)
    cfg = sem.engine.make_pf_config(0, N, T, probs=.1, resampler=1, arith=1, seed=321, mu=[20], n_population=[1000])
    theta = np.array([2.0, 1.0])
    lz = np.zeros(T); z = np.zeros(T); H = np.zeros((T, N, 3)); A = np.zeros((T, N)); ev = np.zeros(1, dtype=np.uint64)
    p = lambda v: v.ctypes.data_as(C.c_void_p)
    rc = L.sem_pf_run_host(C.byref(cfg), p(Y), p(theta), None, p(lz), p(z), p(H), p(A), p(ev))
    assert rc == 0
    ref = c_oracle.pf_run(0, Y, theta, False, .1, N, resampler=1, arith=1, seed=321, mu=[20], npop=[1000])
    assert np.array_equal(H, ref["X_hist"].astype(float)) and np.array_equal(A, ref["ancestry"].astype(float))
    np.testing.assert_allclose(lz, ref["log_zetas"], rtol=1e-11)
    np.testing.assert_allclose(z, np.exp(ref["log_zetas"]), rtol=1e-10)
    assert int(ev[0]) == ref["n_events"]


# ------------------------------------------------------------------ particle-sharded filter
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("model,G,theta,npop,mu", [(0, 1, [2.0, 1.0], [1000], [20]), (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20])])
def test_sharded_filter_equals_single_gpu(sem, c_oracle, world, model, G, theta, npop, mu):
    """W shards with global systematic resampling + particle migration (lock-step emulation of the ranks on one GPU,
    same kernels and host logic as the NCCL path) reproduce the unsharded filter and the oracle exactly."""
    import torch
    from sem_b200 import sharded
    N, T = 3001, (9 if model == 0 else 6)
    Y = _truth_Y(model, T, 5, .1, False, G=G)
    out = sharded.run_local(Y, model, np.array(theta, float), N, world, G=G, probs=.1, seed=4242, filter_id=5, mu=mu,
                            n_population=npop)
    torch.cuda.synchronize()
    assert out["collapsed"] == 0
    X = torch.cat([sh.X_hist for sh in out["shards"]], dim=2).permute(0, 2, 1).cpu().numpy()       # (T,N,C)
    A = torch.cat([sh.ancestry for sh in out["shards"]], dim=1).cpu().numpy()
    arith = sem.engine.resolve_arith(model, "auto")                       # the sharded filter's default
    ref = c_oracle.pf_run(model, Y, theta, False, .1, N, G=G, resampler=1, arith=arith, seed=4242, filter_id=5, mu=mu, npop=npop)
    assert np.array_equal(A, ref["ancestry"])
    assert np.array_equal(X, ref["X_hist"])
    np.testing.assert_allclose(out["log_zetas"], ref["log_zetas"], rtol=1e-11)
    assert sum(sh.n_events for sh in out["shards"]) == ref["n_events"]
    cfg = sem.engine.make_pf_config(model, N, T, G=G, probs=.1, resampler=1, arith=arith, seed=4242, filter_id0=5, mu=mu,
                                    n_population=npop)
    one = sem.engine.run_pf(cfg, Y, np.array(theta, float))
    assert np.array_equal(one.X_hist[0].permute(0, 2, 1).cpu().numpy(), X)
    assert np.array_equal(one.ancestry[0].cpu().numpy(), A)


PEER_CASES = [
    # world, model, G, theta, npop, mu, arith, n_local, block
    (1, 0, 1, [.4, .2], [1000], [20], 4, 3000, 256),              # one rank: the exchange degenerates to push-resampling on one GPU
    (2, 0, 1, [.4, .2], [1000], [20], 4, 2048, 256),              # uniformized32, sorted layout (8 chunks per CTA)
    (3, 0, 1, [.4, .2], [1000], [20], 4, 1500, 300),              # ragged CTAs, sorted layout with helper legs (4W+2 chunks)
    (4, 0, 1, [2.0, 1.0], [1000], [20], 3, 1100, 160),            # direct method, balanced layout (128 + 32)
    (2, 1, 1, [.4, .3, .2], [1000], [20], 4, 2000, 256),          # SEIR (5-word records -> 8-word slots)
    (2, 3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 3, 1024, 128),   # 2 subgroups, group-summed observations (config 5 shape)
    (8, 0, 1, [.4, .2], [1000], [20], 4, 512, 128),               # 8 ranks, 4 CTAs each
]


@pytest.mark.parametrize("world,model,G,theta,npop,mu,arith,n_local,block", PEER_CASES)
def test_peer_exchange_equals_single_gpu_and_oracle(sem, c_oracle, world, model, G, theta, npop, mu, arith, n_local, block):
    """Device-side exchange (sem_pf_run_sharded: partials and child records through peer memory, one cooperative launch
    per rank, the ranks' kernels running concurrently on this GPU): states, GLOBAL ancestor indices, log-likelihoods and
    event counts equal the single-GPU filter of W n particles and the C oracle; the path sample over the shards equals
    the single-GPU iteration result; two passes back to back (generations keep counting, no reset in between)."""
    import torch
    from sem_b200 import sharded
    N, T = world * n_local, (10 if model == 0 else 7)
    Y = _truth_Y(model, T, 5, .1, False, G=G)
    th = np.array(theta, float)
    outs = sharded.run_peer_local(Y, model, th, N, world, G=G, probs=.1, seed=4242, filter_id=5, mu=mu, n_population=npop,
                                  arith=arith, block_particles=block, want_path=True, passes=2, timeout_s=10.0)
    for k, out in enumerate(outs):
        assert out["status"] == [0] * world, out["status"]
        cfg = sem.engine.make_pf_config(model, N, T, G=G, probs=.1, resampler=1, arith=arith, seed=4242, filter_id0=5 + k, mu=mu,
                                        n_population=npop, block_particles=block)
        it = torch.empty((1, sem.engine.ITER_HEADER + T * cfg_cols(model, G)), dtype=torch.float64, device="cuda")
        one = sem.engine.run_pf(cfg, Y, th, iter_out=it)
        torch.cuda.synchronize()
        assert np.array_equal(one.ancestry[0].cpu().numpy(), out["ancestry"])
        assert np.array_equal(one.X_hist[0].cpu().numpy(), out["X_hist"])
        for r in range(world):                                   # every rank holds the same likelihood, bit for bit
            assert np.array_equal(out["log_zetas"][r], out["log_zetas"][0])
        np.testing.assert_allclose(out["log_zetas"][0], one.log_zetas[0].cpu().numpy(), rtol=1e-12)
        assert out["n_events"] == int(one.n_events[0])
        ito = it[0].cpu().numpy()
        for r in range(world):                                   # same chosen particle and trajectory on every rank
            got = out["iteration"][r]
            assert got[1] == 0 and got[3] == ito[3]
            assert np.array_equal(got[sem.engine.ITER_HEADER:], ito[sem.engine.ITER_HEADER:])
            np.testing.assert_allclose(got[0], ito[0], rtol=1e-12)
        if k == 0:
            ref = c_oracle.pf_run(model, Y, theta, False, .1, N, G=G, resampler=1, arith=arith, seed=4242, filter_id=5, mu=mu, npop=npop)
            assert np.array_equal(out["ancestry"], ref["ancestry"])
            assert np.array_equal(np.transpose(out["X_hist"], (0, 2, 1)), ref["X_hist"])
            np.testing.assert_allclose(out["log_zetas"][0], ref["log_zetas"], rtol=1e-11)


def cfg_cols(model, G):
    return 3 if model == 0 else 4 if model == 1 else 3 * G


def test_peer_exchange_collapse_and_reset(sem):
    """A collapsing sharded filter ends on every rank with the same status (no hang), and after reset() the arenas serve
    the next pass."""
    from sem_b200 import sharded
    T = 8
    Y = _truth_Y(0, T, 5, .1, False)
    outs = sharded.run_peer_local(Y, 0, np.array([0.01, 50.0]), 2 * 1024, 2, probs=.1, seed=3, mu=[20], n_population=[1000],
                                  arith=3, block_particles=256, timeout_s=10.0)
    st = outs[0]["status"]
    assert st[0] == st[1] and st[0] > 0, st
    outs = sharded.run_peer_local(Y, 0, np.array([2.0, 1.0]), 2 * 1024, 2, probs=.1, seed=3, mu=[20], n_population=[1000],
                                  arith=3, block_particles=256, timeout_s=10.0)
    assert outs[0]["status"] == [0, 0]


def test_peer_exchange_degenerate_weights(sem, c_oracle):
    """Weight degeneracy: a handful of particles own thousands of children each (the warp-cooperative record store) --
    normal observations with a tiny noise ratio make almost every weight vanish."""
    import torch
    from sem_b200 import sharded
    T, world, n_local = 6, 2, 2048
    Y = _truth_Y(0, T, 5, .02, True)
    th = np.array([2.0, 1.0])
    out = sharded.run_peer_local(Y, 0, th, world * n_local, world, observations=True, probs=.02, seed=12, mu=[20],
                                 n_population=[1000], arith=3, block_particles=256, timeout_s=10.0)[0]
    cfg = sem.engine.make_pf_config(0, world * n_local, T, observations=True, probs=.02, resampler=1, arith=3, seed=12, mu=[20],
                                    n_population=[1000], block_particles=256)
    one = sem.engine.run_pf(cfg, Y, th)
    torch.cuda.synchronize()
    assert out["status"] == [int(one.status[0])] * world
    if out["status"][0] == 0:
        anc = one.ancestry[0].cpu().numpy()
        assert np.array_equal(anc, out["ancestry"]) and np.array_equal(one.X_hist[0].cpu().numpy(), out["X_hist"])
        assert max(np.bincount(anc[p]).max() for p in range(1, T)) > 64      # the case really is degenerate


def test_predict_forward_daily_states(sem, c_oracle):
    """Forward-prediction fan-out (tests/pred_tmps.py:55-73): daily states of many simulations == the oracle's event
    logs sampled at the integer times."""
    n, H = 200, 6
    rng = np.random.RandomState(3)
    thetas = np.stack([rng.uniform(1, 3, n), rng.uniform(.5, 1.5, n)], 1)
    last = np.stack([rng.randint(200, 480, n), rng.randint(0, 30, n), rng.randint(0, 100, n)], 1)
    out = sem.gillespie_algo.predict_forward("sir", thetas, last, H, seed=17).cpu().numpy()
    assert out.shape == (n, H, 3)
    for i in range(0, n, 7):
        ref = c_oracle.ssa(0, 1, last[i], thetas[i], float(H), arith=3, seed=17, sim_index=i, max_rec=4000)
        for d in range(1, H + 1):
            k = np.searchsorted(ref["times"], d, side="right") - 1          # last event with time <= d
            assert np.array_equal(out[i, d - 1], ref["states"][k].astype(np.int32)), (i, d)


@pytest.mark.parametrize("unif", [2, 4])
def test_uniformized_headline_workload_vs_oracle(sem, c_oracle, unif):
    """arith='uniformized' / 'uniformized32' on the BASELINE workload shape (pop 1e4, 101 rows) at 8000 particles: exact
    match with the oracle, and the likelihood estimate agrees with the direct method's within Monte-Carlo error."""
    import torch
    N, T, pop = 8000, 101, 10_000
    Y = bench_truth(T, pop)
    cfg = sem.engine.make_pf_config(0, N, T, probs=.1, resampler=1, arith=unif, seed=31, mu=[20], n_population=[pop])
    r = sem.engine.run_pf(cfg, Y, np.array([.4, .2]))
    o = c_oracle.pf_run(0, Y, [.4, .2], False, .1, N, resampler=1, arith=unif, seed=31, mu=[20], npop=[pop])
    torch.cuda.synchronize()
    assert np.array_equal(r.ancestry[0].cpu().numpy(), o["ancestry"])
    assert np.array_equal(r.X_hist[0].permute(0, 2, 1).cpu().numpy(), o["X_hist"])
    np.testing.assert_allclose(r.log_zetas[0].cpu().numpy(), o["log_zetas"], rtol=1e-11)
    lz = {a: [] for a in (1, unif)}
    lz[2] = lz[unif]
    for a in (1, unif):
        for sd in range(6):
            c = sem.engine.make_pf_config(0, 20000, T, probs=.1, resampler=1, arith=a, seed=100 + sd, mu=[20], n_population=[pop])
            lz[a].append(float(sem.engine.run_pf(c, Y, np.array([.4, .2])).log_zetas[0, -1].cpu()))
    se = np.sqrt(np.var(lz[1]) / 6 + np.var(lz[2]) / 6) + 0.02
    assert abs(np.mean(lz[1]) - np.mean(lz[2])) < 5 * se, (lz, se)


# ------------------------------------------------------------------ statistical parity with the reference (north_star check 3)
def test_logz_distribution_vs_reference(sem):
    """Likelihood-estimate distribution of the CUDA filter vs 300 runs of the unmodified reference particle_filter
    at the same N (tests/golden/make_golden_stats.py): equal means (both unbiased for the same likelihood); with the
    reference's multinomial resampling also the same law of log Z (KS)."""
    import torch
    from scipy import stats
    g = golden("stat_logz_sir")
    N, npop, mu = int(g["n_particles"]), int(g["n_population"]), float(g["mu"])
    ref = g["zetas_last"]
    runs = 1200
    for resampler, arith in [(0, 3), (1, 3), (0, 4), (1, 4), (0, 1), (0, 0), (1, 1), (0, 2)]:   # 3 = fast32, the production default
        cfg = sem.engine.make_pf_config(0, N, len(g["Y"]), n_filters=runs, probs=float(g["probs"]), resampler=resampler,
                                        arith=arith, seed=2024 + 7 * resampler + arith, mu=[mu], n_population=[npop])
        res = sem.engine.run_pf(cfg, g["Y"], np.tile(g["theta"], (runs, 1)))
        torch.cuda.synchronize()
        assert int((res.status != 0).sum()) == 0
        z = np.exp(res.log_zetas[:, -1].cpu().numpy())
        se = np.sqrt(ref.var() / ref.size + z.var() / z.size)
        assert abs(z.mean() - ref.mean()) < 4 * se, (resampler, arith, z.mean(), ref.mean(), se)
        if resampler == 0:
            assert stats.ks_2samp(np.log(z), np.log(ref)).pvalue > 1e-3, (arith,)
        else:
            assert np.log(z).std() <= np.log(ref).std() * 1.1                       # systematic resampling: no more variance


def test_pmcmc_posterior_vs_reference(sem):
    """Posterior of (beta, gamma) from the drop-in particle_mcmc vs a chain of the unmodified reference particle_mcmc on
    the same data and settings: means agree within Monte-Carlo error (standard errors from the chains' ESS)."""
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "stat_pmcmc_sir.npz")):
        pytest.skip("reference PMCMC chain golden not generated")
    g = golden("stat_pmcmc_sir")
    ref = g["thetas"][300:]
    np.random.seed(11)
    thetas, lik, _ = sem.particle_mcmc(g["Y"], sem.ModelType.SIR, list(g["parameters"]), float(g["h"]), n_chains=6000,
                                       probs=float(g["probs"]), n_particles=int(g["n_particles"]),
                                       n_population=int(g["n_population"]), mu=float(g["mu"]), seed=5, resampler="multinomial")
    ours = thetas[500:]
    ess = sem.helpers.effective_sample_size
    for k in range(2):
        se = np.sqrt(ref[:, k].var() / max(ess(ref[:, k]), 4) + ours[:, k].var() / max(ess(ours[:, k]), 4))
        assert abs(ours[:, k].mean() - ref[:, k].mean()) < 4.5 * se, (k, ours[:, k].mean(), ref[:, k].mean(), se)
        assert 0.5 < ours[:, k].std() / ref[:, k].std() < 2.0
    acc_ref, acc_ours = sem.helpers.acceptance_rate(g["thetas"]), sem.helpers.acceptance_rate(thetas)
    assert abs(acc_ref - acc_ours) < 0.12, (acc_ref, acc_ours)


# ---- the DEFAULT path (arith="auto" -> uniformized32, systematic resampling) in the regime where auto selects it
def test_logz_distribution_vs_reference_slow_growth(sem):
    """Slow-growth regime (beta=.4, gamma=.2: growth 0.2 per interval, where engine.resolve_arith picks uniformized32):
    320 runs of the unmodified reference particle_filter (tests/golden/make_golden_stats_slow.py) vs the CUDA filter with
    the shipped defaults.  Both estimators are unbiased for the same likelihood: means within 4 se; with the reference's
    multinomial resampling the whole law of log Z agrees (KS); the direct method agrees too."""
    import torch
    from scipy import stats
    g = golden("stat_logz_sir_slow")
    N, npop, mu = int(g["n_particles"]), int(g["n_population"]), float(g["mu"])
    ref = g["zetas_last"]
    assert sem.engine.resolve_arith(0, "auto", theta=g["theta"]) == sem.engine.ARITH["uniformized32"]
    runs = 4000
    for resampler, arith in [(1, "auto"), (0, "auto"), (1, 3), (0, 3), (0, 0)]:
        a = sem.engine.resolve_arith(0, arith, theta=g["theta"])
        cfg = sem.engine.make_pf_config(0, N, len(g["Y"]), n_filters=runs, probs=float(g["probs"]), resampler=resampler,
                                        arith=a, seed=911 + 7 * resampler + a, mu=[mu], n_population=[npop])
        res = sem.engine.run_pf(cfg, g["Y"], np.tile(g["theta"], (runs, 1)))
        torch.cuda.synchronize()
        assert int((res.status != 0).sum()) == 0
        z = np.exp(res.log_zetas[:, -1].cpu().numpy())
        se = np.sqrt(ref.var() / ref.size + z.var() / z.size)
        assert abs(z.mean() - ref.mean()) < 4 * se, (resampler, arith, z.mean(), ref.mean(), se)
        if resampler == 0:
            assert stats.ks_2samp(np.log(z), np.log(ref)).pvalue > 1e-3, (arith, stats.ks_2samp(np.log(z), np.log(ref)))
        else:
            assert np.log(z).std() <= np.log(ref).std() * 1.1


def test_pmcmc_posterior_vs_reference_slow_growth(sem):
    """Posterior of (beta, gamma) from the drop-in particle_mcmc WITH ITS DEFAULTS (arith='auto' -> uniformized32,
    systematic resampling) vs four chains of the unmodified reference particle_mcmc on the same data and settings:
    posterior means and both endpoints of the 95 % HDI agree within Monte-Carlo error (spread between chains)."""
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "stat_pmcmc_sir_slow.npz")):
        pytest.skip("reference PMCMC chains (slow growth) not generated")
    g = golden("stat_pmcmc_sir_slow")
    burn = 400
    ref = g["thetas"][:, burn:]                                              # (chains, iterations, 2)
    ours = []
    for c in range(4):
        np.random.seed(100 + c)
        th, lik, _ = sem.particle_mcmc(g["Y"], sem.ModelType.SIR, list(g["parameters"]), float(g["h"]), n_chains=6000,
                                       probs=float(g["probs"]), n_particles=int(g["n_particles"]),
                                       n_population=int(g["n_population"]), mu=float(g["mu"]), seed=500 + c)
        ours.append(th[burn:])
    ours = np.stack(ours)
    hdi = sem.helpers.hdi
    for k in range(2):
        stat_ref = np.array([[c[:, k].mean(), *hdi(c[:, k], .95)] for c in ref])       # per chain: mean, HDI low, HDI high
        stat_our = np.array([[c[:, k].mean(), *hdi(c[:, k], .95)] for c in ours])
        for q, name in enumerate(("mean", "hdi_lo", "hdi_hi")):
            se = np.sqrt(stat_ref[:, q].var(ddof=1) / len(stat_ref) + stat_our[:, q].var(ddof=1) / len(stat_our))
            diff = abs(stat_ref[:, q].mean() - stat_our[:, q].mean())
            assert diff < 4.5 * se + 0.004, (k, name, stat_ref[:, q], stat_our[:, q], se)
    r_hat = sem.helpers.gelman_rubin_test([c for c in ours])
    assert np.all(np.asarray(r_hat) < 1.1), r_hat
    acc_ref = np.mean([sem.helpers.acceptance_rate(c) for c in g["thetas"]])
    acc_our = np.mean([sem.helpers.acceptance_rate(c) for c in ours])
    assert acc_our > acc_ref - 0.05, (acc_ref, acc_our)        # systematic resampling: no noisier a likelihood, no fewer acceptances


def test_pmcmc_lookahead_posterior_vs_reference(sem):
    """particle_mcmc(lookahead=8): eight proposals per launch under the rejection assumption.  The chain
    has the law of the sequential loop: posterior means and 95 % HDI endpoints agree with four chains of the unmodified
    reference within Monte-Carlo error, the acceptance rate is the sequential one, and a launch advances the chain by
    more than two iterations on average."""
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "stat_pmcmc_sir_slow.npz")):
        pytest.skip("reference PMCMC chains (slow growth) not generated")
    g = golden("stat_pmcmc_sir_slow")
    burn = 400
    ref = g["thetas"][:, burn:]
    ours, stats_all = [], []
    for c in range(4):
        np.random.seed(300 + c)
        st = {}
        th, lik, traj = sem.particle_mcmc(g["Y"], sem.ModelType.SIR, list(g["parameters"]), float(g["h"]), n_chains=6000,
                                          probs=float(g["probs"]), n_particles=int(g["n_particles"]),
                                          n_population=int(g["n_population"]), mu=float(g["mu"]), seed=700 + c, lookahead=8, stats=st)
        assert np.all(np.isfinite(th)) and np.all(np.isfinite(lik)) and np.all(traj >= 0)
        ours.append(th[burn:]); stats_all.append(st)
    ours = np.stack(ours)
    hdi = sem.helpers.hdi
    for k in range(2):
        stat_ref = np.array([[c[:, k].mean(), *hdi(c[:, k], .95)] for c in ref])
        stat_our = np.array([[c[:, k].mean(), *hdi(c[:, k], .95)] for c in ours])
        for q, name in enumerate(("mean", "hdi_lo", "hdi_hi")):
            se = np.sqrt(stat_ref[:, q].var(ddof=1) / len(stat_ref) + stat_our[:, q].var(ddof=1) / len(stat_our))
            diff = abs(stat_ref[:, q].mean() - stat_our[:, q].mean())
            assert diff < 4.5 * se + 0.004, (k, name, stat_ref[:, q], stat_our[:, q], se)
    acc_ref = np.mean([sem.helpers.acceptance_rate(c) for c in g["thetas"]])
    acc_our = np.mean([sem.helpers.acceptance_rate(c) for c in ours])
    assert abs(acc_our - acc_ref) < 0.08, (acc_ref, acc_our)
    for st in stats_all:
        assert st["lookahead"] == 8 and st["launches"] < 0.7 * 6000, st    # (acceptance ~0.55 here: ~1.8 iterations per launch)


def test_pmcmc_auto_switches_interval_simulation_mid_chain(sem):
    """arith='auto' is resolved per proposal: around growth 0.5 per interval a small filter's chain alternates between the
    direct method and uniformized intervals (both exact); the chain stays well-behaved and both kernels ran."""
    import workloads
    Y = workloads.observe_binomial(workloads.sir_truth((980, 20, 0), 12, 1.0, .5), .2, seed=2)
    np.random.seed(8)
    st = {}
    th, lik, traj = sem.particle_mcmc(Y, sem.ModelType.SIR, [1.0, .5], 4e-3, n_chains=400, probs=.2, n_particles=600, n_population=1000,
                                      mu=20, seed=21, stats=st)
    assert st["arith"] == {"fast32", "uniformized32"}, st["arith"]
    assert np.all(np.isfinite(th)) and np.all(th > 0) and np.all(np.isfinite(lik)) and np.all(traj.sum(2) == 1000)
    assert 20 < st["acceptances"] < 390
    assert abs(th[100:, 0].mean() - 1.0) < .25 and abs(th[100:, 1].mean() - .5) < .15


def test_pmcmc_lookahead_with_p_obs_and_adaptive(sem):
    """lookahead with the estimated observation probability (one p_obs per filter of the batch) and the adaptive
    covariance computed under the rejection assumption: same invariants as the sequential loop."""
    import workloads
    Y = workloads.observe_binomial(workloads.sir_truth((980, 20, 0), 12, 2.0, 1.0), .3, seed=4)
    np.random.seed(5)
    st = {}
    th, lik, traj = sem.particle_mcmc(Y, sem.ModelType.SIR, [2.0, 1.0, .3], 1e-3, adaptive=True, n_chains=1300, probs=None,
                                      n_particles=400, n_population=1000, mu=20, seed=9, lookahead="auto", stats=st)
    assert th.shape == (1300, 3) and np.all((th[:, 2] >= 0) & (th[:, 2] <= 1)) and np.all(th[:, :2] > 0)
    assert np.all(np.isfinite(lik)) and np.all(lik > 0)
    assert np.all(traj.sum(axis=2) == 1000)                        # every sampled trajectory conserves the population
    assert st["lookahead"] == 16 and st["acceptances"] > 20
    assert 0.05 < np.mean(th[300:, 2]) < 0.9


@pytest.mark.parametrize("x0", [(8000, 500, 1500), (9990, 3, 7), (3000, 1500, 5500)])
def test_uniformized_end_state_law_vs_direct_method(sem, x0):
    """Law of the state at the END of one observation interval, headline theta / population 1e4: 1.2e6 propagations of
    one fixed state with the uniformized interval simulation (arith 4, the filters' default) vs the direct method in
    the reference's operation order (arith 0) -- two-sample chi-square on the joint histogram of (infections,
    recoveries); includes a start with I = 3, where extinction inside the interval matters."""
    import torch
    from scipy import stats
    n = 1_200_000
    out = {}
    for arith in (0, 4):
        r = sem.engine.simulate(0, np.array(x0), np.array([.4, .2]), 1.0, arith=arith, seed=99 + arith, n_sims=n)
        torch.cuda.synchronize()
        x = r["x"].cpu().numpy().astype(np.int64)
        assert np.all(x.sum(1) == sum(x0)) and np.all(x >= 0)
        out[arith] = np.stack([x0[0] - x[:, 0], x[:, 2] - x0[2]], 1)          # (infections, recoveries) in the interval
    both = np.concatenate([out[0], out[4]])
    w = [max(1, int(np.ceil(both[:, k].std() / 3))) for k in range(2)]         # ~ 1/3 sd wide cells
    key = lambda d: (d[:, 0] // w[0]) * 100_000 + d[:, 1] // w[1]
    cells, inv = np.unique(np.concatenate([key(out[0]), key(out[4])]), return_inverse=True)
    a = np.bincount(inv[:n], minlength=len(cells)).astype(float)
    b = np.bincount(inv[n:], minlength=len(cells)).astype(float)
    big = (a + b) >= 40
    a2 = np.append(a[big], a[~big].sum()); b2 = np.append(b[big], b[~big].sum())   # the sparse cells pooled into one
    keep = (a2 + b2) > 0
    chi2 = float((((a2 - b2) ** 2) / (a2 + b2))[keep].sum())
    dof = int(keep.sum()) - 1
    pval = stats.chi2.sf(chi2, dof)
    assert dof >= 20 and pval > 1e-4, (x0, chi2, dof, pval)
    for k in range(2):                                                         # and the marginal means, z-test
        se = np.sqrt(out[0][:, k].var() / n + out[4][:, k].var() / n)
        assert abs(out[0][:, k].mean() - out[4][:, k].mean()) < 4.5 * se + 1e-9, (x0, k)


NAMED_SIZE_CASES = [
    # name, model, G, theta, npop, mu, observations, probs, N, T, arith
    ("headline_sir_1e5x101_uniformized32", 0, 1, [.4, .2], [10_000], [20], False, .1, 100_000, 101, 4),
    ("config3_seir_pop1e4_T101", 1, 1, [.4, .1, .1], [10_000], [20], False, .1, 20_000, 101, 4),
    ("config4i_sir_normal_pop1e4_T101", 0, 1, [.4, .2], [10_000], [20], True, .1, 20_000, 101, 4),
    ("config5_shape_sub2_pop1e5_T15", 3, 2, [5, 2, 1, 3, .5], [40_000, 60_000], [600, 800], False, .1, 20_000, 15, 3),
]


@pytest.mark.parametrize("name,model,G,theta,npop,mu,observations,probs,N,T,arith", NAMED_SIZE_CASES, ids=[c[0] for c in NAMED_SIZE_CASES])
def test_named_sizes_match_oracle_exactly(sem, c_oracle, name, model, G, theta, npop, mu, observations, probs, N, T, arith):
    """BASELINE's named configurations at (or near) their named sizes, default arithmetic of each model family, whole-filter
    kernel in its production layout (N = 1e5: sorted, 4W+2 chunks with helper legs): states, ancestors and event counts
    equal the OpenMP oracle's bit for bit, log-likelihoods to 1e-11."""
    import torch
    import workloads
    if model == 0:
        truth = workloads.sir_truth((npop[0] - mu[0], mu[0], 0), T, *theta)
    elif model == 1:
        truth = workloads.seir_truth((npop[0] - mu[0], 0, mu[0], 0), T, *theta)
    else:
        y0 = [(npop[g] - mu[g], mu[g], 0) for g in range(G)]
        truth = workloads.subgroups_truth(y0, T, np.array(theta[:G * G]).reshape(G, G), theta[-1]).reshape(T, G, 3).sum(1)
    Y = workloads.observe_normal(truth, probs, seed=3) if observations else workloads.observe_binomial(truth, probs, seed=3)
    assert sem.engine.resolve_arith(model, "auto", theta=theta) == arith          # the default of this model family and regime
    cfg = sem.engine.make_pf_config(model, N, T, G=G, observations=observations, probs=probs, resampler=1, arith=arith, seed=606,
                                    mu=mu, n_population=npop)
    r = sem.engine.run_pf(cfg, Y, np.array(theta, float))
    o = c_oracle.pf_run(model, Y, theta, observations, probs, N, G=G, resampler=1, arith=arith, seed=606, mu=mu, npop=npop)
    torch.cuda.synchronize()
    assert int(r.status.cpu()[0]) == 0 == o["collapsed"]
    assert r.launches == 1                                                        # the whole-filter kernel
    assert np.array_equal(r.ancestry[0].cpu().numpy(), o["ancestry"])
    assert np.array_equal(r.X_hist[0].permute(0, 2, 1).cpu().numpy(), o["X_hist"])
    np.testing.assert_allclose(r.log_zetas[0].cpu().numpy(), o["log_zetas"], rtol=1e-11)
    assert int(r.n_events[0]) == o["n_events"]


# ------------------------------------------------------------------ robustness / edge cases
@pytest.mark.parametrize("pop,mu0,theta,T,N", [(250_000, 800, [.5, .2], 9, 3000),       # batch means 1e3 .. 1e4: alias tables AND the PTRS fallback (> 4096)
                                               (1000, 2, [.15, .4], 10, 5000),            # dying epidemic: means below 4 (the grid's first table), absorptions
                                               (30_000, 3000, [3.0, .3], 6, 2000)])        # explosive growth: many growth-capped batches per interval
def test_uniformized_candidate_count_regimes_vs_oracle(sem, c_oracle, pop, mu0, theta, T, N):
    """The candidate counts of a uniformized batch come from alias tables for means in [4, 4096] (smaller means use the
    table of 4, larger ones the PTRS sampler): filters whose batch means sit below, inside, across and above that range
    equal the oracle bit for bit."""
    import torch
    import workloads
    Y = workloads.observe_binomial(workloads.sir_truth((pop - mu0, mu0, 0), T, theta[0], theta[1]), .1, seed=3)
    cfg = sem.engine.make_pf_config(0, N, T, probs=.1, resampler=1, arith=4, seed=31337, filter_id0=5, mu=[mu0], n_population=[pop])
    res = sem.engine.run_pf(cfg, Y, np.array(theta, float))
    torch.cuda.synchronize()
    ref = c_oracle.pf_run(0, Y, theta, False, .1, N, resampler=1, arith=4, seed=31337, filter_id=5, mu=[mu0], npop=[pop])
    assert int(res.status.cpu()[0]) == ref["collapsed"]
    if ref["collapsed"] == 0:
        assert np.array_equal(res.X_hist[0].permute(0, 2, 1).cpu().numpy(), ref["X_hist"])
        assert np.array_equal(res.ancestry[0].cpu().numpy(), ref["ancestry"])
        np.testing.assert_allclose(res.log_zetas[0].cpu().numpy(), ref["log_zetas"], rtol=1e-11, atol=1e-11)
        assert int(res.n_events.cpu()[0]) == ref["n_events"]


def test_many_ctas_global_prefix_path(sem, c_oracle):
    """More CTAs than fit the shared-memory prefix stage (nb > 4096) and than one finalize chunk: same answers."""
    import torch
    N, T = 150_001, 4
    Y = _truth_Y(0, T, 5, .1, False)
    cfg = sem.engine.make_pf_config(0, N, T, probs=.1, resampler=1, arith=1, seed=8, mu=[20], n_population=[1000], block_particles=32)
    r = sem.engine.run_pf(cfg, Y, np.array([2.0, 1.0]))
    o = c_oracle.pf_run(0, Y, [2.0, 1.0], False, .1, N, resampler=1, arith=1, seed=8, mu=[20], npop=[1000])
    torch.cuda.synchronize()
    assert np.array_equal(r.ancestry[0].cpu().numpy(), o["ancestry"])
    assert np.array_equal(r.X_hist[0].permute(0, 2, 1).cpu().numpy(), o["X_hist"])
    np.testing.assert_allclose(r.log_zetas[0].cpu().numpy(), o["log_zetas"], rtol=1e-11)


@pytest.mark.parametrize("model,G,theta,npop,mu,arith,resampler", [
    (3, 2, [2.665, 1.706, 1.764, 1.242, 0.934], [3, 4], [0, 1], 3, 1),    # the case tests/fuzz_parity.py found (Poisson(1) > 4)
    (0, 1, [1.2, .4], [6], [4], 4, 1), (0, 1, [1.2, .4], [6], [4], 3, 0), (1, 1, [2.0, .5, .4], [5], [3], 3, 1)])
def test_negative_initial_susceptibles_collapse_like_the_reference(sem, c_oracle, model, G, theta, npop, mu, arith, resampler):
    """pmcmc.py:156-169 sets S0 = n_population - Poisson(mu) without a clamp: with a tiny population some particles start
    with S0 < 0.  scipy has no weight for them (nan), np.random.choice refuses the weights and the reference returns
    (None, None, None) (pmcmc.py:187-192; tests/golden/negative_s0_reference.npz): the filter collapses at the
    first step, on the device as in the oracle, quickly (the exchange once mistook a negative record word for "empty" and
    every such launch sat out its time limit), and the initial states -- negative counts included -- are the oracle's."""
    import time
    import torch
    import workloads
    N, T, F = 1000, 4, 5
    truth = (workloads.subgroups_truth(((npop[0] - mu[0], mu[0], 0), (npop[1] - mu[1], mu[1], 0)), T, np.array(theta[:4]).reshape(2, 2), theta[4]).reshape(T, 2, 3).sum(1)
             if model == 3 else workloads.seir_truth((npop[0] - mu[0], 0, mu[0], 0), T, *theta) if model == 1
             else workloads.sir_truth((npop[0] - mu[0], mu[0], 0), T, *theta))
    Y = workloads.observe_binomial(truth, .4, seed=7)
    thetas = np.array([theta] * F) * (1 + 0.05 * np.arange(F))[:, None]
    cfg = sem.engine.make_pf_config(model, N, T, G=G, n_filters=F, probs=.4, resampler=resampler, arith=arith, seed=99, filter_id0=3, mu=mu, n_population=npop)
    t0 = time.time()
    r = sem.engine.run_pf(cfg, Y, thetas)
    torch.cuda.synchronize()
    assert time.time() - t0 < 5.0
    for f in range(F):
        o = c_oracle.pf_run(model, Y, thetas[f], False, .4, N, G=G, resampler=resampler, arith=arith, seed=99, filter_id=3 + f, mu=mu, npop=npop)
        assert (o["X_hist"][0] < 0).any(), "the case is meant to start some particles below zero"
        assert int(r.status.cpu()[f]) == o["collapsed"] == 1
        assert np.array_equal(r.X_hist[f, 0].t().cpu().numpy(), o["X_hist"][0])
    mt = [sem.ModelType.SIR, sem.ModelType.SEIR, None, sem.ModelType.SIR_SUBGROUPS2][model]
    th = np.array(theta) if model < 2 else (np.array(theta[:4]).reshape(2, 2), theta[4])
    out = sem.particle_filter(Y, mt, th, False, .4, N, npop if model == 3 else npop[0], mu if model == 3 else mu[0], seed=99, filter_id=3)
    assert out == (None, None, None)


def test_given_negative_state_never_hangs_the_exchange(sem, c_oracle):
    """A given X0 whose negative counts sit in a HIDDEN column (extension D5) collapses too -- any negative compartment has
    no weight -- and a filter whose particles are all valid next to it in the same launch is untouched."""
    import torch
    N, T = 600, 5
    Y = _truth_Y(1, T, 5, .1, False)
    Y[:, 1] = np.nan
    X0 = np.tile(np.array([[980, 0, 20, 0]], dtype=np.int32), (N, 1))
    X0b = X0.copy(); X0b[7, 1] = -2; X0b[7, 0] = 982
    for x0, want in ((X0, 0), (X0b, 1)):
        cfg = sem.engine.make_pf_config(1, N, T, probs=.1, resampler=1, arith=3, seed=5, mu=[20], n_population=[1000])
        r = sem.engine.run_pf(cfg, Y, np.array([4.0, 1.0, 1.0]), X0=x0)
        o = c_oracle.pf_run(1, Y, [4.0, 1.0, 1.0], False, .1, N, resampler=1, arith=3, seed=5, X0=x0)
        torch.cuda.synchronize()
        assert int(r.status[0]) == o["collapsed"] == want
        if want == 0:
            assert np.array_equal(r.X_hist[0].permute(0, 2, 1).cpu().numpy(), o["X_hist"])


def test_one_filter_of_a_batch_collapses(sem, c_oracle):
    """Filters of a batch are independent: one collapsing (its status = step) leaves the others untouched."""
    import torch
    T = 8
    Y = _truth_Y(0, T, 5, .1, False)
    thetas = np.array([[2.0, 1.0], [0.01, 50.0], [1.8, 0.9]])          # filter 1: everyone recovers at once -> Y_I impossible
    cfg = sem.engine.make_pf_config(0, 800, T, n_filters=3, probs=.1, resampler=1, arith=1, seed=5, filter_id0=0, mu=[20], n_population=[1000])
    r = sem.engine.run_pf(cfg, Y, thetas)
    torch.cuda.synchronize()
    st = r.status.cpu().numpy()
    assert st[0] == 0 and st[2] == 0 and st[1] > 0
    for f in (0, 1, 2):
        o = c_oracle.pf_run(0, Y, thetas[f], False, .1, 800, resampler=1, arith=1, seed=5, filter_id=f, mu=[20], npop=[1000])
        assert o["collapsed"] == st[f]
        if st[f] == 0:
            assert np.array_equal(r.X_hist[f].permute(0, 2, 1).cpu().numpy(), o["X_hist"])
    assert np.isneginf(r.log_zetas[1, -1].item())


def test_seir_hidden_exposed_column(sem, c_oracle):
    """Extension: a NaN column of Y is not weighed -- SEIR with the E compartment hidden (config 3)."""
    import torch
    T = 8
    Y = _truth_Y(1, T, 5, .1, False)
    Yh = Y.copy(); Yh[:, 1] = np.nan
    cfg = sem.engine.make_pf_config(1, 1500, T, probs=.1, resampler=1, arith=1, seed=77, mu=[20], n_population=[1000])
    r = sem.engine.run_pf(cfg, Yh, np.array([4.0, 1.0, 1.0]))
    o = c_oracle.pf_run(1, Yh, [4.0, 1.0, 1.0], False, .1, 1500, resampler=1, arith=1, seed=77, mu=[20], npop=[1000])
    torch.cuda.synchronize()
    assert int(r.status[0]) == 0 == o["collapsed"]
    assert np.array_equal(r.X_hist[0].permute(0, 2, 1).cpu().numpy(), o["X_hist"])
    np.testing.assert_allclose(r.log_zetas[0].cpu().numpy(), o["log_zetas"], rtol=1e-11)
    full = c_oracle.pf_run(1, Y, [4.0, 1.0, 1.0], False, .1, 1500, resampler=1, arith=1, seed=77, mu=[20], npop=[1000])
    assert o["log_zetas"][-1] > full["log_zetas"][-1]                  # dropping a column from the min can only raise the weights
    z, H, A = sem.particle_filter(Yh, sem.ModelType.SEIR, np.array([4.0, 1.0, 1.0]), False, .1, 1500, 1000, 20, seed=77, arith="fast")
    assert H.shape == (T, 1500, 4) and np.allclose(np.log(z), o["log_zetas"], rtol=1e-11)


@pytest.mark.parametrize("model,G,theta,npop,mu,arith", [(0, 1, [2.0, 1.0], [1000], [20], 1), (0, 1, [2.0, 1.0], [1000], [20], 0),
                                                          (1, 1, [4.0, 1.0, 1.0], [1000], [20], 1), (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1),
                                                          (0, 1, [2.0, 1.0], [1000], [20], 3), (1, 1, [4.0, 1.0, 1.0], [1000], [20], 3),
                                                          (0, 1, [2.0, 1.0], [1000], [20], 4), (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 4),
                                                          (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 3)])
def test_single_cooperative_launch_equals_launch_per_step(sem, model, G, theta, npop, mu, arith):
    """The whole-filter cooperative kernel (grid.sync() as the resampling barrier) and the launch-per-step path are
    bit-identical, including a multi-filter batch and the two-row (ping-pong) history."""
    import torch
    T = 7
    Y = _truth_Y(model, T, 5, .1, False, G=G)
    for F, N, hist in [(1, 3000, True), (3, 700, True), (1, 5000, False)]:
        outs = []
        for per_step, grid_barrier in ((False, False), (True, False), (False, True)):   # default (offspring form for F = 1), per step, grid barrier
            cfg = sem.engine.make_pf_config(model, N, T, G=G, n_filters=F, probs=.1, resampler=1, arith=arith, seed=6, filter_id0=2,
                                            mu=mu, n_population=npop, store_history=hist, launch_per_step=per_step, grid_barrier=grid_barrier)
            r = sem.engine.run_pf(cfg, Y, np.tile(np.array(theta, float), (F, 1)))
            torch.cuda.synchronize()
            outs.append((r.launches, r.X_hist.cpu().numpy(), r.ancestry.cpu().numpy(), r.log_zetas.cpu().numpy(), r.n_events.cpu().numpy(),
                         r.status.cpu().numpy()))
        assert outs[0][0] == 1 and outs[1][0] in (T, T + 1) and outs[2][0] == 1      # T steps (+ the weight-table kernel)
        for k in (3, 5):                                                       # grid-barrier kernel == default kernel, everything
            assert np.array_equal(outs[2][k], outs[0][k])
        if hist and not outs[0][5].any():
            assert np.array_equal(outs[2][1], outs[0][1]) and np.array_equal(outs[2][2], outs[0][2]) and np.array_equal(outs[2][4], outs[0][4])
        assert np.array_equal(outs[0][5], outs[1][5])                          # status (0, or the step of the collapse)
        for f in range(F):
            upto = int(outs[0][5][f]) or T                                     # rows before a collapse are defined, later ones are not
            rows = range(upto) if hist else ([] if outs[0][5][f] else range(2))
            for r_ in rows:
                assert np.array_equal(outs[0][1][f, r_], outs[1][1][f, r_]) and np.array_equal(outs[0][2][f, r_], outs[1][2][f, r_]), (F, N, hist, f, r_)
            assert np.array_equal(outs[0][3][f], outs[1][3][f]), (F, N, hist, f)
            if not outs[0][5][f]:
                assert outs[0][4][f] == outs[1][4][f]


@pytest.mark.parametrize("exact", [False, True])
def test_iteration_result_equals_separate_calls(sem, exact):
    """One PMCMC iteration in one call (sem_b200.h: iteration_result): the packed log-likelihood / status / event count /
    sampled trajectory equal the separate particle_filter + particle_path_sampler calls, on the whole-filter kernel and on
    the launch-per-step path, for one and for several filters."""
    import torch
    T, seed = 9, 0xABCDEF
    Y = _truth_Y(0, T, 5, .1, False)
    for F, N, per_step in [(1, 3000, False), (1, 3000, True), (3, 700, False)]:
        cfg = sem.engine.make_pf_config(0, N, T, n_filters=F, probs=.1, resampler=1, arith=3, seed=seed, filter_id0=17, mu=[20],
                                        n_population=[1000], launch_per_step=per_step, path_exact=exact)
        it = torch.full((F, sem.engine.ITER_HEADER + T * 3), -7.0, dtype=torch.float64, device="cuda")
        res = sem.engine.run_pf(cfg, Y, np.tile(np.array([2.0, 1.0]), (F, 1)), iter_out=it)
        torch.cuda.synchronize()
        r = it.cpu().numpy()
        for f in range(F):
            assert r[f, 0] == float(res.log_zetas[f, -1]) and r[f, 1] == float(res.status[f]) == 0 and r[f, 2] == float(res.n_events[f])
            traj = res.path_sample(f, exact=exact, seed=seed, filter_id=17 + f).cpu().numpy()
            assert np.array_equal(r[f, sem.engine.ITER_HEADER:].reshape(T, 3), traj.astype(float))
            assert 0 <= r[f, 3] < N and np.array_equal(traj[-1], res.X_hist[f, -1, :, int(r[f, 3])].cpu().numpy())
    # a collapsing filter reports its status and no trajectory
    Yc = Y.copy(); Yc[3] = [5000.0, 5000.0, 5000.0]
    cfg = sem.engine.make_pf_config(0, 500, T, probs=.1, resampler=1, arith=3, seed=seed, mu=[20], n_population=[1000])
    it = torch.zeros((1, sem.engine.ITER_HEADER + T * 3), dtype=torch.float64, device="cuda")
    res = sem.engine.run_pf(cfg, Yc, np.array([2.0, 1.0]), iter_out=it)
    torch.cuda.synchronize()
    assert int(res.status[0]) == 4 and it[0, 1].item() == 4.0 and it[0, 3].item() == -1.0 and it[0, 0].item() == float("-inf")


@pytest.mark.parametrize("F,N,model,G,theta,npop,mu", [
    (1, 148 * 140, 0, 1, [2.0, 1.0], [1000], [20]),            # 128 main + 12 time-split particles per CTA (one helper group)
    (2, 74 * 140, 0, 1, [2.0, 1.0], [1000], [20]),             # two filters side by side, same layout
    (1, 148 * 190, 1, 1, [4.0, 1.0, 1.0], [1000], [20]),       # 128 + 62: two helper groups, SEIR
    (1, 148 * 161 + 7, 3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20]),   # 128 + 34 (ragged last CTA), 2 subgroups: blocks of 2 events
])
def test_balanced_layout_matches_oracle(sem, c_oracle, F, N, model, G, theta, npop, mu):
    """The whole-filter kernel's scheduler balancing (extra particles time-split between helper warps, continuation handed
    over in shared memory) is a pure scheduling change: states, ancestors, likelihood and event count equal the oracle's
    and the launch-per-step path's."""
    import torch
    T = 6
    Y = _truth_Y(model, T, 5, .1, False, G=G)
    th = np.tile(np.array(theta, float), (F, 1))
    runs = []
    for per_step in (False, True):
        cfg = sem.engine.make_pf_config(model, N, T, G=G, n_filters=F, probs=.1, resampler=1, arith=3, seed=99, filter_id0=4, mu=mu,
                                        n_population=npop, launch_per_step=per_step)
        r = sem.engine.run_pf(cfg, Y, th)
        torch.cuda.synchronize()
        runs.append(r)
    assert runs[0].launches == 1 and runs[1].launches > 1
    assert torch.equal(runs[0].X_hist, runs[1].X_hist) and torch.equal(runs[0].ancestry, runs[1].ancestry)
    assert torch.equal(runs[0].log_zetas, runs[1].log_zetas) and torch.equal(runs[0].n_events, runs[1].n_events)
    for f in range(F):
        ref = c_oracle.pf_run(model, Y, theta, False, .1, N, G=G, resampler=1, arith=3, seed=99, filter_id=4 + f, mu=mu, npop=npop)
        assert np.array_equal(runs[0].X_hist[f].permute(0, 2, 1).cpu().numpy(), ref["X_hist"])
        assert np.array_equal(runs[0].ancestry[f].cpu().numpy(), ref["ancestry"])
        np.testing.assert_allclose(runs[0].log_zetas[f].cpu().numpy(), ref["log_zetas"], rtol=1e-11, atol=1e-11)
        assert int(runs[0].n_events[f]) == ref["n_events"]


@pytest.mark.gpu
@pytest.mark.parametrize("F,N,model,theta,npop,mu,T", [
    (1, 148 * 140, 0, [2.0, 1.0], [1000], [20], 6),             # 5 chunks of 32: 4 main warps + one helper group (4W+1)
    (1, 148 * 170, 0, [2.0, 1.0], [1000], [20], 6),             # 6 chunks: two helper groups (4W+2)
    (1, 148 * 128 - 7, 0, [2.0, 1.0], [1000], [20], 6),         # 4 chunks, no helpers, ragged last CTA
    (1, 148 * 224, 0, [0.9, 0.5], [3000], [12], 8),             # 7 chunks (4W+3): no helpers
    (2, 74 * 190, 1, [4.0, 1.0, 1.0], [1000], [20], 6),         # SEIR, two filters side by side, 6 chunks
    (1, 148 * 140 + 33, 0, [1.1, 1.0], [400], [1.5], 12),       # most particles start or go extinct (nothing to run: K = 0)
])
def test_sorted_layout_matches_oracle(sem, c_oracle, F, N, model, theta, npop, mu, T):
    """The uniformized whole-filter kernel sorts a CTA's particles by the candidate count of their batch, runs them at
    their sorted position (helper legs for the last chunks) and sends the end state home through shared memory: a pure
    scheduling change.  States, ancestors, likelihood and event count equal the launch-per-step path's and the oracle's."""
    import torch
    Y = _truth_Y(model, T, 5, .1, False)
    th = np.tile(np.array(theta, float), (F, 1))
    runs = []
    for per_step in (False, True):
        cfg = sem.engine.make_pf_config(model, N, T, n_filters=F, probs=.1, resampler=1, arith=4, seed=123, filter_id0=2, mu=mu,
                                        n_population=npop, launch_per_step=per_step)
        r = sem.engine.run_pf(cfg, Y, th)
        torch.cuda.synchronize()
        runs.append(r)
    assert runs[0].launches == 1 and runs[1].launches > 1
    assert torch.equal(runs[0].X_hist, runs[1].X_hist) and torch.equal(runs[0].ancestry, runs[1].ancestry)
    assert torch.equal(runs[0].n_events, runs[1].n_events) and torch.equal(runs[0].status, runs[1].status)
    np.testing.assert_allclose(runs[0].log_zetas.cpu().numpy(), runs[1].log_zetas.cpu().numpy(), rtol=1e-12, atol=1e-12)
    for f in range(F):
        ref = c_oracle.pf_run(model, Y, theta, False, .1, N, resampler=1, arith=4, seed=123, filter_id=2 + f, mu=mu, npop=npop)
        assert int(runs[0].status[f]) == ref["collapsed"]
        assert np.array_equal(runs[0].X_hist[f].permute(0, 2, 1).cpu().numpy(), ref["X_hist"])
        assert np.array_equal(runs[0].ancestry[f].cpu().numpy(), ref["ancestry"])
        np.testing.assert_allclose(runs[0].log_zetas[f].cpu().numpy(), ref["log_zetas"], rtol=1e-11, atol=1e-11)
        assert int(runs[0].n_events[f]) == ref["n_events"]


# ------------------------------------------------------------------ ODE data synthesiser

def test_ode_synthesiser_vs_reference_and_oracle(sem):
    """sir / seir / sir_subgroups _simulate_discrete (pmcmc.py:54-113) through sem_ode_daily: rows within 1e-7 of the
    population of the unmodified reference's (odeint), equal to the numpy restatement of the RK4 scheme to rounding, same
    DataFrame columns; and a batch of 1000 parameter sets equals the one-by-one results."""
    from oracle import sem_oracle as so
    g = golden("ode_synth")
    df = sem.pmcmc.sir_simulate_discrete((4800, 20, 0), g["t200"], 2, 1)
    assert list(df.columns) == ["time", "susceptible", "infected", "removed"]
    assert np.abs(df.to_numpy(dtype=float) - g["sir"]).max() < 1e-7 * 4820
    np.testing.assert_allclose(df.to_numpy(dtype=float)[:, 1:], so.ode_daily(0, 1, (4800, 20, 0), [2, 1], g["t200"]), rtol=1e-11, atol=1e-9)
    df = sem.pmcmc.seir_simulate_discrete((4800, 0, 20, 0), g["t200"], 2.0, 1.0, 1.0)
    assert list(df.columns) == ["time", "susceptible", "exposed", "infected", "removed"]
    assert np.abs(df.to_numpy(dtype=float) - g["seir"]).max() < 1e-7 * 4820
    for name, pop, beta, gamma in (("sub2", g["pop2"], g["beta2"], .5), ("sub3", g["pop3"], g["beta3"], .4)):
        df = sem.pmcmc.sir_subgroups_simulate_discrete(pop, g["t200"], beta, gamma)
        assert list(df.columns)[-1] == "time" and list(df.columns)[:3] == ["susceptible0", "infected0", "removed0"]
        assert np.abs(df.to_numpy(dtype=float) - g[name]).max() < 1e-7 * pop.sum()
    df = sem.pmcmc.sir_simulate_discrete((9980, 20, 0), g["t1500"], .4, .2)
    assert np.abs(df.to_numpy(dtype=float) - g["sir_slow"]).max() < 1e-7 * 10000
    rng = np.random.RandomState(0)
    th = np.column_stack([rng.uniform(.2, 3, 1000), rng.uniform(.1, 1, 1000)])
    batch = sem.pmcmc.simulate_discrete_batch(sem.ModelType.SIR, (4800, 20, 0), th, g["t200"])
    assert batch.shape == (1000, 15, 3)
    for k in (0, 17, 999):
        np.testing.assert_allclose(batch[k], so.ode_daily(0, 1, (4800, 20, 0), th[k], g["t200"]), rtol=1e-11, atol=1e-9)
    assert np.allclose(batch.sum(2), 4820.0)
