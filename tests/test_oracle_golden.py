"""The oracles (oracle/sem_oracle.py, oracle/sem_oracle.c) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden, golden_names, mt_doubles
from oracle import sem_oracle as so

MODEL_OF = {"sir": 0, "seir": 1, "subgroups": 2}


# ---------------------------------------------------------------- SSA (gillespie_algo.py)
@pytest.mark.parametrize("name", golden_names("ssa_"))
def test_ssa_python_oracle_bit_exact(name):
    g = golden(name)
    model = str(g["model"])
    st = so.MTStream(g["mt_key"], g["mt_pos"])
    tmax = float(g["max_time"])
    if model == "sir":
        _, _, t, s = so.ssa_sir(g["population"], g["theta"], tmax, st, True)
    elif model == "seir":
        _, _, t, s = so.ssa_seir(g["population"], g["theta"], tmax, st, True)
    else:
        G = g["population"].shape[0]
        th = g["theta"]
        _, _, t, s = so.ssa_subgroups(g["population"], th[:G * G].reshape(G, G), th[-1], tmax, st, True)
    assert np.array_equal(t, g["times"])
    assert np.array_equal(s.reshape(g["states"].shape), g["states"])


@pytest.mark.parametrize("name", golden_names("ssa_"))
def test_ssa_c_oracle_bit_exact(name, c_oracle):
    g = golden(name)
    model = MODEL_OF[str(g["model"])]
    pop = g["population"]
    G = pop.shape[0] if model == 2 else 1
    n_ev = g["times"].size - 1
    u = mt_doubles(g["mt_key"], g["mt_pos"], 2 * (n_ev + 2))
    out = c_oracle.ssa(model, G, pop, g["theta"], float(g["max_time"]), arith=0, u=u, max_rec=n_ev + 8)
    assert out["n_rec"] == n_ev + 1
    assert np.array_equal(out["times"], g["times"])          # same libm log => bit-equal times
    assert np.array_equal(out["states"], g["states"])
    # FAST arithmetic: same integer path, times to a few ulp
    out2 = c_oracle.ssa(model, G, pop, g["theta"], float(g["max_time"]), arith=1, u=u, max_rec=n_ev + 8)
    assert np.array_equal(out2["states"], g["states"])
    np.testing.assert_allclose(out2["times"], g["times"], rtol=1e-13, atol=0)


# ---------------------------------------------------------------- particle filter (pmcmc.py:123-233)
PF_CASES = [n for n in golden_names("pf_") if n != "pf_sir_collapse"]


@pytest.mark.parametrize("name", PF_CASES)
def test_pf_python_oracle_bit_exact(name):
    g = golden(name)
    if int(g["n_particles"]) > 200:
        pytest.skip("python oracle is slow; covered by the C oracle")
    st = so.MTStream(g["mt_key"], g["mt_pos"])
    out = so.particle_filter(g["Y"], so.MODEL_IDS[str(g["model"])], g["theta"], bool(g["observations"]),
                             float(g["probs"]), g["hidden_process"][0], st)
    assert not out["collapsed"]
    assert np.array_equal(out["hidden_process"], g["hidden_process"])
    assert np.array_equal(out["ancestry_matrix"], g["ancestry_matrix"])
    assert np.array_equal(out["zetas"], g["zetas"])


def run_c_pf_flat(c_oracle, g, arith=0):
    model = so.MODEL_IDS[str(g["model"])]
    N = int(g["n_particles"])
    G = g["mu"].size if model >= 2 else 1
    T = g["Y"].shape[0]
    n = 4096
    while True:
        flat = mt_doubles(g["mt_key"], g["mt_pos"], n)
        try:
            out = c_oracle.pf_run(model, g["Y"], g["theta"], bool(g["observations"]), float(g["probs"]), N, G=G,
                                  arith=arith, X0=g["hidden_process"][0], flat_u=flat, want_logw=True)
            return out, flat
        except RuntimeError:
            n *= 4


@pytest.mark.parametrize("name", PF_CASES)
def test_pf_c_oracle_vs_reference(name, c_oracle):
    g = golden(name)
    out, flat = run_c_pf_flat(c_oracle, g)
    assert out["collapsed"] == 0
    assert np.array_equal(out["X_hist"], g["hidden_process"].astype(np.int32))       # trajectories bit-exact
    assert np.array_equal(out["ancestry"], g["ancestry_matrix"].astype(np.int32))    # resampling indices bit-exact
    np.testing.assert_allclose(out["log_zetas"], np.log(g["zetas"]), rtol=1e-9, atol=1e-12)
    # per-particle CSR replay reproduces the flat replay
    ssa_u, off = c_oracle.flat_to_csr(flat, out["ssa_start"], out["ssa_end"])
    model = so.MODEL_IDS[str(g["model"])]
    G = g["mu"].size if model >= 2 else 1
    out2 = c_oracle.pf_run(model, g["Y"], g["theta"], bool(g["observations"]), float(g["probs"]),
                           int(g["n_particles"]), G=G, X0=g["hidden_process"][0], res_u=out["res_u"], ssa_u=ssa_u,
                           ssa_off=off)
    assert np.array_equal(out2["X_hist"], out["X_hist"]) and np.array_equal(out2["ancestry"], out["ancestry"])
    assert np.array_equal(out2["log_zetas"], out["log_zetas"])


def test_pf_collapse(c_oracle):
    g = golden("pf_sir_collapse")
    N = int(g["n_particles"])
    np.random.seed(int(g["seed"]))
    I0 = np.random.poisson(g["mu"][0], N)                 # pmcmc.py:157
    X0 = np.stack([int(g["n_population"][0]) - I0, I0, 0 * I0], 1)
    flat = mt_doubles(g["mt_key"], g["mt_pos"], 4096)
    out = c_oracle.pf_run(0, g["Y"], g["theta"], False, float(g["probs"]), N, X0=X0, flat_u=flat)
    assert out["collapsed"] == 1                          # reference returns (None,None,None), pmcmc.py:191-192
    st = so.MTStream(g["mt_key"], g["mt_pos"])
    assert so.particle_filter(g["Y"], 0, g["theta"], False, float(g["probs"]), X0, st)["collapsed"]


@pytest.mark.parametrize("name", golden_names("path_"))
def test_path_sampler(name, c_oracle):
    g = golden(name)
    src = golden(str(g["source"]))
    tr = so.particle_path_sampler(src["hidden_process"], src["ancestry_matrix"], int(g["chosen"]))
    assert np.array_equal(tr, g["trajectory"])
    tr_c = c_oracle.path_sample(src["hidden_process"].astype(np.int32), src["ancestry_matrix"].astype(np.int32),
                                int(g["chosen"]))
    assert np.array_equal(tr_c, g["trajectory"].astype(np.int32))


# ---------------------------------------------------------------- observation weights (pmcmc.py:178-181)
def test_weights_known_answers(c_oracle):
    g = golden("weights_known_answers")
    lp = c_oracle.binom_logpmf(g["binom_k"], g["binom_n"], g["binom_p"])
    ref = g["binom_pmf"]
    zero = ref == 0
    assert np.all(np.isneginf(lp[zero & (g["binom_k"] != np.floor(g["binom_k"]))]))
    assert np.all(np.exp(lp[zero]) == 0)                                   # support / underflow agree
    np.testing.assert_allclose(np.exp(lp[~zero]), ref[~zero], rtol=2e-12)   # vs scipy/Boost
    big = ref > 1e-300
    np.testing.assert_allclose(lp[big], np.log(ref[big]), rtol=1e-11, atol=1e-11)
    ln = c_oracle.norm_logpdf(g["norm_y"], g["norm_x"], g["norm_probs"])
    ok = g["norm_pdf"] > 1e-300
    np.testing.assert_allclose(ln[ok], np.log(g["norm_pdf"][ok]), rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------- ABC (abc_algo.py)
def test_abc_python_and_c(c_oracle):
    g = golden("abc_sir_small")
    obs = g["observed"]
    n = g["distance"].size
    streams, offs = [], [0]
    for i in range(n):
        st = so.MTStream(g["mt_key"][i], g["mt_pos"][i])
        tr, d = so.abc_trial(obs, g["theta"][i, 0], g["theta"][i, 1], g["n_start"][i], st)
        assert d == g["distance"][i]
        assert np.array_equal(tr[:, 2], g["I_sim"][i]) and np.array_equal(tr[:, 3], g["R_sim"][i])
        streams.append(np.array(st.log)); offs.append(offs[-1] + len(st.log))
    out = c_oracle.abc_trials(obs, n, float(g["threshold"]), theta=g["theta"], n_start=g["n_start"],
                              ssa_u=np.concatenate(streams), ssa_off=np.array(offs))
    np.testing.assert_allclose(out["distance"], g["distance"], rtol=1e-13)
    assert np.array_equal(out["traj"][:, :, 1], g["I_sim"].astype(np.int32))
    assert np.array_equal(out["traj"][:, :, 2], g["R_sim"].astype(np.int32))
    acc = out["distance"] <= float(g["threshold"])
    assert np.array_equal(out["theta"][acc, 0], g["post_beta"]) and np.array_equal(out["theta"][acc, 1], g["post_gamma"])
    assert np.array_equal(out["traj"][acc], g["trajectories"][:, :, 1:].astype(np.int32))
    assert np.array_equal(g["trajectories"][0, :, 0], np.arange(obs.shape[0]))


# ---------------------------------------------------------------- Philox known answers (Random123 kat_vectors)
def test_philox_known_answers(c_oracle):
    assert c_oracle.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert c_oracle.philox4x32([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert c_oracle.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ---------------------------------------------------------------- statistical pin of the Philox mode (Pattern A)
def test_oracle_philox_logz_distribution_vs_reference(c_oracle):
    """The oracle's free-running (Philox) filter samples the same likelihood-estimate distribution as 300 runs of the
    unmodified reference particle_filter (tests/golden/make_golden_stats.py)."""
    from scipy import stats
    g = golden("stat_logz_sir")
    N, npop, mu = int(g["n_particles"]), int(g["n_population"]), float(g["mu"])
    ref = g["zetas_last"]
    for arith in (0, 1, 2, 3, 4):
        z = np.array([np.exp(c_oracle.pf_run(0, g["Y"], g["theta"], False, float(g["probs"]), N, resampler=0, arith=arith,
                                             seed=1000 + s, mu=[mu], npop=[npop])["log_zetas"][-1]) for s in range(300)])
        se = np.sqrt(ref.var() / ref.size + z.var() / z.size)
        assert abs(z.mean() - ref.mean()) < 4 * se, (arith, z.mean(), ref.mean(), se)      # unbiased for the same Z
        assert stats.ks_2samp(np.log(z), np.log(ref)).pvalue > 1e-3, arith             # same N, multinomial => same law


@pytest.mark.parametrize("mu", [10.0, 37.5, 185.3, 2411.0])
def test_batch_count_sampler_is_poisson(c_oracle, mu):
    """The candidate count of a uniformized32 batch (PTRS with the slow path's logarithms folded, poisson_draw_u) follows
    Poisson(mu): mean, variance and a Kolmogorov-Smirnov distance against scipy's cdf (discrete => conservative)."""
    from scipy import stats
    n = 40000
    k = np.array([c_oracle.poisson_u_philox(mu, 77, i, 3) for i in range(n)])
    assert np.all(k == np.floor(k)) and k.min() >= 0
    assert abs(k.mean() - mu) < 4.5 * np.sqrt(mu / n)
    assert abs(k.var() - mu) < 4.5 * mu * np.sqrt(2.0 / n + 1.0 / (mu * n))
    ks = np.max(np.abs(np.searchsorted(np.sort(k), np.arange(k.min(), k.max() + 1), side="right") / n
                       - stats.poisson.cdf(np.arange(k.min(), k.max() + 1), mu)))
    assert ks < 1.95 / np.sqrt(n), ks                           # alpha ~ 1e-3 for a continuous law


def test_fast32_stream_layout(c_oracle):
    """arith 3 (fast32): event k of a simulation takes words (2(k&1), 2(k&1)+1) of Philox call k>>1 as 32-bit uniforms
    u = w / 2^32; everything else is the FAST arithmetic.  Restated here in plain Python from the raw Philox words."""
    import math
    seed, sim = 0xFEEDFACE12345, 7
    x0, theta, tmax = [480.0, 20.0, 0.0], [2.0, 1.0], 0.75
    out = c_oracle.ssa(0, 1, x0, theta, tmax, arith=3, seed=seed, sim_index=sim, max_rec=4000)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    S, I, R = x0
    N = S + I + R
    bN = theta[0] * (1.0 / N)
    t, k, times, states = 0.0, 0, [0.0], [[S, I, R]]
    while I > 0:
        w = c_oracle.philox4x32([k >> 1, sim, 0, (7 << 24)], key)          # DOM_SIM = 7, filter id 0
        u1, u2 = w[2 * (k & 1)] / 2.0**32, w[2 * (k & 1) + 1] / 2.0**32
        k += 1
        r0, r1 = bN * S * I, theta[1] * I
        a0 = r0 + r1
        tau = -math.log(1.0 - u1) / a0
        if t + tau > tmax:
            break
        t += tau
        if r0 <= u2 * a0:
            I -= 1; R += 1
        else:
            S -= 1; I += 1
        times.append(t); states.append([S, I, R])
    assert out["pairs"] == k and out["n_rec"] == len(times)
    assert np.array_equal(out["states"], np.array(states))
    np.testing.assert_allclose(out["times"], times, rtol=1e-15)
    # and the 52-bit stream of arith 1 is a different stream over the same law: one call per event
    out52 = c_oracle.ssa(0, 1, x0, theta, tmax, arith=1, seed=seed, sim_index=sim, max_rec=4000)
    assert out52["pairs"] != out["pairs"] or not np.array_equal(out52["states"], out["states"])


# ------------------------------------------------------------------ ODE data synthesiser
ODE_CASES = [("sir", 0, 1, (4800, 20, 0), [2, 1], "t200"), ("sir_slow", 0, 1, (9980, 20, 0), [.4, .2], "t1500"),
             ("seir", 1, 1, (4800, 0, 20, 0), [2., 1., 1.], "t200"), ("sir_coarse", 0, 1, (990, 10, 0), [1.2, .4], "tc")]


@pytest.mark.parametrize("name,model,G,y0,th,tkey", ODE_CASES, ids=[c[0] for c in ODE_CASES])
def test_ode_synthesiser_scheme_vs_reference(name, model, G, y0, th, tkey):
    """The RK4 scheme the CUDA synthesiser uses (restated in numpy) against rows made by the unmodified reference
    (odeint + its daily sub-sampling, tests/golden/make_golden_ode.py): within 1e-7 of the population (odeint's own
    tolerance is 1.5e-8 relative)."""
    g = golden("ode_synth")
    rows = so.ode_daily(model, G, y0, th, g[tkey])
    ref = g[name]
    assert np.array_equal(ref[:, 0], np.arange(ref.shape[0]))              # the reference's `time` column = the day
    assert rows.shape == ref[:, 1:].shape
    assert np.abs(rows - ref[:, 1:]).max() < 1e-7 * sum(y0)


def test_ode_synthesiser_subgroups_vs_reference():
    g = golden("ode_synth")
    for name, G, pop, beta, gamma in (("sub2", 2, g["pop2"], g["beta2"], .5), ("sub3", 3, g["pop3"], g["beta3"], .4)):
        rows = so.ode_daily(2, G, pop.reshape(-1), list(beta.reshape(-1)) + [gamma], g["t200"])
        ref = g[name]
        assert np.array_equal(ref[:, -1], np.arange(ref.shape[0]))          # (subgroup frames carry `time` last)
        assert np.abs(rows - ref[:, :-1]).max() < 1e-7 * pop.sum()


def test_daily_rows_of_grid_is_the_reference_selection():
    import sem_b200
    g = golden("ode_synth")
    for tkey in ("t200", "t1500", "tc"):
        t = g[tkey]
        row, n = sem_b200.engine.daily_rows_of_grid(t)
        days = np.ceil(t).astype(int)
        assert n == days[-1] + 1
        for d in range(n):                                                  # pmcmc.py:68-74: last index whose ceil(time) is d
            k = len(days) - 1 - list(days)[::-1].index(d)
            assert row[k] == d
        assert (row >= 0).sum() == n
    with pytest.raises(ValueError):                                         # a day without a grid point: list.index raises
        sem_b200.engine.daily_rows_of_grid(np.array([0.0, 2.5]))


@pytest.mark.parametrize("model,G,theta,npop,mu", [(0, 1, [1.2, .4], [6], [4]), (3, 2, [2.7, 1.7, 1.8, 1.2, .9], [3, 4], [0, 1])])
def test_negative_counts_collapse_the_filter(c_oracle, model, G, theta, npop, mu):
    """S0 = n_population - Poisson(mu) can be negative (pmcmc.py:156-169); scipy has no weight for such
    a particle (nan) and the reference's filter returns (None, None, None) (pmcmc.py:187-192): collapse at the first step."""
    import workloads
    T, N = 4, 1000
    truth = (workloads.subgroups_truth(((npop[0] - mu[0], mu[0], 0), (npop[1] - mu[1], mu[1], 0)), T, np.array(theta[:4]).reshape(2, 2), theta[4]).reshape(T, 2, 3).sum(1)
             if model == 3 else workloads.sir_truth((npop[0] - mu[0], mu[0], 0), T, *theta))
    Y = workloads.observe_binomial(truth, .4, seed=7)
    for arith in (1, 3, 4) if model == 0 else (1, 3):
        o = c_oracle.pf_run(model, Y, theta, False, .4, N, G=G, resampler=1, arith=arith, seed=99, filter_id=3, mu=mu, npop=npop)
        assert (o["X_hist"][0] < 0).any() and o["collapsed"] == 1
    X0 = np.tile(np.array([[2, 4, 0]], dtype=np.int32), (N, 1))          # the same sizes with valid particles only: no collapse
    if model == 0:
        assert c_oracle.pf_run(0, Y, theta, False, .4, N, resampler=1, arith=3, seed=99, X0=X0)["collapsed"] == 0


def test_reference_collapses_on_a_negative_initial_count():
    """The pin of D9: the unmodified reference, run by tests/golden/make_golden_negative.py (numpy seed 1, SIR, population 6,
    mu 4, 1000 particles), returned (None, None, None); scipy's values for a negative n are recorded beside it."""
    g = golden("negative_s0_reference")
    assert bool(g["returned_none"]) and int(g["n_negative_s0"]) > 0
    assert np.isnan(g["binom_pmf_negative_n"]).all() and np.isnan(g["norm_pdf_negative_scale"]).all()
    assert str(g["simulator_error"]).startswith("ValueError") and "non-negative" in str(g["simulator_error"])
