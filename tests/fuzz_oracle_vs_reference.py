"""Randomised pin of the C oracle to the UNMODIFIED reference (CPU only; needs /root/reference, so it runs in the build
container, not on the GPU box; not collected by pytest):

    python tests/fuzz_oracle_vs_reference.py [cases] [seed] [pf] [abc] [ssa]      (default: all three parts)

Each case draws a random small particle-filter problem -- model, population (down to a handful), parameters, observation
model and its parameter, number of particles and rows, data that may be impossible under the model -- runs the reference's
particle_filter (pmcmc.py:123-233, jobs=1, numpy seeded) and replays the SAME uniform stream through oracle/sem_oracle.c:
trajectories and resampling indices bit for bit, likelihoods to 1e-9 relative, and the same collapse / no-collapse verdict
(including the reference's collapse on a negative initial count).  Then abc_algo (every trial the
reference ran: trajectories, distances, accepted set) and the simulators alone (gillespie_algo.py) with event-by-event logs.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # conftest helpers
from conftest import mt_doubles  # noqa: E402
from oracle import c_oracle as co, ref_harness  # noqa: E402

warnings.filterwarnings("ignore")
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
parts = sys.argv[3:] or ["pf", "abc", "ssa"]
co.build()
bad = 0
stats = dict(pf_full=0, pf_collapsed=0, pf_negative_init=0, pf_reference_raised_on_hidden_negative=0, pf_d2_linear_underflow=0, pf_sub2_hidden_negative=0, ssa=0, ssa_events=0, abc_cases=0, abc_trials=0, abc_accepted=0)
NAMES = ["SIR", "SEIR", "SIR_SUBGROUPS", "SIR_SUBGROUPS2"]


def random_problem():
    model = int(rng.choice([0, 0, 1, 2, 3]))
    G = 2 if model >= 2 else 1
    pop = int(rng.choice([6, 15, 60, 300]))
    i0 = max(1, int(pop * rng.choice([.05, .2, .5])))
    T = int(rng.randint(2, 6))
    if model == 0:
        theta = np.array([rng.uniform(.1, 3), rng.uniform(.1, 1.5)]); th_ref = theta
        npop, mu = pop, i0
        base = np.stack([np.maximum(pop - i0 * (1 + np.arange(T)), 0), i0 + np.arange(T), np.arange(T)], 1)
    elif model == 1:
        theta = np.array([rng.uniform(.1, 3), rng.uniform(.2, 1), rng.uniform(.1, 1.5)]); th_ref = theta
        npop, mu = pop, i0
        base = np.stack([np.maximum(pop - i0 * (1 + np.arange(T)), 0), np.arange(T), i0 + np.arange(T), np.arange(T)], 1)
    else:
        betas = rng.uniform(.1, 3, (2, 2)); gamma = float(rng.uniform(.1, 1))
        theta = np.concatenate([betas.reshape(-1), [gamma]]); th_ref = (betas, gamma)
        npop, mu = [pop // 2 + 1, pop - pop // 2 + 1], [max(i0 // 2, 0), max(i0 // 2, 1)]
        per = np.concatenate([np.stack([np.maximum(npop[g] - mu[g] * (1 + np.arange(T)), 0), mu[g] + np.arange(T), np.arange(T)], 1) for g in range(2)], 1)
        base = per if model == 2 else per[:, :3] + per[:, 3:]
    normal = bool(rng.randint(2))
    probs = float(rng.choice([.05, .1, .3, .5, .9, 1.0])) if not normal else float(rng.choice([.05, .1, .5]))
    if normal:
        Y = np.floor(base * (1 + probs * rng.normal(size=base.shape) * .3))
    else:
        Y = rng.binomial(base.astype(int), probs).astype(float)
    N = int(rng.choice([4, 8, 20, 50]))
    return model, G, theta, th_ref, npop, mu, Y, normal, probs, N


def fuzz_pf(k):
    global bad
    model, G, theta, th_ref, npop, mu, Y, normal, probs, N = random_problem()
    seed = int(rng.randint(0, 2**31 - 1))
    tag = f"pf {k}: {NAMES[model]} N {N} T {len(Y)} npop {npop} mu {mu} normal {normal} probs {probs} theta {np.round(theta, 3).tolist()} seed {seed}"
    try:
        ref = ref_harness.run_particle_filter(Y, NAMES[model], th_ref, normal, probs, N, npop, mu, seed)
    except ValueError as e:                                  # SUBGROUPS2 only: a hidden negative group reaches the simulator
        np.random.seed(seed)
        I0 = [np.random.poisson(m, N) for m in np.atleast_1d(mu)]
        hidden_negative = any((np.atleast_1d(npop)[g] - I0[g] < 0).any() for g in range(len(I0)))
        if not (model == 3 and hidden_negative and "non-negative" in str(e)):
            bad += 1
            print("REFERENCE RAISED", tag, repr(e), flush=True)
        else:                                               # the reference crashed; nothing to compare
            stats["pf_reference_raised_on_hidden_negative"] += 1
        return
    # the reference's initial particles: redo its Poisson draws (pmcmc.py:156-169)
    np.random.seed(seed)
    Cn = 3 if model == 0 else 4 if model == 1 else 6
    X0 = np.zeros((N, Cn), dtype=np.int64)
    for g, m in enumerate(np.atleast_1d(mu)):
        inf = np.random.poisson(m, N)
        X0[:, 3 * g * (model >= 2) + (2 if model == 1 else 1)] = inf
        X0[:, 3 * g * (model >= 2)] = np.atleast_1d(npop)[g] - inf
    negative = bool((X0 < 0).any())
    stats["pf_negative_init"] += int(negative)
    n = 1 << 14
    while True:
        flat = mt_doubles(ref["mt_key"], ref["mt_pos"], n)
        try:
            out = co.pf_run(model, Y, theta, normal, probs, N, G=G, arith=0, X0=X0, flat_u=flat, want_logw=True)
            break
        except RuntimeError:
            n *= 4
    if ref["collapsed"]:
        stats["pf_collapsed"] += 1
        ok = out["collapsed"] != 0
        if not ok:
            # the reference works in the linear domain, where the weights of ALL particles can underflow to 0
            # (sum = 0 -> nan -> ValueError -> None); the log-domain filter goes on with a likelihood of exp(-10^3...10^8)
            lw = out["log_w"][1:]
            if bool((lw.max(axis=1) < -700).any()):
                ok = True
                stats["pf_d2_linear_underflow"] += 1
    elif model == 3 and negative and out["collapsed"] != 0:
        # SUBGROUPS2 weighs the group sums, which can hide a negative group; the reference goes on until
        # such a particle is resampled (then its simulator raises out of the filter); here any negative count collapses
        stats["pf_sub2_hidden_negative"] += 1
        ok = True
    else:
        stats["pf_full"] += 1
        ok = (out["collapsed"] == 0 and np.array_equal(ref["hidden_process"][0], X0.astype(float)) and
              np.array_equal(out["X_hist"], ref["hidden_process"].astype(np.int32)) and
              np.array_equal(out["ancestry"], ref["ancestry_matrix"].astype(np.int32)))
        if ok:
            z = ref["zetas"]
            pos = z > 0
            ok = np.allclose(out["log_zetas"][pos], np.log(z[pos]), rtol=1e-9, atol=1e-12) and bool(np.all(out["log_zetas"][~pos] < -700))
    if not ok:
        bad += 1
        print("MISMATCH", tag, "reference collapsed", ref["collapsed"], "oracle", out["collapsed"], "negative init", negative, flush=True)


def fuzz_ssa(k):
    global bad
    model = int(rng.choice([0, 1, 2]))
    pop = int(rng.choice([5, 40, 300]))
    tmax = float(rng.choice([.5, 1.0, 4.0]))
    seed = int(rng.randint(0, 2**31 - 1))
    if model == 0:
        x0 = [int(rng.randint(0, pop + 1)), int(rng.randint(0, max(2, pop // 5))), int(rng.randint(0, 5))]
        theta = np.array([rng.uniform(.1, 3), rng.uniform(.1, 1.5)]); th = theta; name = "sir"; G = 1
    elif model == 1:
        x0 = [int(rng.randint(0, pop + 1)), int(rng.randint(0, 5)), int(rng.randint(0, max(2, pop // 5))), 0]
        theta = np.array([rng.uniform(.1, 3), rng.uniform(.2, 1), rng.uniform(.1, 1.5)]); th = theta; name = "seir"; G = 1
    else:
        G = int(rng.choice([2, 3]))
        x0 = [[int(rng.randint(0, pop + 1)), int(rng.randint(0, max(2, pop // 5))), 0] for _ in range(G)]
        betas = rng.uniform(.1, 3, (G, G)); gamma = float(rng.uniform(.1, 1))
        theta = np.concatenate([betas.reshape(-1), [gamma]]); th = (betas, gamma); name = "subgroups"
    tag = f"ssa {k}: {name} x0 {x0} tmax {tmax} theta {np.round(theta, 3).tolist()} seed {seed}"
    ref = ref_harness.run_ssa(name, x0, th, tmax, seed)
    n_ev = ref["times"].size - 1
    u = mt_doubles(ref["mt_key"], ref["mt_pos"], 2 * (n_ev + 2))
    out = co.ssa(model, G, np.array(x0).reshape(-1), theta, tmax, arith=0, u=u, max_rec=n_ev + 8)
    stats["ssa"] += 1; stats["ssa_events"] += n_ev
    if not (out["n_rec"] == n_ev + 1 and np.array_equal(out["times"], ref["times"]) and np.array_equal(out["states"], ref["states"])):
        bad += 1
        print("MISMATCH", tag, flush=True)


def fuzz_abc(k):
    """abc_algo (abc_algo.py:17-109): every trial the reference ran -- its prior draw, Poisson-perturbed start and the
    MT19937 state at the simulator's entry are captured -- replayed through the oracle: day-by-day trajectories, distance,
    and the accepted set in the reference's order."""
    global bad
    T = int(rng.randint(3, 10))
    pop = int(rng.choice([20, 80, 400]))
    i0 = max(1, pop // int(rng.choice([5, 20])))
    beta, gamma = float(rng.uniform(.5, 2.5)), float(rng.uniform(.2, 1))
    t = np.arange(T)
    I = np.minimum(i0 * np.exp((beta - gamma) * t * .5), pop * .6); R = np.minimum(gamma * np.cumsum(I), pop * .4)
    truth = np.stack([np.maximum(pop - I - R, 0), I, R], 1)
    obs = truth + rng.normal(0, .1 * truth + 1e-12)
    pri = {"beta": [0, float(rng.uniform(2, 5))], "gamma": [float(rng.choice([0, .1])), float(rng.uniform(1, 3))]}
    thr = float(rng.choice([.08, .15, .3]) * pop)
    want = int(rng.choice([1, 2, 4]))
    seed = int(rng.randint(0, 2**31 - 1))
    tag = f"abc {k}: T {T} pop {pop} priors {pri} threshold {thr} samples {want} seed {seed}"
    import signal

    def too_long(*_):
        raise TimeoutError
    signal.signal(signal.SIGALRM, too_long)
    signal.alarm(45)                                         # the reference samples until `want` trials are accepted
    try:
        ref = ref_harness.run_abc(obs, want, thr, pri, seed)
    except TimeoutError:
        stats["abc_skipped_slow"] = stats.get("abc_skipped_slow", 0) + 1
        return
    finally:
        signal.alarm(0)
    tr = ref["trials"]
    n = len(tr)
    cap = 2 * (3 * int(max(x["n_start"].sum() for x in tr)) + 8)
    us = [mt_doubles(x["mt_key"], x["mt_pos"], cap) for x in tr]
    out = co.abc_trials(obs, n, thr, theta=np.array([x["theta"] for x in tr]), n_start=np.array([x["n_start"] for x in tr]),
                        ssa_u=np.concatenate(us), ssa_off=np.arange(n + 1) * cap)
    stats["abc_cases"] += 1; stats["abc_trials"] += n
    d = np.array([x["distance"] for x in tr])
    acc = out["distance"] <= thr
    stats["abc_accepted"] += int(acc.sum())
    ok = (np.allclose(out["distance"], d, rtol=1e-13, atol=0) and
          np.array_equal(out["traj"][:, :, 1], np.array([x["I_sim"] for x in tr]).astype(np.int32)) and
          np.array_equal(out["traj"][:, :, 2], np.array([x["R_sim"] for x in tr]).astype(np.int32)) and
          np.array_equal(out["theta"][acc, 0], np.array(ref["posterior"]["beta"])) and
          np.array_equal(out["theta"][acc, 1], np.array(ref["posterior"]["gamma"])) and
          np.array_equal(out["traj"][acc], ref["trajectories"][:, :, 1:].astype(np.int32)))
    if not ok:
        bad += 1
        print("MISMATCH", tag, flush=True)


n_abc = max(1, cases // 4)
for k in range(cases if "pf" in parts else 0):
    fuzz_pf(k)
for k in range(n_abc if "abc" in parts else 0):
    fuzz_abc(k)
for k in range(cases if "ssa" in parts else 0):
    fuzz_ssa(k)
print(f"{cases if 'pf' in parts else 0} filters + {n_abc if 'abc' in parts else 0} ABC runs + {cases if 'ssa' in parts else 0} simulations against the unmodified reference: {bad} bad; {stats}")
sys.exit(1 if bad else 0)
