"""Randomised parity sweep (not collected by pytest): random particle-filter configurations -- model, sizes, population,
parameters, observation model, interval simulation, several filters per launch -- on the GPU against the C oracle,
bit for bit (states, ancestors, event counts; log-likelihoods to 1e-11).

    python tools/fuzz_parity.py [cases] [seed]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine
from oracle import c_oracle as co

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for k in range(cases):
    model = int(rng.choice([0, 0, 1, 3]))
    arith = int(rng.choice([4, 4, 3])) if model < 2 else 3
    N = int(rng.choice([1, 2, 31, 33, 200, 1000, 3000, 7000]))
    T = int(rng.randint(2, 11))
    pop = int(rng.choice([5, 40, 300, 2000, 20000]))
    normal = bool(rng.randint(2))
    probs = float(rng.uniform(.05, .5))
    F = int(rng.choice([1, 1, 2, 5]))
    rs = int(rng.choice([1, 1, 0]))                          # systematic (offspring-form kernel) or multinomial (ancestor search)
    G = 2 if model == 3 else 1
    if model == 0:
        theta = [float(rng.uniform(.05, 3)), float(rng.uniform(.05, 1.5))]
    elif model == 1:
        theta = [float(rng.uniform(.05, 3)), float(rng.uniform(.1, 1)), float(rng.uniform(.05, 1.5))]
    else:
        theta = [float(v) for v in rng.uniform(.1, 3, 4)] + [float(rng.uniform(.1, 1))]
    i0 = max(0, int(pop * rng.choice([0, .01, .05, .3])))
    if model == 3:
        npop = [pop // 2 + 1, pop - pop // 2 + 1]; mu = [max(i0 // 2, 0), max(i0 // 2, 1)]
        y0 = ((npop[0] - mu[0], mu[0], 0), (npop[1] - mu[1], mu[1], 0))
        truth = workloads.subgroups_truth(y0, T, np.array(theta[:4]).reshape(2, 2), theta[4]).reshape(T, 2, 3).sum(1)
    elif model == 1:
        npop = [pop]; mu = [i0]
        truth = workloads.seir_truth((pop - i0, 0, i0, 0), T, *theta)
    else:
        npop = [pop]; mu = [i0]
        truth = workloads.sir_truth((pop - i0, i0, 0), T, *theta)
    Y = workloads.observe_normal(truth, probs, seed=k) if normal else workloads.observe_binomial(truth, probs, seed=k)
    seed = int(rng.randint(1, 2**31)); fid = int(rng.randint(0, 1000))
    thetas = np.array([theta] * F) * (1 + 0.05 * np.arange(F))[:, None]
    tag = f"case {k}: model {model} arith {arith} resampler {rs} N {N} T {T} pop {pop} F {F} normal {normal} mu {mu} theta {np.round(theta, 3).tolist()}"
    try:
        cfg = engine.make_pf_config(model, N, T, G=G, n_filters=F, observations=normal, probs=probs, resampler=rs, arith=arith, seed=seed,
                                    filter_id0=fid, mu=mu, n_population=npop)
        res = engine.run_pf(cfg, Y, thetas)
        torch.cuda.synchronize()
        for f in range(F):
            ref = co.pf_run(model, Y, thetas[f], normal, probs, N, G=G, resampler=rs, arith=arith, seed=seed, filter_id=fid + f, mu=mu, npop=npop)
            st = int(res.status.cpu()[f])
            ok = st == ref["collapsed"]
            if ok and st == 0:
                ok = (np.array_equal(res.X_hist[f].permute(0, 2, 1).cpu().numpy(), ref["X_hist"]) and
                      np.array_equal(res.ancestry[f].cpu().numpy(), ref["ancestry"]) and
                      np.allclose(res.log_zetas[f].cpu().numpy(), ref["log_zetas"], rtol=1e-11, atol=1e-11) and
                      int(res.n_events.cpu()[f]) == ref["n_events"])
            if not ok:
                bad += 1
                print("MISMATCH", tag, "filter", f, "status", st, ref["collapsed"], flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e), flush=True)
print(f"{cases} cases, {bad} bad")
sys.exit(1 if bad else 0)
