"""Randomised parity sweep (not collected by pytest): random configurations on the GPU against the C oracle, bit for bit.

    python tests/fuzz_parity.py [cases] [seed] [mode ...]          modes: filter simulate abc peer dropin iteration   (default: all)

filter    particle filter: model, sizes, population, parameters, observation model, interval simulation, resampler, several
          filters per launch, CTA shape (balanced / sorted / helper layouts), kernel (offspring form, grid barrier, launch per
          step), given X0 or Poisson init (incl. S0 < 0), hidden (NaN) columns of Y, the iteration's path sample
          (states, ancestors, event counts, status; log-likelihoods to 1e-11; trajectory of the chosen particle)
simulate  sem_ssa_simulate: final states of batches of runs, per-run parameters and initial states, all interval simulations
abc       sem_abc_run: sampled parameters, trajectories, distances, event count
dropin    sem_b200.particle_filter (the reference's call: host numpy in, float64 numpy out, arith / kernel chosen by `auto`)
iteration engine.PreparedIteration (what particle_mcmc runs): a batch of proposals with per-filter theta and p_obs, several
          launches on the same buffers
peer      the device-side exchange with 2..4 ranks' kernels on this GPU == the single-rank filter (and through it the oracle)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine, sharded
from oracle import c_oracle as co

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
modes = sys.argv[3:] or ["filter", "simulate", "abc", "peer", "dropin", "iteration"]
bad = 0
stats = dict(filters_compared_in_full=0, filters_collapsed=0, peer_compared_in_full=0, peer_collapsed=0, abc_accepted=0, sim_runs=0,
             dropin_compared_in_full=0, dropin_collapsed=0, iterations_compared_in_full=0, iterations_collapsed=0)


def cols(model, G):
    return 3 if model == 0 else 4 if model == 1 else 3 * G


def random_model(models=(0, 0, 1, 3)):
    model = int(rng.choice(models))
    G = 2 if model >= 2 else 1
    if model == 0:
        theta = [float(rng.uniform(.05, 3)), float(rng.uniform(.05, 1.5))]
    elif model == 1:
        theta = [float(rng.uniform(.05, 3)), float(rng.uniform(.1, 1)), float(rng.uniform(.05, 1.5))]
    else:
        theta = [float(v) for v in rng.uniform(.1, 3, 4)] + [float(rng.uniform(.1, 1))]
    return model, G, theta


def truth_and_init(model, theta, pop, i0, T):
    if model >= 2:
        npop = [pop // 2 + 1, pop - pop // 2 + 1]; mu = [max(i0 // 2, 0), max(i0 // 2, 1)]
        y0 = ((npop[0] - mu[0], mu[0], 0), (npop[1] - mu[1], mu[1], 0))
        truth = workloads.subgroups_truth(y0, T, np.array(theta[:4]).reshape(2, 2), theta[4]).reshape(T, 2, 3)
        return npop, mu, (truth.sum(1) if model == 3 else truth.reshape(T, 6))
    if model == 1:
        return [pop], [i0], workloads.seir_truth((pop - i0, 0, i0, 0), T, *theta)
    return [pop], [i0], workloads.sir_truth((pop - i0, i0, 0), T, *theta)


def fuzz_filter(k):
    global bad
    model, G, theta = random_model()
    arith = int(rng.choice([4, 4, 3, 1])) if model < 2 else int(rng.choice([3, 3, 1]))
    N = int(rng.choice([1, 2, 31, 33, 200, 1000, 3000, 7000]))
    T = int(rng.randint(2, 11))
    pop = int(rng.choice([5, 40, 300, 2000, 20000]))
    normal = bool(rng.randint(2))
    probs = float(rng.uniform(.05, .5))
    F = int(rng.choice([1, 1, 2, 5]))
    rs = int(rng.choice([1, 1, 0]))
    block = int(rng.choice([0, 0, 32, 96, 160, 256, 300]))
    kernel = str(rng.choice(["default", "default", "grid_barrier", "per_step"]))
    given_x0 = bool(rng.randint(4) == 0)
    hide = bool(rng.randint(5) == 0)
    exact = bool(rng.randint(2))
    i0 = max(0, int(pop * rng.choice([0, .01, .05, .3])))
    npop, mu, truth = truth_and_init(model, theta, pop, i0, T)
    Y = workloads.observe_normal(truth, probs, seed=k) if normal else workloads.observe_binomial(truth, probs, seed=k)
    Y = np.asarray(Y, dtype=np.float64)
    if hide:
        Y[:, int(rng.randint(Y.shape[1]))] = np.nan
    seed = int(rng.randint(1, 2**31)); fid = int(rng.randint(0, 1000))
    thetas = np.array([theta] * F) * (1 + 0.05 * np.arange(F))[:, None]
    Cn = cols(model, G)
    X0 = None
    if given_x0:                                             # the same initial particles for every filter of the launch
        X0 = np.zeros((N, Cn), dtype=np.int32)
        for g in range(G):
            inf = rng.poisson(max(mu[g], 1), N)
            X0[:, 3 * g * (model >= 2) + (2 if model == 1 else 1)] = inf
            X0[:, 3 * g * (model >= 2)] = npop[g] - inf
    tag = (f"filter {k}: model {model} arith {arith} resampler {rs} N {N} T {T} pop {pop} F {F} normal {normal} block {block} "
           f"kernel {kernel} x0 {given_x0} hide {hide} exact {exact} mu {mu} theta {np.round(theta, 3).tolist()} seed {seed} fid {fid}")
    try:
        cfg = engine.make_pf_config(model, N, T, G=G, n_filters=F, observations=normal, probs=probs, resampler=rs, arith=arith, seed=seed,
                                    filter_id0=fid, mu=mu, n_population=npop, block_particles=block, path_exact=exact,
                                    launch_per_step=kernel == "per_step", grid_barrier=kernel == "grid_barrier")
        it = torch.full((F, engine.ITER_HEADER + T * Cn), -7.0, dtype=torch.float64, device="cuda")
        res = engine.run_pf(cfg, Y, thetas, X0=X0, iter_out=it)
        torch.cuda.synchronize()
        itn = it.cpu().numpy()
        for f in range(F):
            ref = co.pf_run(model, Y, thetas[f], normal, probs, N, G=G, resampler=rs, arith=arith, seed=seed, filter_id=fid + f,
                            X0=X0, mu=None if given_x0 else mu, npop=None if given_x0 else npop)
            st = int(res.status.cpu()[f])
            ok = st == ref["collapsed"] and int(itn[f, 1]) == st
            stats["filters_compared_in_full" if st == 0 else "filters_collapsed"] += 1
            if ok and st == 0:
                ok = (np.array_equal(res.X_hist[f].permute(0, 2, 1).cpu().numpy(), ref["X_hist"]) and
                      np.array_equal(res.ancestry[f].cpu().numpy(), ref["ancestry"]) and
                      np.allclose(res.log_zetas[f].cpu().numpy(), ref["log_zetas"], rtol=1e-11, atol=1e-11) and
                      int(res.n_events.cpu()[f]) == ref["n_events"])
                chosen = int(itn[f, 3])
                ok = ok and 0 <= chosen < N and np.isclose(itn[f, 0], ref["log_zetas"][-1], rtol=1e-11, atol=1e-11)
                if ok:
                    traj = co.path_sample(ref["X_hist"], ref["ancestry"], chosen, exact=exact)
                    ok = np.array_equal(itn[f, engine.ITER_HEADER:].reshape(T, Cn), traj.astype(np.float64))
            if not ok:
                bad += 1
                print("MISMATCH", tag, "filter", f, "status", st, ref["collapsed"], flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e), flush=True)


def fuzz_simulate(k):
    global bad
    model, G, theta = random_model((0, 1, 2))
    arith = int(rng.choice([0, 1, 3, 4, 2])) if model < 2 else int(rng.choice([0, 1, 3]))
    n = int(rng.choice([1, 7, 64, 500]))
    pop = int(rng.choice([5, 60, 800, 5000]))
    tmax = float(rng.choice([.3, 1.0, 2.5, 7.0]))
    seed = int(rng.randint(1, 2**31)); s0 = int(rng.randint(0, 5000))
    Cn = cols(model, G)
    shared = bool(rng.randint(2))
    x0 = np.zeros((n, Cn), dtype=np.int64)
    for g in range(G):
        inf = rng.randint(0, max(2, pop // 10), n)
        x0[:, 3 * g * (model >= 2) + (2 if model == 1 else 1)] = inf
        x0[:, 3 * g * (model >= 2)] = rng.randint(0, pop + 1, n)
    th = np.array([theta] * n) * rng.uniform(.5, 1.5, (n, 1))
    if shared:
        x0, th = x0[0], th[0]
    tag = f"simulate {k}: model {model} arith {arith} n {n} pop {pop} tmax {tmax} shared {shared} seed {seed} sim0 {s0}"
    try:
        out = engine.simulate(model, x0, th, tmax, G=G, arith=arith, seed=seed, sim_index0=s0, n_sims=n)
        x = out["x"].cpu().numpy()
        stats["sim_runs"] += n
        for i in range(n):
            ref = co.ssa(model, G, x0 if shared else x0[i], th if shared else th[i], tmax, arith=arith, seed=seed, sim_index=s0 + i)
            if not np.array_equal(x[i], ref["x"].astype(np.int32)):
                bad += 1
                print("MISMATCH", tag, "run", i, x[i].tolist(), ref["x"].tolist(), flush=True)
                break
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e), flush=True)


def fuzz_abc(k):
    global bad
    T = int(rng.choice([3, 8, 20, 150]))
    pop = int(rng.choice([50, 400, 3000]))
    i0 = max(1, pop // 50)
    beta, gamma = float(rng.uniform(.3, 2.5)), float(rng.uniform(.1, 1))
    obs = np.round(workloads.sir_truth((pop - i0, i0, 0), T, beta, gamma))
    pri = [0.0, float(rng.uniform(1, 5)), float(rng.choice([0.0, .05])), float(rng.uniform(.5, 3))]
    n = int(rng.choice([1, 33, 700, 4000]))
    arith = int(rng.choice([0, 1, 3]))
    seed = int(rng.randint(1, 2**31)); t0 = int(rng.randint(0, 10**6))
    thr = float(rng.choice([0.0, .1, .5, 2.0]) * pop)
    early = bool(rng.randint(2))
    tag = f"abc {k}: T {T} pop {pop} n {n} arith {arith} prior {np.round(pri, 3).tolist()} thr {thr} early {early} seed {seed} trial0 {t0}"
    try:
        out = engine.abc_trials(obs, n, thr, pri, seed=seed, trial0=t0, arith=arith, want_traj=True, early_reject=early)
        ref = co.abc_trials(obs, n, thr, tuple(pri), arith=arith, seed=seed, trial0=t0)
        d, dr = out["distance"].cpu().numpy(), ref["distance"]
        ok = np.array_equal(out["theta"].cpu().numpy(), ref["theta"])
        stats["abc_accepted"] += int((dr <= thr).sum())
        if early:                                            # early rejection: the accepted set and its distances / trajectories
            acc = dr <= thr
            ok = ok and np.array_equal(d <= thr, acc) and np.allclose(d[acc], dr[acc], rtol=1e-13) and \
                np.array_equal(out["traj"].cpu().numpy()[acc], ref["traj"][acc])
        else:
            ok = ok and np.array_equal(out["traj"].cpu().numpy(), ref["traj"]) and np.allclose(d, dr, rtol=1e-13) and \
                int(out["n_events"].cpu()[0]) == ref["n_events"]
        if not ok:
            bad += 1
            print("MISMATCH", tag, flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e), flush=True)


def fuzz_peer(k):
    global bad
    model, G, theta = random_model((0, 0, 1, 3))
    arith = int(rng.choice([4, 3])) if model < 2 else 3
    world = int(rng.choice([2, 2, 3, 4]))
    n_local = int(rng.choice([64, 500, 1024, 2500]))
    block = int(rng.choice([0, 128, 160, 256, 300]))
    T = int(rng.randint(3, 9))
    pop = int(rng.choice([5, 300, 2000]))
    normal = bool(rng.randint(2))
    probs = float(rng.uniform(.05, .5))
    i0 = max(0, int(pop * rng.choice([.01, .05, .3])))
    npop, mu, truth = truth_and_init(model, theta, pop, i0, T)
    Y = workloads.observe_normal(truth, probs, seed=k) if normal else workloads.observe_binomial(truth, probs, seed=k)
    seed = int(rng.randint(1, 2**31)); fid = int(rng.randint(0, 1000))
    N = world * n_local
    Cn = cols(model, G)
    tag = (f"peer {k}: world {world} model {model} arith {arith} n_local {n_local} block {block} T {T} pop {pop} normal {normal} "
           f"mu {mu} theta {np.round(theta, 3).tolist()} seed {seed} fid {fid}")
    try:
        out = sharded.run_peer_local(Y, model, np.array(theta), N, world, G=G, observations=normal, probs=probs, seed=seed, filter_id=fid,
                                     mu=mu, n_population=npop, arith=arith, block_particles=block, want_path=True, timeout_s=5.0)[0]
        cfg = engine.make_pf_config(model, N, T, G=G, observations=normal, probs=probs, resampler=1, arith=arith, seed=seed, filter_id0=fid,
                                    mu=mu, n_population=npop, block_particles=block)
        it = torch.empty((1, engine.ITER_HEADER + T * Cn), dtype=torch.float64, device="cuda")
        one = engine.run_pf(cfg, Y, np.array(theta), iter_out=it)
        torch.cuda.synchronize()
        st = int(one.status[0])
        ok = out["status"] == [st] * world
        stats["peer_compared_in_full" if st == 0 else "peer_collapsed"] += 1
        if ok and st == 0:
            ito = it[0].cpu().numpy()
            ok = (np.array_equal(one.ancestry[0].cpu().numpy(), out["ancestry"]) and np.array_equal(one.X_hist[0].cpu().numpy(), out["X_hist"]) and
                  np.allclose(out["log_zetas"][0], one.log_zetas[0].cpu().numpy(), rtol=1e-12) and out["n_events"] == int(one.n_events[0]) and
                  all(np.array_equal(out["iteration"][r][engine.ITER_HEADER:], ito[engine.ITER_HEADER:]) and out["iteration"][r][3] == ito[3]
                      for r in range(world)))
        if not ok:
            bad += 1
            print("MISMATCH", tag, "status", out["status"], st, flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e), flush=True)


def fuzz_dropin(k):
    global bad
    model, G, theta = random_model((0, 0, 1, 2, 3))
    N = int(rng.choice([8, 100, 1000, 4000]))
    T = int(rng.randint(2, 12))
    pop = int(rng.choice([40, 300, 4820, 20000]))
    normal = bool(rng.randint(2))
    probs = float(rng.uniform(.05, .5))
    rs = str(rng.choice(["systematic", "systematic", "multinomial"]))
    i0 = max(1, int(pop * rng.choice([.005, .05, .2])))
    npop, mu, truth = truth_and_init(model, theta, pop, i0, T)
    Y = workloads.observe_normal(truth, probs, seed=k) if normal else workloads.observe_binomial(truth, probs, seed=k)
    seed = int(rng.randint(1, 2**31)); fid = int(rng.randint(0, 1000))
    mt = [sem_b200.ModelType.SIR, sem_b200.ModelType.SEIR, sem_b200.ModelType.SIR_SUBGROUPS, sem_b200.ModelType.SIR_SUBGROUPS2][model]
    th_in = np.array(theta) if model < 2 else (np.array(theta[:4]).reshape(2, 2), theta[4])      # (betas, gamma) like pmcmc.py:214-215
    tag = f"dropin {k}: model {model} N {N} T {T} pop {pop} normal {normal} resampler {rs} mu {mu} theta {np.round(theta, 3).tolist()} seed {seed} fid {fid}"
    try:
        arith = engine.resolve_arith(model, "auto", theta=np.array(theta), n_particles=N)
        z, H, A = sem_b200.particle_filter(Y, mt, th_in, normal, probs, N, npop if model >= 2 else pop, mu if model >= 2 else mu[0],
                                           resampler=rs, seed=seed, filter_id=fid)
        ref = co.pf_run(model, Y, theta, normal, probs, N, G=G, resampler=int(rs == "systematic"), arith=arith, seed=seed, filter_id=fid, mu=mu, npop=npop)
        stats["dropin_compared_in_full" if z is not None else "dropin_collapsed"] += 1
        ok = (z is None) == (ref["collapsed"] != 0)
        if ok and z is not None:
            ok = (z.dtype == H.dtype == A.dtype == np.float64 and np.array_equal(H, ref["X_hist"].astype(np.float64)) and
                  np.array_equal(A, ref["ancestry"].astype(np.float64)) and np.allclose(z, np.exp(ref["log_zetas"]), rtol=1e-10, atol=0) and
                  np.allclose(np.log(z[z > 0]), ref["log_zetas"][z > 0], rtol=1e-11, atol=1e-11))    # (zetas, like the reference's, underflow to 0)
        if not ok:
            bad += 1
            print("MISMATCH", tag, flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e), flush=True)


def fuzz_iteration(k):
    global bad
    model, G, theta = random_model((0, 0, 1, 3))
    arith = int(rng.choice([4, 3])) if model < 2 else 3
    N = int(rng.choice([64, 1000, 2048, 5000]))
    T = int(rng.randint(3, 10))
    F = int(rng.choice([1, 2, 4, 8, 16]))
    pop = int(rng.choice([300, 2000, 20000]))
    normal = bool(rng.randint(2))
    hist = True                                             # (the packed result needs the history)
    exact = bool(rng.randint(2))
    i0 = max(1, int(pop * rng.choice([.01, .05, .2])))
    npop, mu, truth = truth_and_init(model, theta, pop, i0, T)
    Y = workloads.observe_normal(truth, .2, seed=k) if normal else workloads.observe_binomial(truth, .2, seed=k)
    seed = int(rng.randint(1, 2**31))
    Cn = cols(model, G)
    tag = f"iteration {k}: model {model} arith {arith} N {N} T {T} F {F} pop {pop} normal {normal} history {hist} exact {exact} mu {mu} seed {seed}"
    try:
        cfg = engine.make_pf_config(model, N, T, G=G, n_filters=F, observations=normal, probs=.2, resampler=1, arith=arith, seed=seed,
                                    mu=mu, n_population=npop, store_history=hist, path_exact=exact)
        prep = engine.PreparedIteration(cfg, Y)
        for launch in range(3):
            fid = int(rng.randint(0, 10**6))
            thetas = np.array([theta] * F) * rng.uniform(.8, 1.25, (F, 1))
            per = bool(rng.randint(2))
            pr = rng.uniform(.1, .4, F) if per else float(rng.uniform(.1, .4))
            it = prep.run(thetas, fid, probs=pr).copy()
            for f in range(F):
                ref = co.pf_run(model, Y, thetas[f], normal, pr[f] if per else pr, N, G=G, resampler=1, arith=arith, seed=seed, filter_id=fid + f, mu=mu, npop=npop)
                st = int(it[f, 1])
                stats["iterations_compared_in_full" if st == 0 else "iterations_collapsed"] += 1
                ok = st == ref["collapsed"]
                if ok and st == 0:
                    chosen = int(it[f, 3])
                    ok = np.isclose(it[f, 0], ref["log_zetas"][-1], rtol=1e-11, atol=1e-11) and int(it[f, 2]) == ref["n_events"] and 0 <= chosen < N
                    if ok and hist:
                        ok = np.array_equal(it[f, engine.ITER_HEADER:].reshape(T, Cn), co.path_sample(ref["X_hist"], ref["ancestry"], chosen, exact=exact).astype(np.float64))
                if not ok:
                    bad += 1
                    print("MISMATCH", tag, "launch", launch, "filter", f, "per-filter p_obs", per, "status", st, ref["collapsed"], flush=True)
    except Exception as e:                                   # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e), flush=True)


run = dict(dropin=fuzz_dropin, iteration=fuzz_iteration, filter=fuzz_filter, simulate=fuzz_simulate, abc=fuzz_abc, peer=fuzz_peer)
for m in modes:
    before = bad
    for k in range(cases):
        run[m](k)
    print(f"{m}: {cases} cases, {bad - before} bad", flush=True)
print(f"total: {bad} bad; {stats}")
sys.exit(1 if bad else 0)
