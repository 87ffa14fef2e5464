import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine
def run(name, model, N, T, G, obs, probs, theta, mu, npop, Y, ariths):
    for arith in ariths:
        cfg = engine.make_pf_config(model, N, T, G=G, observations=obs, probs=probs, arith=arith, seed=11, mu=mu, n_population=npop)
        out = engine.alloc_pf_outputs(cfg)
        Yd = torch.from_numpy(Y).cuda(); th = torch.tensor(theta, dtype=torch.float64).cuda()
        for _ in range(2): res = engine.run_pf(cfg, Yd, th, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record()
        for i in range(K):
            cfg.filter_id0 = i + 1
            res = engine.run_pf(cfg, Yd, th, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        ev = int(res.n_events[0])
        print(f"{name:30s} {arith:14s} {ms:9.3f} ms/pass {N*(T-1)/ms/1e3:8.1f} M p-steps/s {ev/ms/1e6:7.1f} G ev/s {ev/(N*(T-1)):8.1f} ev/p-step status {int(res.status[0])} logZ {float(res.log_zetas[0,-1]):.3f} launches {res.launches}", flush=True)
ar = ["fast32", "uniformized32"]
for name, w in workloads.CONFIGS.items():
    run(name, w["model"], w["n_particles"], w["T"], w["G"], w["observations"], w["probs"], w["theta"], w["mu"], w["n_population"], workloads.config_Y(name), ar)
for pop, beta, gamma, T in [(1000, 2.0, 1.0, 15), (1000, 1.0, 0.5, 25), (4820, 2.0, 1.0, 15), (10000, 0.8, 0.2, 40), (100000, 0.4, 0.2, 101)]:
    Y = workloads.observe_binomial(workloads.sir_truth((pop - 20, 20, 0), T, beta, gamma), .1, seed=0)
    for N in (1000, 100000):
        run(f"sir pop{pop} b{beta} g{gamma} T{T} N{N}", 0, N, T, 1, False, .1, (beta, gamma), [20], [pop], Y, ar)
