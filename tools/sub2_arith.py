import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine
def run(name, model, N, T, G, obs, probs, theta, mu, npop, Y, ariths, K=2):
    for arith in ariths:
        cfg = engine.make_pf_config(model, N, T, G=G, observations=obs, probs=probs, arith=arith, seed=11, mu=mu, n_population=npop)
        out = engine.alloc_pf_outputs(cfg)
        Yd = torch.from_numpy(Y).cuda(); th = torch.tensor(theta, dtype=torch.float64).cuda()
        res = engine.run_pf(cfg, Yd, th, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            cfg.filter_id0 = i + 1
            res = engine.run_pf(cfg, Yd, th, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        ev = int(res.n_events[0])
        print(f"{name:34s} {arith:14s} {ms:10.2f} ms/pass {ev/ms/1e6:7.1f} G ev/s {ev/(N*(T-1)):10.1f} ev/p-step status {int(res.status[0])} logZ {float(res.log_zetas[0,-1]):.3f} launches {res.launches}", flush=True)
T = 15
for pop_scale, N in [(10, 100000), (100, 100000), (100, 20000)]:
    npop = [4000 * pop_scale, 6000 * pop_scale]; mu = [60 * pop_scale, 80 * pop_scale]
    y0 = ((npop[0] - mu[0], mu[0], 0), (npop[1] - mu[1], mu[1], 0))
    theta = (5, 2, 1, 3, .5)
    full = workloads.subgroups_truth(y0, T, np.array(theta[:4]).reshape(2, 2), theta[-1])
    Y = workloads.observe_binomial(full.reshape(T, 2, 3).sum(1), .1, seed=0)
    run(f"sub2 pop{sum(npop)} N{N}", 3, N, T, 2, False, .1, theta, mu, npop, Y, ["fast32", "uniformized32"])
