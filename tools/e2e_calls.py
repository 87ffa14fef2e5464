import os, sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import sem_b200, workloads
w = workloads.HEADLINE; Y = workloads.headline_Y(); th = list(w["theta"])
def call(n):
    st = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, th, 1e-6, n_chains=n, probs=.1, n_particles=100000, n_population=10000, mu=20, seed=3, stats=st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"n_chains={n}: total {1e3*dt:.1f} ms, {1e3*dt/st['filter_runs']:.3f} ms per iteration ({st['filter_runs']} runs)", flush=True)
call(3); call(3); call(50); call(100); call(100)
pr = cProfile.Profile(); pr.enable(); call(100); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
