"""Latency of one PMCMC iteration at BASELINE config 1 size (SIR, pop 1000, 1000 particles, 15 rows)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sem_b200, workloads
Y = workloads.observe_binomial(workloads.sir_truth((980, 20, 0), 15, 2.0, 1.0), .1, seed=0)
for N in (1000, 10000):
    st = {}
    np.random.seed(0)
    sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, [2.0, 1.0], 1e-4, n_chains=20, probs=.1, n_particles=N, n_population=1000, mu=20, seed=1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, [2.0, 1.0], 1e-4, n_chains=400, probs=.1, n_particles=N, n_population=1000, mu=20, seed=2, stats=st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"N={N}: {1e3*dt/st['filter_runs']:.3f} ms per PMCMC iteration ({st['filter_runs']} filter runs), {N*14*st['filter_runs']/dt/1e6:.2f} M particle-steps/s")
    for L, hh in [(1, 1e-2), (4, 1e-2), (8, 1e-2), (16, 1e-2), (1, 6e-2), (8, 6e-2), (16, 6e-2)]:
        st = {}
        np.random.seed(0)
        sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, [2.0, 1.0], hh, n_chains=40, probs=.1, n_particles=N, n_population=1000, mu=20, seed=1, lookahead=L)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        th, _, _ = sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, [2.0, 1.0], hh, n_chains=2000, probs=.1, n_particles=N, n_population=1000, mu=20, seed=2, stats=st, lookahead=L)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"   h={hh:g} lookahead {L:2d} (used {st.get('lookahead', 1)}): {1e3*dt/2000:.3f} ms per PMCMC iteration ({2000/dt:.0f} iterations/s; {st['launches']} launches for 2000 iterations, "
              f"acceptance {sem_b200.helpers.acceptance_rate(th):.2f}, {1e3*dt/st['launches']:.3f} ms per launch)")
    cfg = sem_b200.engine.make_pf_config(0, N, 15, probs=.1, seed=3, mu=[20], n_population=[1000])
    out = sem_b200.engine.alloc_pf_outputs(cfg)
    Yd = torch.from_numpy(Y).cuda(); th = torch.tensor([2.0, 1.0], dtype=torch.float64).cuda()
    for _ in range(5): sem_b200.engine.run_pf(cfg, Yd, th, out=out)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): sem_b200.engine.run_pf(cfg, Yd, th, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"   device-only filter pass: {e0.elapsed_time(e1)/200:.3f} ms")
