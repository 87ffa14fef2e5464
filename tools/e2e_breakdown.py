"""Where the host time of one PMCMC iteration goes at the headline size (device pass 4.5 ms): per-call timers around the
pieces of engine.PreparedIteration.run and of the MH loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine, _lib
w = workloads.HEADLINE
Y = workloads.headline_Y()
cfg = engine.make_pf_config(0, w["n_particles"], w["T"], probs=.1, seed=1, mu=[w["mu"]], n_population=[w["n_population"]], theta=list(w["theta"]))
prep = engine.PreparedIteration(cfg, Y)
th = np.array(w["theta"])
for i in range(3):
    prep.run(th, i, probs=.1, arith=cfg.arith)
T = {}
def tick(k, t0):
    t1 = time.perf_counter(); T[k] = T.get(k, 0.0) + (t1 - t0); return t1
n = 100
stream = torch.cuda.current_stream(prep.dev)
t_all = time.perf_counter()
for i in range(n):
    t = time.perf_counter()
    prop = np.random.multivariate_normal(th, 1e-6 * np.eye(2)); t = tick("mvn", t)
    a = engine.resolve_arith(0, "auto", theta=prop, n_particles=100000); t = tick("resolve", t)
    prep.cfg.filter_id0 = i + 10; prep.th_host[...] = prop; t = tick("stage", t)
    with torch.cuda.stream(stream):
        prep.dev_th.copy_(prep.pin_th, non_blocking=True); t = tick("h2d enqueue", t)
        _lib.check(prep.L.sem_pf_run(C.byref(prep.cfg), C.byref(prep.buf), C.c_void_p(stream.cuda_stream)), "run"); t = tick("sem_pf_run (enqueue)", t)
        prep.pin_it.copy_(prep.dev_it, non_blocking=True); t = tick("d2h enqueue", t)
    stream.synchronize(); t = tick("synchronize (device pass)", t)
    r = prep.it_host[0]; lz = float(r[0]); traj = r[4:].reshape(w["T"], 3).copy(); u = np.log(np.random.uniform()); t = tick("read result", t)
tot = time.perf_counter() - t_all
print(f"{n} iterations: {1e3 * tot / n:.3f} ms per iteration")
for k, v in T.items():
    print(f"  {k:28s} {1e6 * v / n:9.1f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    prep.cfg.filter_id0 = i + 500
    _lib.check(prep.L.sem_pf_run(C.byref(prep.cfg), C.byref(prep.buf), C.c_void_p(stream.cuda_stream)), "run")
e1.record(); torch.cuda.synchronize()
print(f"back-to-back device passes: {e0.elapsed_time(e1) / 20:.3f} ms each")
t0 = time.perf_counter()
for i in range(50):
    prep.cfg.filter_id0 = i + 900
    prep.run(th, i + 900, probs=.1, arith=cfg.arith)
print(f"prep.run: {1e3 * (time.perf_counter() - t0) / 50:.3f} ms each")
np.random.seed(0); st = {}
t0 = time.perf_counter()
sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, list(th), 1e-6, n_chains=100, probs=.1, n_particles=100000, n_population=10000, mu=20, seed=3, stats=st)
print(f"particle_mcmc: {1e3 * (time.perf_counter() - t0) / st['filter_runs']:.3f} ms per iteration")
