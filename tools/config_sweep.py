"""Extra measurements: one filter pass of the other BASELINE configurations (workloads.CONFIGS), device events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine
ariths = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fast32", "fast"]
for name, w in workloads.CONFIGS.items():
    Y = workloads.config_Y(name)
    for arith in ariths:
        cfg = engine.make_pf_config(w["model"], w["n_particles"], w["T"], G=w["G"], observations=w["observations"], probs=w["probs"],
                                    arith=arith, seed=11, mu=w["mu"], n_population=w["n_population"])
        out = engine.alloc_pf_outputs(cfg)
        Yd = torch.from_numpy(Y).cuda(); th = torch.tensor(w["theta"], dtype=torch.float64).cuda()
        for _ in range(2):
            res = engine.run_pf(cfg, Yd, th, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record()
        for i in range(K):
            cfg.filter_id0 = i + 1
            res = engine.run_pf(cfg, Yd, th, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        ev = int(res.n_events[0]); N, T = w["n_particles"], w["T"]
        print(f"{name:28s} {arith:7s} {ms:9.3f} ms/pass  {N*(T-1)/ms/1e3:8.1f} M particle-steps/s  {ev/ms/1e6:7.1f} G events/s  "
              f"{ev/(N*(T-1)):8.1f} ev/p-step  status {int(res.status[0])} logZ {float(res.log_zetas[0,-1]):.3f} launches {res.launches}", flush=True)
