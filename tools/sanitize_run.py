"""Small whole-filter runs for compute-sanitizer (memcheck / racecheck / synccheck): the balanced layout of the direct
method (helper warps, named barriers, shared-memory hand-over), the sorted layout of the uniformized loop (counting sort
and exchange in shared memory; 4W+1 / 4W+2 chunks with helper legs, 4W chunks without), the weight table and the
iteration epilogue are all active."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine
T, N = 6, 148 * 140                      # 140 particles per CTA: 128 main + 12 time-split
Y = workloads.observe_binomial(workloads.sir_truth((980, 20, 0), T, 2.0, 1.0), .1, seed=1)
for model, theta, npop, mu, Yc, arith, N in [(0, [2.0, 1.0], [1000], [20], Y, "fast32", N), (0, [2.0, 1.0], [1000], [20], Y, "uniformized32", N),
                                            (0, [2.0, 1.0], [1000], [20], Y, "uniformized32", 148 * 170),
                                            (0, [2.0, 1.0], [1000], [20], Y, "uniformized32", 148 * 128 - 7)]:
    cfg = engine.make_pf_config(model, N, T, probs=.1, seed=5, mu=mu, n_population=npop, arith=arith)
    it = torch.zeros((1, engine.ITER_HEADER + T * 3), dtype=torch.float64, device="cuda")
    res = engine.run_pf(cfg, Yc, np.array(theta), iter_out=it)
    torch.cuda.synchronize()
    print(arith, N, "launches", res.launches, "status", int(res.status[0]), "logZ", float(res.log_zetas[0, -1]), "events", int(res.n_events[0]))
out = engine.abc_trials(workloads.observe_normal(workloads.sir_truth((480, 20, 0), 8, 2.0, 1.0), .1, seed=7), 2048, 45.0, [0, 5, 0, 5], seed=9)
torch.cuda.synchronize()
print("abc ok", int(out["n_events"][0]))
