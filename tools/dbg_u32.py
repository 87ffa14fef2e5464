import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import sem_b200
from oracle import c_oracle as co
from test_gpu_parity import _truth_Y
model, G, theta, npop, mu, N, T, normal, probs, resampler, arith, block = (3, 2, [5, 2, 1, 3, .5], [415, 620], [15, 20], 1200, 7, False, .1, 1, 4, 32)
Y = _truth_Y(model, T, 5, probs, normal, G=G)
seed = 0xC0FFEE1234
ref = co.pf_run(model, Y, theta, normal, probs, N, G=G, resampler=resampler, arith=arith, seed=seed, filter_id=3, mu=mu, npop=npop)
print("oracle collapsed", ref["collapsed"], ref["log_zetas"])
for per_step in (False, True):
    cfg = sem_b200.engine.make_pf_config(model, N, T, G=G, observations=normal, probs=probs, resampler=resampler, arith=arith,
                                         seed=seed, filter_id0=3, mu=mu, n_population=npop, block_particles=block, launch_per_step=per_step)
    res = sem_b200.engine.run_pf(cfg, Y, np.array(theta, float))
    torch.cuda.synchronize()
    X = res.X_hist[0].permute(0, 2, 1).cpu().numpy(); A = res.ancestry[0].cpu().numpy()
    print("per_step", per_step, "launches", res.launches, "status", int(res.status[0]), res.log_zetas[0].cpu().numpy())
    for p in range(T):
        bad = np.nonzero((X[p] != ref["X_hist"][p]).any(1))[0]
        badA = np.nonzero(A[p] != ref["ancestry"][p])[0]
        if bad.size or badA.size:
            print("  step", p, "state mismatches", bad.size, bad[:8], "ancestor mismatches", badA.size, badA[:8])
            for j in bad[:3]:
                print("    j", j, "gpu", X[p][j], "ref", ref["X_hist"][p][j], "anc gpu", A[p][j], "ref", ref["ancestry"][p][j], "parent", ref["X_hist"][p-1][ref["ancestry"][p][j]])
            break
