"""Multi-GPU check (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

1. the particle-sharded filter over NCCL (all-gather of weight summaries + all-to-all-v particle migration) equals the
   single-GPU filter of the same size and seed, and the C oracle;
2. abc_algo with the trial ids split across ranks returns the single-process accepted set.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sem_b200  # noqa: E402
import workloads  # noqa: E402
from sem_b200 import sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, T, pop = 40_001, 12, 1000
    Y = workloads.observe_binomial(workloads.sir_truth((pop - 20, 20, 0), T, 2.0, 1.0), .1, seed=1)
    out = sharded.run_distributed(Y, 0, np.array([2.0, 1.0]), N, probs=.1, seed=777, filter_id=2, mu=[20], n_population=[pop])
    assert out["collapsed"] == 0
    sh = out["shard"]
    cfg = sem_b200.engine.make_pf_config(0, N, T, probs=.1, resampler=1, seed=777, filter_id0=2, mu=[20], n_population=[pop])
    one = sem_b200.engine.run_pf(cfg, Y, np.array([2.0, 1.0]))
    torch.cuda.synchronize()
    lo, cnt = sh.j0, sh.n_local
    assert torch.equal(one.X_hist[0][:, :, lo:lo + cnt], sh.X_hist), "sharded states differ from the single-GPU filter"
    assert torch.equal(one.ancestry[0][:, lo:lo + cnt], sh.ancestry), "sharded ancestors differ"
    np.testing.assert_allclose(out["log_zetas"], one.log_zetas[0].cpu().numpy(), rtol=1e-11)
    ev = torch.tensor([sh.n_events], dtype=torch.int64, device="cuda")
    dist.all_reduce(ev)
    assert int(ev) == int(one.n_events[0])
    if rank == 0:
        from oracle import c_oracle as co
        ref = co.pf_run(0, Y, [2.0, 1.0], False, .1, N, resampler=1, arith=cfg.arith, seed=777, filter_id=2, mu=[20], npop=[pop])
        assert np.array_equal(one.ancestry[0].cpu().numpy(), ref["ancestry"])
        print(f"sharded filter over {world} GPUs == single GPU == oracle: OK  logZ={out['log_zetas'][-1]:.6f}")
    # the same filter with the exchange on the device (peer memory, one cooperative launch per rank and pass)
    n_local, blk = 20_000, 160
    for model, th, kw in [(0, np.array([.4, .2]), dict(mu=[20], n_population=[pop])),
                          (0, np.array([2.0, 1.0]), dict(mu=[20], n_population=[pop]))]:
        pf = None
        for k in range(2):                                          # two passes through the same arenas (no reset in between)
            o = sharded.run_peer_distributed(Y, model, th, n_local * world, probs=.1, seed=777, filter_id=2 + k, want_path=True, pf=pf,
                                             block_particles=blk, **kw)
            pf = o["shard"]
            assert o["collapsed"] == 0, o["collapsed"]
            cfg = sem_b200.engine.make_pf_config(model, n_local * world, T, probs=.1, resampler=1, arith=pf.cfg.arith, seed=777,
                                                 filter_id0=2 + k, block_particles=blk, **kw)
            it = torch.empty((1, sem_b200.engine.ITER_HEADER + T * 3), dtype=torch.float64, device="cuda")
            one = sem_b200.engine.run_pf(cfg, Y, th, iter_out=it)
            torch.cuda.synchronize()
            lo = rank * n_local
            assert torch.equal(one.X_hist[0][:, :, lo:lo + n_local], pf.X_hist), "device exchange: states differ from the single-GPU filter"
            assert torch.equal(one.ancestry[0][:, lo:lo + n_local], pf.ancestry), "device exchange: ancestors differ"
            np.testing.assert_allclose(o["log_zetas"], one.log_zetas[0].cpu().numpy(), rtol=1e-12)
            ito = it[0].cpu().numpy()
            assert o["iteration"][3] == ito[3] and np.array_equal(o["iteration"][4:], ito[4:]), "path sample over the shards differs"
        dist.barrier()
        pf.close()
        if rank == 0:
            print(f"device-side exchange over {world} GPUs (theta {th.tolist()}, arith {pf.cfg.arith}) == single GPU, incl. path sample: OK")
    # ABC across ranks
    obs = workloads.observe_normal(workloads.sir_truth((480, 20, 0), 10, 2.0, 1.0), .1, seed=7)
    st = {}
    post, traj = sem_b200.abc_algo.abc_algo(obs, 6, 45.0, {"beta": [0, 5], "gamma": [0, 5]}, seed=11, batch=8192, stats=st)
    if rank == 0:
        from oracle import c_oracle as co
        ref = co.abc_trials(obs, st["trials"], 45.0, (0, 5, 0, 5), arith=3, seed=11, trial0=0, want_traj=False)
        acc = np.nonzero(ref["distance"] <= 45.0)[0][:6]
        assert np.array_equal(st["accepted_ids"], acc) and np.array_equal(np.array(post["beta"]), ref["theta"][acc, 0])
        print(f"ABC sharded over {world} GPUs: accepted ids {list(acc)} OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
