#!/usr/bin/env python
"""Benchmark of the particle-filter hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the bootstrap particle filter (X_0 init, then T-1 x [weigh, accumulate likelihood,
resample, gather, Gillespie-SSA propagate]) = one PMCMC likelihood evaluation, on the BASELINE.json metric
workload: SIR, population 1e4, 1e5 particles, 100 observation intervals (workloads.HEADLINE).  One pass is one
cooperative kernel launch (pf_persistent).
Metric: particle-steps/s (one particle advanced across one observation interval, including its share of
weighting + resampling).  Prints ONE JSON line (rank 0).

N > 1 (torchrun): ONE filter of N x 1e5 particles sharded over the N ranks (SURVEY 8(e)(3), weak scaling): global
systematic resampling, the resampling barrier and the particle migration run inside the ranks' kernels over peer
memory (NVLink), one cooperative launch per rank and pass (sem_pf_run_sharded).  value = the sharded filter's
particle-steps / max-over-ranks device time.  The figure for N independent filters (independent PMCMC chains,
SURVEY 8(e)(2), no data-path exchange) is kept as the secondary key "independent_chains".

--impl reference: the reference's CPU algorithm for the same path, timed on this box's host cores.  The reference
itself is pure Python and cannot travel to the GPU box, so this arm runs its C restatement (oracle/sem_oracle.c,
pinned bit-exact to the reference's outputs by tests/golden) with all host threads on a bounded sample of the
same workload.  BASELINE.md holds the Python reference's own measured rate (~220 particle-steps/s/core).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import workloads  # noqa: E402

# Roofline inputs.  SURVEY 8(d) declared a budget of 128 thread-instructions per SSA event "to be replaced by ncu
# smsp__inst_executed / events once the kernel exists".  The measured per-kernel figures (executed thread-instructions per
# event, DRAM bytes per launch) are NOT pasted here: they are read from profiles/r02_kernel_profile.json, which
# tools/profile_kernels.py (under ncu) + tools/make_kernel_profile.py write together with the sha256 of the library they
# were measured on and the sha256 of its sources + compiler flags.  If the library loaded now matches neither, the measured roofline is withheld (frac = null,
# "stale_profile") and only the declared-budget figure is printed.
I_ALG_DECLARED = 128
B_ALG = 48             # algorithmic HBM bytes per particle-step, SIR: 8C+24
LANES = 148 * 4 * 32   # issue lanes: SMs x schedulers x warp width
KERNEL_PROFILE = os.path.join(ROOT, "profiles", "r02_kernel_profile.json")


def kernel_profile(tag):
    """(record, stale) of a profiled kernel; record None when the profile file is missing."""
    import hashlib
    try:
        prof = json.load(open(KERNEL_PROFILE))
        from sem_b200 import _lib
        from sem_b200 import build as _build
        sha = hashlib.sha256(open(_lib.LIB_PATH, "rb").read()).hexdigest()
        current = prof["lib_sha256"] == sha or (prof.get("src_sha256") and prof["src_sha256"] == _build.source_hash())
        return prof["kernels"].get(tag), not current
    except Exception:
        return None, True


def workload_config(w, N, T, theta, world, resampler="systematic", arith="uniformized32", exchange="auto"):
    """The `config` object of the JSON line -- the same for the GPU arm and the reference arm (which times a bounded
    sample of this workload; the sample is described in its cpu_baseline.sample)."""
    return {"workload": w["name"], "model": "SIR", "n_particles": N, "n_obs": T, "population": w["n_population"],
            "theta": [float(v) for v in theta], "obs_model": "binomial p=0.1", "resampler": resampler, "arith": arith,
            "l2": "flushed between timed iterations (256 MiB write)", "n_particles_global": world * N,
            "parallelism": (f"one filter of {world}x{N} particles sharded over {world} GPUs: global systematic resampling, "
                            "CTA weight partials + child records exchanged through peer memory inside one cooperative "
                            "launch per rank (no NCCL call, no host work per step)") if world > 1 else
                           ("1 GPU, whole-filter kernel with grid barrier + ancestor search" if exchange == "pull" else
                            "1 GPU, whole-filter kernel, resampling in offspring form (the sharded filter's kernel with one rank)")}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, k): k.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
                     for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") and k not in
                     ("nvmlClocksThrottleReasonNone", "nvmlClocksThrottleReasonAll")}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, nm in names.items():
                        if bit and (r & bit):
                            self.reasons.add(nm)
                except Exception:
                    pass
                time.sleep(0.02)
        except Exception as e:      # NVML unavailable: report it rather than inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


def host_threads():
    # every host thread the process may use, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port(Y, w, n_particles, threads, seed=1):
    """C restatement of the reference (oracle/sem_oracle.c) on the host cores: bounded sample of the workload."""
    from oracle import c_oracle as co
    co.build()
    t0 = time.perf_counter()
    o = co.pf_run(0, Y, list(w["theta"]), w["observations"], w["probs"], n_particles, resampler=1, arith=0, seed=seed,
                  mu=[w["mu"]], npop=[w["n_population"]], n_threads=threads)
    dt = time.perf_counter() - t0
    T = Y.shape[0]
    used = threads if threads > 0 else co.lib().so_num_threads()
    return dict(value=n_particles * (T - 1) / dt, unit="particle-steps/s", cores=used, kind="port",
                sample=f"{n_particles} particles x {T - 1} steps of the same workload, reference fp64 operation order, "
                       f"{o['n_events']} SSA events in {dt:.2f} s",
                events_per_s=o["n_events"] / dt, seconds=dt)


def cpu_reference(Y, w, n_particles, jobs, seed=0):
    """The UNMODIFIED Python reference (oracle/_ref, copied there by oracle/make_ref.py): pmcmc.particle_filter
    (pmcmc.py:123) on a bounded sample of the workload -- fewer particles, every observation interval, same data, same
    theta.  jobs=1: one core (deterministic); jobs=-1: the reference's own per-particle process pool on all host cores."""
    from oracle import ref_run
    if jobs != 1:                                             # loky workers import gillespie_algo by name
        os.environ["PYTHONPATH"] = ref_run.REF + os.pathsep + os.environ.get("PYTHONPATH", "")
    r = ref_run.time_particle_filter(Y, list(w["theta"]), w["observations"], w["probs"], n_particles, w["n_population"], w["mu"],
                                     jobs=jobs, seed=seed)
    cores = 1 if jobs == 1 else host_threads()
    return dict(value=r["particle_steps"] / r["seconds"], unit="particle-steps/s", cores=cores, kind="reference",
                sample=f"unmodified reference pmcmc.particle_filter(jobs={jobs}): {n_particles} particles x {Y.shape[0] - 1} observation "
                       f"intervals of the same workload in {r['seconds']:.1f} s" + (" (filter collapsed)" if r["collapsed"] else ""),
                seconds=r["seconds"], log_likelihood=r["log_z"])


def cpu_baseline(Y, w):
    """cpu_baseline of the GPU arm's line: the Python reference on ONE core (bounded sample, ~20 s), with the C port's
    figure beside it; the C port alone when oracle/_ref did not travel."""
    from oracle import ref_run
    port = cpu_port(Y, w, 5000, 1)
    port.pop("seconds")
    if not ref_run.available():
        port["note"] = "oracle/_ref absent: Python reference not timed here (BASELINE.md: ~220 particle-steps/s/core)"
        return port
    cb = cpu_reference(Y, w, 50, 1)
    cb.pop("seconds")
    cb["c_port_1_thread"] = port
    return cb


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on this box's host cores, all of them (jobs=-1 is the
    reference's process pool, one task per particle per step, pmcmc.py:201-220).  One step = one particle_filter call on a
    bounded sample: 64 particles (not the 1e5 of the GPU arm: the reference needs ~12 h per pass at that size; its cost is
    linear in the particle count) x all 100 observation intervals of the same data.  Falls back to the C port when
    oracle/_ref is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_run
    w = workloads.HEADLINE
    Y = workloads.headline_Y()
    T = Y.shape[0]
    threads = host_threads()
    use_ref = ref_run.available()
    n_sample = 64 if use_ref else 10_000
    run = (lambda n, sd: cpu_reference(Y, w, n, -1, seed=sd)) if use_ref else (lambda n, sd: cpu_port(Y, w, n, threads, seed=sd))
    for i in range(args.warmup):
        run(16 if use_ref else 1000, 100 + i)
    t_tot, last = 0.0, None
    for i in range(args.steps):
        last = run(n_sample, i)
        t_tot += last["seconds"]
    value = args.steps * n_sample * (T - 1) / t_tot
    cb = dict(value=value, unit="particle-steps/s", cores=last["cores"], kind=last["kind"],
              sample=f"each step = {n_sample} particles x {T - 1} observation intervals of the {w['name']} workload, "
                     + ("unmodified Python reference, pmcmc.particle_filter(jobs=-1) = its own process pool over all host cores"
                        if use_ref else "C restatement of the Python reference, all host threads"))
    if use_ref:
        port = cpu_port(Y, w, 10_000, threads)
        port.pop("seconds")
        cb["c_port_all_threads"] = port
    print(json.dumps({
        "impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, w["n_particles"], T, w["theta"], args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "same workload (data, theta, population, horizon, metric) as the GPU arm; the timed steps run a bounded sample "
                f"of {n_sample} particles because the reference's cost is linear in the particle count (about 12 h per pass at 1e5)",
    }))


def run_abc(args):
    """Extra measurement (not the driver's headline): ABC rejection trials of BASELINE config 2
    (tests/test_abc_sir.py: y0=(4800,20,0), beta=2, gamma=1, T=15, priors U(0,5)^2, threshold 150)."""
    import torch
    import torch.distributed as dist
    from sem_b200 import engine
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    obs = workloads.observe_normal(workloads.sir_truth((4800, 20, 0), 15, 2.0, 1.0), .1, seed=0)
    n, K, W = args.trials, args.steps, max(args.warmup, 3)
    for i in range(W):
        engine.abc_trials(obs, n, 150.0, [0, 5, 0, 5], seed=5, trial0=(rank * 1000 + i) * n, arith=args.arith, early_reject=args.early_reject)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    events = acc = 0
    ev0.record()
    outs = []
    for i in range(K):
        outs.append(engine.abc_trials(obs, n, 150.0, [0, 5, 0, 5], seed=5, trial0=(rank * 1000 + W + i) * n, arith=args.arith,
                                      early_reject=args.early_reject))
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    for o in outs:
        events += int(o["n_events"].cpu()[0]); acc += int((o["distance"] <= 150.0).sum().cpu())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.cpu()[0])
    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle as co
            t0 = time.perf_counter()
            o = co.abc_trials(obs, 2000, 150.0, (0, 5, 0, 5), arith=0, seed=5, trial0=0, want_traj=False, n_threads=1)
            dt = time.perf_counter() - t0
            cb = dict(value=2000 / dt, unit="epidemics/s", cores=1, kind="port", sample="2000 trials, reference operation order",
                      events_per_s=o["n_events"] / dt)
        print(json.dumps({"metric": "simulated epidemics/s", "value": world * K * n / (ms / 1e3), "unit": "epidemics/s",
                          "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                          "scaling": "weak", "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "abc_sir_pop4820_T15_prior_U(0,5)^2_thr150", "trials_per_step_per_gpu": n,
                                     "early_reject": bool(args.early_reject), "arith": args.arith},
                          "events_per_s": world * events / (ms / 1e3), "events_per_trial": events / (K * n),
                          "acceptance_rate": acc / (K * n), "gpu_launches": K, "cpu_baseline": cb}))
    if world > 1:
        dist.destroy_process_group()


def run_sharded(args):
    """Extra measurement (not the driver's headline): ONE filter sharded over all ranks (BASELINE config 5 shape:
    2-subgroup SIR, group-summed binomial observations, global systematic resampling, particle migration by
    all-to-all-v).  Weak scaling: N_global = ranks x --particles (default 1e5); population scaled to --population
    (default 1e5: config 5's 1e6 would take ~10 s per pass).  Wall clock with device syncs, max over ranks."""
    import torch
    import torch.distributed as dist
    from sem_b200 import sharded
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29533",
                            rank=rank, world_size=world)
    pop = args.population
    npop = [int(.4 * pop), pop - int(.4 * pop)]
    mu = [max(1, npop[0] // 66), max(1, npop[1] // 75)]
    theta = np.array([5, 2, 1, 3, .5])
    T = 15
    y0 = [(npop[0] - mu[0], mu[0], 0), (npop[1] - mu[1], mu[1], 0)]
    truth = workloads.subgroups_truth(y0, T, theta[:4].reshape(2, 2), theta[4]).reshape(T, 2, 3).sum(1)
    Y = workloads.observe_binomial(truth, .1, seed=0)
    n_local = args.particles or 100_000
    n_global = n_local * world
    K, W = args.steps, max(min(args.warmup, 3), 1)
    times, events = [], 0
    for i in range(W + K):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        out = sharded.run_distributed(Y, 3, theta, n_global, G=2, probs=.1, seed=4242, filter_id=i, mu=mu, n_population=npop,
                                      arith=args.arith, store_history=False)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert out["collapsed"] == 0, "sharded filter collapsed"
        if i >= W:
            times.append(float(dt.cpu()[0]))
            ev = torch.tensor([out["shard"].n_events], dtype=torch.int64, device=dev)
            dist.all_reduce(ev)
            events = int(ev.cpu()[0])
    if rank == 0:
        s_pass = sum(times) / K
        print(json.dumps({"metric": "particle-steps/s", "value": n_global * (T - 1) / s_pass, "unit": "particle-steps/s", "n_gpus": world,
                          "steps": K, "warmup": W, "ms_per_step": 1e3 * s_pass, "higher_is_better": True, "scaling": "weak", "dtype": "f64",
                          "data": "synthetic",
                          "config": {"workload": f"sharded_pf_sub2_pop{pop}_Nlocal{n_local}_T15", "n_global": n_global, "arith": args.arith,
                                     "exchange": "all_gather(2 doubles/rank) + all_to_all_v(int32 records) per step, host-planned"},
                          "events_per_s": events / s_pass, "events_per_particle_step": events / (n_global * (T - 1)),
                          "log_likelihood": float(out["log_zetas"][-1])}))
    dist.destroy_process_group()


def run_config5(args):
    """BASELINE config 5 AS DEFINED (tests/test_pmcmc_sir_subgrps2.py:13-41 scaled to pop 1e6): particle_mcmc on the 2-subgroup
    SIR with group-summed binomial observations, ONE filter of --particles (default 1e7) particles sharded over all ranks with
    global systematic resampling and the path sample over the shards, for a few MH iterations.  A rank's 1.25e6 particles
    are not co-resident on a GPU, so the exchange is the host-driven one (NCCL all-gather + all-to-all-v per step; a step is
    ~1 s of SSA).  Wall clock around the public call, max over ranks."""
    import torch
    import torch.distributed as dist
    import sem_b200
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29534",
                            rank=rank, world_size=world)
    pop = args.population if args.population != 100_000 else 1_000_000
    npop = [int(.4 * pop), pop - int(.4 * pop)]                       # [400000, 600000] at pop 1e6
    mu = [int(.015 * npop[0]), int(8000 / 600000 * npop[1])]          # [6000, 8000]
    theta = np.array([5, 2, 1, 3, .5])
    T = 15
    y0 = [(npop[0] - mu[0], mu[0], 0), (npop[1] - mu[1], mu[1], 0)]
    truth = workloads.subgroups_truth(y0, T, theta[:4].reshape(2, 2), theta[4]).reshape(T, 2, 3).sum(1)
    Y = workloads.observe_binomial(truth, .1, seed=0)
    N = args.particles or 10_000_000
    n_it = max(2, min(args.steps, 4))
    st = {}
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    thetas, lik, trajs = sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR_SUBGROUPS2, list(theta), 1e-8, n_chains=n_it, probs=.1,
                                                n_particles=N, n_population=npop, mu=mu, seed=2026, sharded=True, return_log=True, stats=st)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt.cpu()[0])
    same = torch.tensor(np.concatenate([lik, trajs.reshape(-1)]), device=dev)
    ref = same.clone()
    dist.broadcast(ref, src=0)
    agree = torch.tensor([float(torch.equal(same, ref))], device=dev)
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    if rank == 0:
        runs = st["filter_runs"]
        print(json.dumps({"metric": "particle-steps/s", "value": runs * N * (T - 1) / sec, "unit": "particle-steps/s", "n_gpus": world,
                          "steps": runs, "warmup": 0, "ms_per_step": 1e3 * sec / runs, "higher_is_better": True, "scaling": "strong",
                          "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"config5_pmcmc_sir_subgroups2_pop{pop}_N{N}_T{T}", "n_population": npop, "mu": mu,
                                     "theta": theta.tolist(), "exchange": st["exchange"], "api": "sem_b200.particle_mcmc(sharded=True)",
                                     "parallelism": f"one filter of {N} particles sharded over {world} GPUs, global systematic resampling, "
                                                    "path sample over the shards"},
                          "filter_runs": runs, "seconds_per_likelihood": sec / runs, "log_likelihoods": [float(v) for v in lik],
                          "acceptances": st["acceptances"], "all_ranks_agree": bool(agree.cpu()[0] == 1.0),
                          "sampled_trajectory_last": trajs[:, -1, :].tolist(), "gpu_launches": st["launches"]}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=None, help="override n_particles (not the headline then)")
    ap.add_argument("--resampler", default="systematic")
    ap.add_argument("--arith", default="auto", help="auto = uniformized32 for SIR/SEIR filters, fast32 otherwise (and for ABC)")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--exchange", default="auto", choices=["auto", "pull"],
                    help="N = 1 only: pull = force the whole-filter kernel with grid barrier + ancestor search (pf_persistent); "
                         "auto = sem_pf_run's default (pf_persistent_x with one rank: resampling in offspring form)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-abc", action="store_true", help="skip the secondary ABC measurement (config 2)")
    ap.add_argument("--workload", default="pf", choices=["pf", "abc", "sharded", "config5"],
                    help="pf = the BASELINE metric workload (default); abc = ABC rejection trials (config 2); sharded = one "
                         "particle-sharded filter over all ranks (config 5 shape); the last two are extra measurements")
    ap.add_argument("--population", type=int, default=100_000, help="total population of the sharded workload")
    ap.add_argument("--trials", type=int, default=1 << 20, help="ABC trials per step per GPU")
    ap.add_argument("--early-reject", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    from sem_b200 import engine as _engine            # report the resolved interval simulation, not "auto"
    args.arith = _engine.ARITH_NAMES[_engine.resolve_arith({"pf": 0, "abc": None, "sharded": 3, "config5": 3}[args.workload], args.arith)]
    if args.workload == "config5":
        return run_config5(args)
    if args.workload == "abc":
        return run_abc(args)
    if args.workload == "sharded":
        return run_sharded(args)

    import torch
    import torch.distributed as dist
    import sem_b200
    from sem_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = dict(workloads.HEADLINE)
    if args.particles:
        w["n_particles"] = args.particles
    Y = workloads.headline_Y()
    T, N = Y.shape[0], w["n_particles"]
    theta = np.array(w["theta"], dtype=np.float64)
    K, W = args.steps, max(args.warmup, 3)

    # ---------------------------------------------------------------- device-resident timing ("value")
    cfg = engine.make_pf_config(0, N, T, probs=w["probs"], observations=w["observations"], resampler=args.resampler,
                                arith=args.arith, seed=1234, filter_id0=rank * 4096, mu=[w["mu"]],
                                n_population=[w["n_population"]], block_particles=args.block, grid_barrier=args.exchange == "pull")
    out = engine.alloc_pf_outputs(cfg, dev)
    Yd = torch.from_numpy(Y).to(dev)
    thd = torch.from_numpy(theta).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream()
    sharded_mode = world > 1
    peer = None
    if sharded_mode:                                                      # one filter of world x N particles over the ranks
        from sem_b200 import sharded as shd
        peer = shd.PeerFilter(rank, world, 0, Y, N, theta=theta, probs=w["probs"], observations=w["observations"],
                              arith=args.arith, seed=1234, mu=[w["mu"]], n_population=[w["n_population"]],
                              block_particles=args.block, want_handle=world > 1)
        if world > 1:
            shd.connect_distributed(peer)

    class _Res:                                                           # what the report below reads
        pass

    def one_pass(i):
        if peer is not None:
            peer.run(theta, filter_id=i)
            r = _Res()
            r.n_events, r.status, r.log_zetas, r.launches = peer.out[4], peer.out[3], peer.out[2], 1
            return r
        cfg.filter_id0 = (rank * 4096 + i) & 0xFFFFFF
        return engine.run_pf(cfg, Yd, thd, out=out)

    for i in range(W):
        res = one_pass(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    events_total = 0
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()                                                    # L2 flush between timed iterations (untimed)
        evs[i][0].record(stream)
        res = one_pass(W + i)
        evs[i][1].record(stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    events_last = int(res.n_events.cpu()[0])
    status = int(res.status.cpu()[0])
    logz = float(res.log_zetas[0, -1].cpu())
    assert status == 0, f"filter collapsed at step {status}"
    t_max = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t_max.cpu()[0])
    value = world * K * N * (T - 1) / (dev_ms_max / 1e3)
    events_all = torch.tensor([events_last], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(events_all)                                      # the sharded filter's events are spread over the ranks
    independent = None
    if world > 1:                                                        # secondary: N independent filters (no exchange at all)
        K2 = max(3, min(K, 10))
        cfg.filter_id0 = rank * 4096
        for i in range(2):
            engine.run_pf(cfg, Yd, thd, out=out)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K2):
            cfg.filter_id0 = (rank * 4096 + 100 + i) & 0xFFFFFF
            engine.run_pf(cfg, Yd, thd, out=out)
        e1.record(stream)
        torch.cuda.synchronize()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        independent = {"value": world * K2 * N * (T - 1) / (float(t2.cpu()[0]) / 1e3), "unit": "particle-steps/s", "steps": K2,
                       "note": "N independent filters of 1e5 particles, one per GPU (independent PMCMC chains); L2 not flushed"}

    # ---------------------------------------------------------------- secondary: ABC trials sharded by trial id (config 2)
    # 10^7 simulated epidemics over the ranks (tests/test_abc_sir.py: y0=(4800,20,0), beta=2, gamma=1, T=15, priors
    # U(0,5)^2, threshold 150): every rank runs its own id range, no data-path collective; device events, max over ranks
    abc = None
    if not args.no_abc:
        obs_abc = workloads.observe_normal(workloads.sir_truth((4800, 20, 0), 15, 2.0, 1.0), .1, seed=0)
        n_abc = 10_000_000 // world
        for i in range(2):
            engine.abc_trials(obs_abc, 1 << 18, 150.0, [0, 5, 0, 5], seed=5, trial0=(rank * 100 + i) << 24)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        o_abc = engine.abc_trials(obs_abc, n_abc, 150.0, [0, 5, 0, 5], seed=5, trial0=rank * n_abc)
        a1.record(stream)
        torch.cuda.synchronize()
        ta = torch.tensor([a0.elapsed_time(a1), float(o_abc["n_events"].cpu()[0]), float((o_abc["distance"] <= 150.0).sum().cpu())],
                          dtype=torch.float64, device=dev)
        if world > 1:
            tmax = ta.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(ta, op=dist.ReduceOp.SUM); ta[0] = tmax[0]
        abc = {"metric": "simulated epidemics/s", "value": world * n_abc / (float(ta[0].cpu()) / 1e3), "unit": "epidemics/s",
               "n_trials": world * n_abc, "ms": float(ta[0].cpu()), "events_per_s": float(ta[1].cpu()) / (float(ta[0].cpu()) / 1e3),
               "accepted": int(ta[2].cpu()), "config": "BASELINE config 2: ABC rejection for SIR, 10^7 epidemics sharded by trial id over the ranks",
               "launches": 1}

    # ---------------------------------------------------------------- end-to-end through the public API
    e2e = None
    if not args.no_e2e:
        np.random.seed(rank)
        st = {}
        n_it = max(8, min(4 * K, 120))                                   # MH iterations in the timed call (its one-time setup is inside)
        shard_kw = dict(sharded=True) if world > 1 else {}
        sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, list(theta), 1e-6, n_chains=3, probs=w["probs"], n_particles=world * N,
                               n_population=w["n_population"], mu=w["mu"], seed=77, **shard_kw)          # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        sem_b200.particle_mcmc(Y, sem_b200.ModelType.SIR, list(theta), 1e-6, n_chains=n_it, probs=w["probs"],
                               n_particles=world * N, n_population=w["n_population"], mu=w["mu"], seed=99, stats=st,
                               resampler=args.resampler, arith=args.arith, **shard_kw)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.cpu()[0])
        runs = st["filter_runs"]
        e2e = {"value": world * runs * N * (T - 1) / e2e_s, "unit": "particle-steps/s",
               "h2d_bytes_per_step": int(world * (theta.nbytes + Y.nbytes / runs)), "d2h_bytes_per_step": int(world * 8 * (4 + T * 3)),
               "api": "sem_b200.particle_mcmc (drop-in for pmcmc.py:251): host numpy in, host numpy out, one filter pass + "
                      "path sample per MH iteration" + ("; sharded=True: ONE chain whose filter of world x 1e5 particles is "
                      "sharded over the ranks (setup of the peer arenas inside the timed call)" if world > 1 else ""),
               "iterations": runs, "pmcmc_iters_per_s": (1 if world > 1 else world) * runs / e2e_s}
        # the other public call: particle_filter returning the reference's full (T,N,C)+(T,N) float64 history on the host
        sem_b200.particle_filter(Y, sem_b200.ModelType.SIR, theta, w["observations"], w["probs"], N, w["n_population"], w["mu"],
                                 seed=4, resampler=args.resampler, arith=args.arith)                    # warm-up (staging buffers)
        t0 = time.perf_counter()
        z, H, A = sem_b200.particle_filter(Y, sem_b200.ModelType.SIR, theta, w["observations"], w["probs"], N,
                                           w["n_population"], w["mu"], seed=5, resampler=args.resampler, arith=args.arith)
        th = time.perf_counter() - t0
        e2e["full_history"] = {"value": N * (T - 1) / th, "unit": "particle-steps/s", "d2h_bytes_per_step": int(H.nbytes + A.nbytes + z.nbytes),
                               "api": "sem_b200.particle_filter (pmcmc.py:123) returning float64 numpy history", "per_gpu": True}
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.result()

    if rank == 0:
        hbm_peak, sm_max, how = measured_peaks()
        f_sm = (clocks["sm_mhz"] or sm_max) * 1e6
        per_gpu_ms = dev_ms / K
        events_per_s = float(events_all.cpu()[0]) / world / (per_gpu_ms / 1e3)          # per GPU
        events_last = float(events_all.cpu()[0]) / world
        issue_peak = LANES * f_sm
        launches = res.launches
        default_path = (args.arith == "uniformized32" and args.resampler == "systematic" and N == workloads.HEADLINE["n_particles"]
                        and launches == 1)
        rec, stale = kernel_profile("headline_sir_grid_barrier" if args.exchange == "pull" and world == 1 else "headline_sir") \
            if default_path else (None, True)
        measured = rec is not None and not stale
        i_alg = rec["thread_inst_per_event"] if measured else None
        traffic = rec["dram_bytes"] if measured else None
        kname = "pf_persistent_x" if (sharded_mode or (launches == 1 and args.exchange != "pull" and args.resampler == "systematic")) else \
            ("pf_persistent" if launches == 1 else "pf_step")
        roofline = {"bound": "issue", "kernel": kname + f"<SirModel, {args.arith}>",
                    "achieved": events_per_s * i_alg / 1e9 if measured else None, "peak": issue_peak / 1e9, "unit": "Gthread-inst/s",
                    "frac": events_per_s * i_alg / issue_peak if measured else None,
                    "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu)",
                    "events_per_s": events_per_s,
                    "events_per_particle_step": events_last / (N * (T - 1)), "I_alg": i_alg,
                    "I_alg_source": ("profiles/r02_kernel_profile.json: ncu smsp__thread_inst_executed.sum / SSA events of one launch of this "
                                     "kernel, library sha256 matches the one loaded") if measured else
                                    ("withheld: " + ("no profile of this configuration" if rec is None else
                                                     "stale_profile (profiles/r02_kernel_profile.json was measured on another build of the library)")),
                    "stale_profile": bool(rec is not None and stale),
                    "issue_active_pct_ncu": rec["issue_active_pct"] if measured else None,
                    "active_lanes_per_inst_ncu": rec["lanes_per_inst"] if measured else None,
                    "frac_with_declared_I_alg_128": events_per_s * I_ALG_DECLARED / issue_peak,
                    "frac_at_round1_I_alg_60.6": events_per_s * 60.6 / issue_peak,
                    "sm_mhz_used": f_sm / 1e6, "launches_per_pass": launches, "avg_launch_us": 1e3 * per_gpu_ms / launches,
                    "note": "SSA propagate is bound by SM instruction issue, not HBM (SURVEY 8(d)); peak = 148 SMs x 4 "
                            "schedulers x 32 lanes x SM clock sampled during the run; frac = issue-active x active lanes / 32; sm_100a "
                            "issues the IMAD.WIDE of Philox once per ~4 cycles, so the reachable fraction for this mix is ~0.65 "
                            "(DESIGN.md); events = fired SSA events (uniformized modes) or uniform pairs drawn (direct modes); frac uses the "
                            "MEASURED instructions per event of this build, so removing instructions lowers it while raising "
                            "throughput: frac_at_round1_I_alg_60.6 prices the same events at round 1's instruction count "
                            "(the unit in which the round-1 review set its 0.58 target)"}
        hbm_ach = N * (T - 1) * B_ALG / (per_gpu_ms / 1e3) / 1e9
        roofline_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                        "traffic": traffic, "peak_source": how, "bytes_per_launch": N * (T - 1) * B_ALG / launches,
                        "note": "algorithmic bytes 8C+24 per particle-step; measured DRAM traffic is lower because the "
                                "step's working set (4.8 MB) stays in the 126 MB L2"}
        line = {
            "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, N, T, theta, world, args.resampler, args.arith, args.exchange),
            "independent_chains": independent, "abc_config2": abc,
            "pmcmc_iters_per_s": world * K / (dev_ms_max / 1e3), "events_per_s": events_per_s * world,
            "log_likelihood": logz, "gpu_launches": K * res.launches, "clocks": clocks,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "e2e": e2e, "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(Y, w)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
