"""One measured pass of every hot kernel (after two warm-up passes each), for ncu.

    python tools/profile_kernels.py                 # plain run: event-timed ms + event counts -> gpurun_out/profile_events.json
    ncu --metrics <list> --clock-control none --csv --log-file gpurun_out/profile_metrics.csv \
        -k regex:"pf_persistent|abc_kernel|pf_step|pf_offspring|pf_init" python tools/profile_kernels.py

tools/make_kernel_profile.py (run in the container) joins the two files with the hash of the library they were taken
on into profiles/r02_kernel_profile.json, which bench.py reads for its roofline figures.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch  # noqa: E402,E401
import sem_b200, workloads  # noqa: E402,E401
from sem_b200 import engine, sharded, _lib  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "profile_events.json")
records = []


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); r = fn(); e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)


def pf_case(tag, kernel, model, G, theta, T, N, Y, mu, npop, observations=False, probs=.1, grid_barrier=False):
    cfg = engine.make_pf_config(model, N, T, G=G, observations=observations, probs=probs, arith="auto", theta=theta, seed=11, mu=mu,
                                n_population=npop, grid_barrier=grid_barrier)
    out = engine.alloc_pf_outputs(cfg)
    Yd = torch.from_numpy(Y).cuda(); th = torch.tensor(theta, dtype=torch.float64).cuda()
    for i in range(2):
        cfg.filter_id0 = i
        engine.run_pf(cfg, Yd, th, out=out)
    cfg.filter_id0 = 7
    res, ms = timed(lambda: engine.run_pf(cfg, Yd, th, out=out))
    assert int(res.status[0]) == 0
    records.append(dict(tag=tag, kernel=kernel, launches_per_pass=res.launches, passes=3, ms=ms, n_events=int(res.n_events[0]),
                        units=N * (T - 1), unit="particle-steps", arith=engine.ARITH_NAMES[cfg.arith], n_particles=N, n_obs=T))
    print(records[-1], flush=True)


w = workloads.HEADLINE
Yh = workloads.headline_Y()
if os.environ.get("SEM_PROFILE_ONLY") == "headline_sir":     # (the full ncu capture of the headline kernel: nothing else to replay)
    pf_case("headline_sir", "pf_persistent_x<sem::SirModel, 4>", 0, 1, list(w["theta"]), w["T"], w["n_particles"], Yh, [w["mu"]], [w["n_population"]])
    sys.exit(0)
pf_case("headline_sir", "pf_persistent_x<sem::SirModel, 4>", 0, 1, list(w["theta"]), w["T"], w["n_particles"], Yh, [w["mu"]], [w["n_population"]])
pf_case("headline_sir_grid_barrier", "pf_persistent<sem::SirModel, 4>", 0, 1, list(w["theta"]), w["T"], w["n_particles"], Yh, [w["mu"]],
        [w["n_population"]], grid_barrier=True)
for name, kern in [("seir_pop1e4_N1e5_T101", "pf_persistent_x<sem::SeirModel, 4>"), ("sir_noisy_pop1e4_N1e5_T101", "pf_persistent_x<sem::SirModel, 4>"),
                   ("sub2_pop1e4_N1e5_T15", "pf_persistent_x<sem::SubModel<2>, 3>")]:
    c = workloads.CONFIGS[name]
    pf_case(name, kern, c["model"], c["G"], list(c["theta"]), c["T"], c["n_particles"], workloads.config_Y(name), c["mu"], c["n_population"],
            observations=c["observations"], probs=c["probs"])

# ABC rejection trials (config 2)
obs = workloads.observe_normal(workloads.sir_truth((4800, 20, 0), 15, 2.0, 1.0), .1, seed=0)
n = 1 << 20
for i in range(2):
    engine.abc_trials(obs, n, 150.0, [0, 5, 0, 5], seed=5, trial0=i * n)
o, ms = timed(lambda: engine.abc_trials(obs, n, 150.0, [0, 5, 0, 5], seed=5, trial0=5 * n))
records.append(dict(tag="abc_config2", kernel="abc_kernel<3, 0>", launches_per_pass=1, passes=3, ms=ms, n_events=int(o["n_events"].cpu()[0]), units=n,
                    unit="epidemics"))
print(records[-1], flush=True)

# host-driven sharded filter (launch-per-step kernels; config 5 shape at pop 1e4, two shards on this GPU)
c = workloads.CONFIGS["sub2_pop1e4_N1e5_T15"]
Y5 = workloads.config_Y("sub2_pop1e4_N1e5_T15")
for i in range(2):
    sharded.run_local(Y5, 3, np.array(c["theta"], float), 200_000, 2, G=2, probs=.1, seed=4242, filter_id=i, mu=c["mu"], n_population=c["n_population"])
r, ms = timed(lambda: sharded.run_local(Y5, 3, np.array(c["theta"], float), 200_000, 2, G=2, probs=.1, seed=4242, filter_id=9, mu=c["mu"],
                                        n_population=c["n_population"]))
records.append(dict(tag="sharded_host_driven_sub2", kernel="pf_step<sem::SubModel<2>, 3, false>", launches_per_pass=sum(s.launches for s in r["shards"]),
                    passes=3, ms=ms, n_events=sum(s.n_events for s in r["shards"]), units=200_000 * 14, unit="particle-steps",
                    note="wall of the whole host-driven pass (2 shards in lock-step on one GPU), kernels: pf_init, pf_step, pf_offspring"))
print(records[-1], flush=True)

sha = hashlib.sha256(open(_lib.LIB_PATH, "rb").read()).hexdigest()
os.makedirs(os.path.dirname(OUT), exist_ok=True)
if not os.environ.get("SEM_PROFILE_NO_WRITE"):
    from sem_b200 import build as _build
    json.dump(dict(lib_sha256=sha, src_sha256=_build.source_hash(), records=records), open(OUT, "w"), indent=1)
