"""Join gpurun_out/profile_events.json (plain run) and gpurun_out/profile_metrics.csv (the same script under ncu) into
profiles/<name>.json: per kernel the measured launch duration, executed warp / thread instructions, DRAM bytes and the
events of the profiled pass -- the inputs of bench.py's roofline -- with the sha256 of the library they were measured on.

    python tools/make_kernel_profile.py [profiles/r02_kernel_profile.json]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_kernel_profile.json")
ev = json.load(open(os.path.join(ROOT, "gpurun_out", "profile_events.json")))
rows = []
with open(os.path.join(ROOT, "gpurun_out", "profile_metrics.csv")) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
# long format: one row per (launch ID, metric)
launch = {}
for r in rows:
    d = launch.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
order = [launch[k] for k in sorted(launch)]
out = dict(lib_sha256=ev["lib_sha256"], src_sha256=ev.get("src_sha256"), how="tools/profile_kernels.py under ncu --metrics (cold-cache, serialised launches) + its plain run "
           "(event-timed ms, event counts); third pass of every workload", kernels={})
pos = 0
for k, rec in enumerate(ev["records"]):
    if k == len(ev["records"]) - 1:                          # the launch-per-step workload comes last: all remaining launches, 3 equal passes
        mine = order[pos:]
        per = len(mine) // rec["passes"]
        last = mine[-per:]
        rec["launches_per_pass"] = per
    else:                                                    # whole-filter kernels: one launch per pass
        n = rec["launches_per_pass"] * rec["passes"]
        mine = order[pos:pos + n]
        pos += n
        last = mine[-rec["launches_per_pass"]:]
    s = lambda m: sum(l.get(m, 0.0) for l in last)
    warp_inst, thr_inst = s("smsp__inst_executed.sum"), s("smsp__thread_inst_executed.sum")
    out["kernels"][rec["tag"]] = dict(
        kernel=rec["kernel"], kernels_seen=sorted(set(l["kernel"] for l in last)), launches_per_pass=rec["launches_per_pass"],
        ms_plain=rec["ms"], ncu_duration_ms=s("gpu__time_duration.sum") / 1e6, n_events=rec["n_events"], units=rec["units"], unit=rec["unit"],
        warp_inst=warp_inst, thread_inst=thr_inst, lanes_per_inst=thr_inst / warp_inst if warp_inst else None,
        thread_inst_per_event=thr_inst / rec["n_events"] if rec["n_events"] else None,
        dram_bytes=s("dram__bytes_read.sum") + s("dram__bytes_write.sum"),
        issue_active_pct=(sum(l.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * l.get("gpu__time_duration.sum", 0.0) for l in last)
                          / max(s("gpu__time_duration.sum"), 1e-30)),
        **{k: rec[k] for k in ("arith", "n_particles", "n_obs", "note") if k in rec})
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
