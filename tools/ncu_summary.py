"""Summarise an ncu report (.ncu-rep) into a small text file for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof_peak.ncu-rep profiles/r01_pf_step_peak.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    lines = [f"ncu summary of {rep}", "(captured with: ncu --set full --clock-control none --import-source on; one launch)", ""]
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        lines.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for m in METRICS:
            if m in hdr:
                lines.append(f"  {m:70s} {r[hdr.index(m)]:>18s} {units[hdr.index(m)]}")
        lines.append("  warp stall reasons (avg warps stalled per issue-active cycle):")
        st = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                v = float(r[hdr.index(h)])
                st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        for v, nm in sorted(st, reverse=True):
            if v >= 0.02:
                lines.append(f"    {nm:28s} {v:8.3f}")
        lines.append("")
    # hottest SASS lines
    rows = ncu_csv(rep, "source", ["--print-source", "sass"])
    hdr = None
    body = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            body.append(r)
    if hdr and body:
        ia, isrc, ins, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        a0 = int(body[0][ia], 16)
        tot_s = sum(int(r[ins]) for r in body) or 1
        tot_e = sum(int(r[iex]) for r in body) or 1
        lines.append(f"SASS: {len(body)} instructions, {tot_s} stall samples, {tot_e} warp-instructions executed")
        agg = {}
        for r in body:
            k = (int(r[ia], 16) - a0) // 0x400
            a = agg.setdefault(k, [0, 0]); a[0] += int(r[ins]); a[1] += int(r[iex])
        lines.append("  share of samples / executed instructions per 1 KiB of code (offset: samples%  inst%):")
        for k in sorted(agg):
            if agg[k][0] > tot_s * 0.02 or agg[k][1] > tot_e * 0.02:
                lines.append(f"    +0x{k * 0x400:05x}: {100 * agg[k][0] / tot_s:5.1f}%  {100 * agg[k][1] / tot_e:5.1f}%")
        lines.append("  hottest instructions (samples, executed, SASS):")
        for r in sorted(body, key=lambda r: -int(r[ins]))[:20]:
            lines.append(f"    +0x{int(r[ia], 16) - a0:05x} {int(r[ins]):6d} {int(r[iex]):10d}  {r[isrc][:90]}")
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
