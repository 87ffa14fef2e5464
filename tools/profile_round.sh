#!/bin/bash
# One GPU session that produces the evidence under profiles/ (run through gpurun from the repo root):
#   plain run + ncu metric pass of every hot kernel (tools/profile_kernels.py), one full ncu capture of the headline kernel,
#   the launch list of the default bench, the bench line itself.   usage: bash tools/profile_round.sh <tag>
tag=${1:-r02}
out=gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
python tools/profile_kernels.py > $out/profile_plain.log 2>&1 || { tail -5 $out/profile_plain.log; exit 1; }
SEM_PROFILE_NO_WRITE=1 ncu --metrics $M --clock-control none --csv --log-file $out/profile_metrics.csv \
    -k regex:"pf_persistent|abc_kernel|pf_step|pf_offspring|pf_init|weight_table" python tools/profile_kernels.py > $out/profile_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-e2e --no-abc > $out/${tag}_bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_default_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-abc > $out/${tag}_launches_ncu.log 2>&1
SEM_PROFILE_ONLY=headline_sir ncu --set full --clock-control none --import-source on -k regex:pf_persistent_x --launch-skip 2 -c 1 \
    -f -o $out/${tag}_pf_persistent_x python tools/profile_kernels.py > $out/${tag}_ncu_full.log 2>&1
python bench.py > $out/${tag}_bench.log 2>&1
tail -c 400 $out/${tag}_bench.log
