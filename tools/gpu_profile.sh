#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU): launch list of a bench step + one --set full capture of
# the hot kernels.  Outputs land in gpurun_out/ (scratch); the summaries copied to profiles/ are the committed ones.
set -x
mkdir -p gpurun_out
TAG=${1:-r01b}
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --launch-list-only > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"msm_accumulate|ntt_pass_kernel|quot_evaluate_h|sort_scatter|lookup_mark_leftover|perm_num_den|msm_digits|witness_expand|witness_iszero" \
    -c 40 -f -o gpurun_out/${TAG}_full python tools/prof_once.py 22 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2> gpurun_out/${TAG}_full_raw.err
ls -la gpurun_out/${TAG}_full.ncu-rep
# the report itself is too large to bring back with everything else: keep only the CSV
rm -f gpurun_out/${TAG}_full.ncu-rep
wc -l gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_full_raw.csv
tail -3 gpurun_out/${TAG}_full.log
