#!/bin/bash
# `ncu --set full` over every launch of one steady-state eager training step of the bench workload (row N3: an ncu capture
# for every kernel that ships). The report stays on the box; its raw page comes back as CSV (one row per launch):
#   gpurun --timeout 600 -- 'bash scripts/ncu_step_capture.sh r02n c4'
# Summarise with scripts/summarize_ncu_raw.py. Numbers under ncu are never bench values.
tag=${1:-check}; cfg=${2:-c4}; skip=${3:-900}; count=${4:-170}
out=gpurun_out
mkdir -p $out
timeout 500 ncu --set full --clock-control none -s $skip -c $count -f -o /tmp/${tag}_step \
    python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-extra --mode eager > $out/${tag}_ncu_step_bench.log 2>&1
ncu -i /tmp/${tag}_step.ncu-rep --page raw --csv > $out/${tag}_${cfg}_step_full_raw.csv 2>/dev/null
ls -la /tmp/${tag}_step.ncu-rep $out/${tag}_${cfg}_step_full_raw.csv
