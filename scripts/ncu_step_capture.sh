#!/bin/bash
# ncu over every launch of one steady-state eager training step of the bench workload (row N3: an ncu capture for every
# kernel that ships), plus `--set full` of the weight-gradient kernels. Reports stay on the box; their raw pages come back
# as CSV (one row per launch):
#   gpurun --timeout 1300 -- 'bash scripts/ncu_step_capture.sh r02r c4'
# Summarise with scripts/summarize_ncu_raw.py. Numbers under ncu are never bench values.
# (`--set full` costs ~10 s per launch here - 17 passes, each restoring the step's device memory - so the whole-step pass
# uses the sections that carry the roofline inputs: 6 passes.)
tag=${1:-check}; cfg=${2:-c4}; skip=${3:-900}; count=${4:-165}
out=gpurun_out
mkdir -p $out
timeout 700 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -s $skip -c $count -f -o /tmp/${tag}_step \
    python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-extra --mode eager > $out/${tag}_ncu_step_bench.log 2>&1
ncu -i /tmp/${tag}_step.ncu-rep --page raw --csv > $out/${tag}_${cfg}_step_raw.csv 2>/dev/null
timeout 400 ncu --set full --import-source on --clock-control none -k regex:tc_ -s 60 -c 24 -f -o /tmp/${tag}_wgrad \
    python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-extra --mode eager > $out/${tag}_ncu_wgrad_bench.log 2>&1
ncu -i /tmp/${tag}_wgrad.ncu-rep --page raw --csv > $out/${tag}_${cfg}_wgrad_full_raw.csv 2>/dev/null
for i in 0 3; do python scripts/ncu_top_stalls.py /tmp/${tag}_wgrad.ncu-rep $i 30 > $out/${tag}_${cfg}_wgrad_stalls_$i.txt 2>&1; done
ls -la /tmp/${tag}_*.ncu-rep $out/${tag}_${cfg}_*
tail -3 $out/${tag}_ncu_step_bench.log
