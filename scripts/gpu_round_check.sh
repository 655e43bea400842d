#!/bin/bash
# One gpurun call that refreshes every piece of evidence the round is judged on, with its own time limits:
#   gpurun --timeout 420 -- 'bash scripts/gpu_round_check.sh r02a'
# writes gpurun_out/<tag>_{gputests.log,smoke.log,bench.json,bench_reference_arm.json,launches.csv,launch_summary.txt}.
# Copy what should be kept into profiles/ afterwards. Nothing here reads /root/reference.
# The launch list skips the launches of the warm-up steps: the count printed by the first (unprofiled) bench run is
# used to find where the last eager step starts.
tag=${1:-check}
out=gpurun_out
mkdir -p $out
echo "== pytest -m gpu"
(timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) | tee $out/${tag}_gputests.log
echo "== smoke"
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee $out/${tag}_smoke.log
echo "== bench (default flags: graph mode, cpu baseline on rank 0)"
timeout 200 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 1500 $out/${tag}_bench.json
echo "== bench --impl reference"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
tail -c 400 $out/${tag}_bench_reference_arm.json
echo "== ncu launch list of one eager step (numbers under ncu are never bench values)"
# bench --mode eager runs ~10 identical eager steps of ~160 launches each (warm-up, timed, end-to-end) after ~80 start-up
# launches: a window of 222 launches (2 x 111) starting at 900 is two whole steady-state steps (check "n=" in the summary: every
# per-step count doubled)
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 222 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mode eager > $out/${tag}_launches_bench.log 2>&1
python scripts/summarize_launches.py $out/${tag}_launches.csv 1.0 | tee $out/${tag}_launch_summary.txt | head -30
