out=gpurun_out; tag=${1:-r02v}
echo "== smoke"; (timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee $out/${tag}_smoke.log
echo "== all gpu tests"; (timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -40) | tee $out/${tag}_gputests.log
echo "== bench"; timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 400 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
for lag in 0 1 4; do echo "== bench lag $lag"; DEEPFLOWS_SIDE_LAG=$lag timeout 300 python bench.py --no-extra --no-cpu-baseline > $out/${tag}_bench_lag$lag.json 2> $out/${tag}_bench_lag$lag.err; head -c 300 $out/${tag}_bench_lag$lag.json; echo; done
timeout 300 python scripts/step_timeline.py --config c4 2> $out/${tag}_timeline_c4.err | c++filt > $out/${tag}_timeline_c4.txt; tail -3 $out/${tag}_timeline_c4.err; tail -25 $out/${tag}_timeline_c4.txt | cut -c1-150
