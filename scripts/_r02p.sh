out=gpurun_out; tag=${1:-r02s}
echo "== smoke"; (timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee $out/${tag}_smoke.log
echo "== fused tests"; (timeout 400 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | grep -v "^tests/\|^    \|^$" | tail -40) | tee $out/${tag}_fused.log
echo "== all gpu tests"; (timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -40) | tee $out/${tag}_gputests.log
echo "== bench"; timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
echo "== bench slots off"; DFB_STAT_SLOTS=0 timeout 300 python bench.py --no-extra --no-cpu-baseline > $out/${tag}_bench_noslots.json 2> $out/${tag}_bench_noslots.err; head -c 300 $out/${tag}_bench_noslots.json
