out=gpurun_out; tag=r02o
for v in "" "DEEPFLOWS_FUSE=0" "DFB_CONV_PERSISTENT=0" "DFB_STEM_TC=0" "DFB_FP32_TC=0"; do
  echo "== smoke $v"; (env $v timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee -a $out/${tag}_smoke.log
done
echo "== l1 tests"; (timeout 300 python -m pytest tests/test_gpu_l1.py -q 2>&1 | tail -30) | tee $out/${tag}_l1.log
echo "== all gpu tests"; (timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -30) | tee $out/${tag}_gputests.log
echo "== bench"; timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 1200 $out/${tag}_bench.json
