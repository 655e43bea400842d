out=gpurun_out; tag=r02q
for v in "" "DFB_FP32_TC=0"; do
  echo "== smoke $v"; (env $v timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee -a $out/${tag}_smoke.log
done
echo "== l1 + fullsize tests"; (timeout 400 python -m pytest tests/test_gpu_l1.py tests/test_gpu_fullsize.py -q 2>&1 | grep -v "^tests/\|^    \|^$" | tail -60) | tee $out/${tag}_l1.log
echo "== all gpu tests"; (timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -12) | tee $out/${tag}_gputests.log
echo "== bench"; timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json
