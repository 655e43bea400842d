out=gpurun_out; tag=${1:-r04zz}
echo "== dist tests"; (timeout 300 python -m pytest tests/test_gpu_dist_nccl.py -x -q 2>&1 | tail -4) | tee $out/${tag}_disttests.log
echo "== l1 / train tests"; (timeout 300 python -m pytest tests/test_gpu_l1.py tests/test_gpu_train.py tests/test_gpu_graph.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4) | tee $out/${tag}_tests.log
timeout 200 python bench.py --no-cpu-baseline --no-extra > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 260 $out/${tag}_bench.json; echo
