out=gpurun_out; tag=${1:-r04s}
echo "== tests"; (timeout 600 python -m pytest tests/test_gpu_l0.py tests/test_gpu_train.py tests/test_gpu_graph.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -8) | tee $out/${tag}_tests.log
echo "== bench"; timeout 200 python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 300 $out/${tag}_bench.json; echo; tail -2 $out/${tag}_bench.err
timeout 200 python scripts/step_timeline.py --config c4 2> $out/${tag}_timeline_c4.err | c++filt > $out/${tag}_timeline_c4.txt; tail -2 $out/${tag}_timeline_c4.err; head -1 $out/${tag}_timeline_c4.txt
