out=gpurun_out; tag=${1:-r04u}
echo "== tests"; (timeout 300 python -m pytest tests/test_gpu_l1.py -k linear_small -x -q 2>&1 | tail -3)
for f in 1 0; do echo "== bench DFB_LINEAR_SMALL=$f"; DEEPFLOWS_LINEAR_SMALL=$f timeout 200 python bench.py --no-cpu-baseline --no-extra > $out/${tag}_bench_ls$f.json 2> $out/${tag}_bench_ls$f.err; head -c 250 $out/${tag}_bench_ls$f.json; echo; done
timeout 200 python scripts/step_timeline.py --config c4 2> $out/${tag}_timeline_c4.err | c++filt > $out/${tag}_timeline_c4.txt; tail -2 $out/${tag}_timeline_c4.err; head -1 $out/${tag}_timeline_c4.txt
