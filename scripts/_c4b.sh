out=gpurun_out; tag=${1:-r04t}
echo "== tests"; (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) | tee $out/${tag}_tests.log
echo "== bench"; timeout 200 python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 300 $out/${tag}_bench.json; echo; tail -2 $out/${tag}_bench.err
