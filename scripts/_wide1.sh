out=gpurun_out; tag=${1:-r04i}
mkdir -p $out
echo "== tests"; (timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_l1.py -x -q 2>&1 | tail -12) | tee $out/${tag}_tests.log
for c in c5-vgg16 c5-resnet; do
echo "== bench $c"; timeout 400 python bench.py --config $c --no-cpu-baseline > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err; head -c 300 $out/${tag}_bench_$c.json; echo; tail -3 $out/${tag}_bench_$c.err
done
