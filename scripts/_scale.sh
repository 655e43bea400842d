out=gpurun_out; tag=${1:-r03c}
for n in 8 4 2; do
echo "== bench $n gpus"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err; head -c 230 $out/${tag}_bench_${n}gpu.json; echo; grep "diverged\|Error" $out/${tag}_bench_${n}gpu.err | head -3
done
echo "== 1 gpu"; timeout 100 python bench.py --no-extra --no-cpu-baseline > $out/${tag}_bench_1gpu.json; head -c 230 $out/${tag}_bench_1gpu.json
