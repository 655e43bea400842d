out=gpurun_out; tag=${1:-r04z}; n=${2:-8}
mkdir -p $out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
head -c 230 $out/${tag}_bench_${n}gpu.json; echo; grep -o '"dp_check.*' $out/${tag}_bench_${n}gpu.json | head -c 300; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_${n}gpu.err | head -5
