out=gpurun_out; tag=${1:-r04r}; n=${2:-2}
mkdir -p $out
for mb in 4 8 16 2; do
echo "== bench $n gpus hybrid bucket_mb=$mb"
DEEPFLOWS_DP_BUCKET_MB=$mb timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu_mb$mb.json 2> $out/${tag}_bench_${n}gpu_mb$mb.err
head -c 230 $out/${tag}_bench_${n}gpu_mb$mb.json; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_${n}gpu_mb$mb.err | head -5
done
