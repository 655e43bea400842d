out=gpurun_out; tag=${1:-r04n}; n=${2:-8}
mkdir -p $out
for cfg in nccl tail all; do
echo "== bench $n gpus $cfg"
if [ $cfg = nccl ]; then export DEEPFLOWS_DP_TRANSPORT=nccl; else export DEEPFLOWS_DP_TRANSPORT=peer DEEPFLOWS_DP_PEER_PARTS=$cfg; fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu_$cfg.json 2> $out/${tag}_bench_${n}gpu_$cfg.err
head -c 230 $out/${tag}_bench_${n}gpu_$cfg.json; echo; grep -o '"dp_check.*' $out/${tag}_bench_${n}gpu_$cfg.json | head -c 300; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_${n}gpu_$cfg.err | head -5
done
