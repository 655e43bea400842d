out=gpurun_out; tag=${1:-r04l}; n=${2:-2}
mkdir -p $out
for parts in tail mid all; do
echo "== bench $n gpus peer parts=$parts"
DEEPFLOWS_DP_PEER_PARTS=$parts timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu_$parts.json 2> $out/${tag}_bench_${n}gpu_$parts.err
head -c 230 $out/${tag}_bench_${n}gpu_$parts.json; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_${n}gpu_$parts.err | head -5
done
echo "== nccl"; DEEPFLOWS_DP_TRANSPORT=nccl timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline 2>/dev/null | head -c 230; echo
for parts in tail all; do
DFB_TIMELINE_PREFIX=$out/${tag}_timeline_${n}gpu_$parts DEEPFLOWS_DP_PEER_PARTS=$parts timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 scripts/step_timeline.py > $out/${tag}_timeline_${n}gpu_$parts.log 2>&1
done
ls $out/${tag}_timeline*
