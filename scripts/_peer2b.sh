out=gpurun_out; tag=${1:-r04h}; n=${2:-2}
mkdir -p $out
echo "== dist tests"; (timeout 400 python -m pytest tests/test_gpu_dist_nccl.py -x -q 2>&1 | tail -8) | tee $out/${tag}_disttests.log
for t in peer nccl peer nccl; do
echo "== bench $n gpus transport=$t"
DEEPFLOWS_DP_TRANSPORT=$t timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu_$t.json 2> $out/${tag}_bench_${n}gpu_$t.err
head -c 230 $out/${tag}_bench_${n}gpu_$t.json; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_${n}gpu_$t.err | head -5
done
