out=gpurun_out; tag=${1:-r04a}; n=${2:-2}
mkdir -p $out
echo "== dist tests"; (timeout 400 python -m pytest tests/test_gpu_dist_nccl.py tests/test_gpu_fused.py -x -q 2>&1 | tail -25) | tee $out/${tag}_disttests.log
for t in peer nccl; do
echo "== bench $n gpus transport=$t"
DEEPFLOWS_DP_TRANSPORT=$t timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu_$t.json 2> $out/${tag}_bench_${n}gpu_$t.err
head -c 260 $out/${tag}_bench_${n}gpu_$t.json; echo; grep -o '"dp_check.*' $out/${tag}_bench_${n}gpu_$t.json | head -c 400; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_${n}gpu_$t.err | head -5
done
echo "== 1 gpu"; timeout 100 python bench.py --no-extra --no-cpu-baseline > $out/${tag}_bench_1gpu.json; head -c 260 $out/${tag}_bench_1gpu.json; echo
DFB_POOL2_FAST=0 timeout 100 python bench.py --no-extra --no-cpu-baseline > $out/${tag}_bench_1gpu_nopool2.json; head -c 260 $out/${tag}_bench_1gpu_nopool2.json
echo "== gemm bench"; timeout 60 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/${tag}_gemm_bench.txt
