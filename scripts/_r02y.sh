out=gpurun_out; tag=${1:-r02y}
echo "== nccl dp tests"; timeout 300 python -m pytest tests/test_gpu_dist_nccl.py -q -x 2>&1 | grep -v "^$" | tail -30 > $out/${tag}_dp_tests.log; tail -3 $out/${tag}_dp_tests.log
for v in "DEEPFLOWS_BUCKETS_ON_SIDE=1"; do
echo "== bench 2 gpus $v"; env $v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extra --no-cpu-baseline > $out/${tag}_bench_2gpu_$v.json 2> $out/${tag}_bench_2gpu_$v.err; head -c 260 $out/${tag}_bench_2gpu_$v.json; echo; grep "diverged" $out/${tag}_bench_2gpu_$v.err | head -2
done
echo "== 1 gpu"; timeout 100 python bench.py --no-extra --no-cpu-baseline | head -c 260
