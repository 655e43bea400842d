out=gpurun_out; tag=${1:-r02y}
echo "== nccl dp tests"; timeout 300 python -m pytest tests/test_gpu_dist_nccl.py -q -x 2>&1 | grep -v "^$" | tail -60 > $out/${tag}_dp_tests.log; grep -n "Error\|assert\|diverged\|rank" $out/${tag}_dp_tests.log | head -20
for v in "DEEPFLOWS_SIDE_LAG=0" "DFB_STAT_SLOTS=0" "DEEPFLOWS_FUSE=0" "A=1"; do
echo "== bench 2 gpus $v"; env $v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --no-extra --no-cpu-baseline > $out/${tag}_bench_2gpu_$v.json 2> $out/${tag}_bench_2gpu_$v.err; head -c 200 $out/${tag}_bench_2gpu_$v.json; grep "diverged" $out/${tag}_bench_2gpu_$v.err | head -2
done
