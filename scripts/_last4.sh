out=gpurun_out; tag=${1:-r04end}
(timeout 100 python -m pytest tests/test_gpu_dist_nccl.py -x -q 2>&1 | tail -3) | tee $out/${tag}_disttests.log
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extra --no-cpu-baseline > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err
head -c 230 $out/${tag}_bench_2gpu.json; echo; grep -o '"dp_check.*' $out/${tag}_bench_2gpu.json | head -c 300; echo; grep -i "diverged\|Error" $out/${tag}_bench_2gpu.err | head -3
