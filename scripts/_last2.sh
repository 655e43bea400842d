out=gpurun_out; tag=${1:-r04x}
echo "== dist tests"; (timeout 300 python -m pytest tests/test_gpu_dist_nccl.py -x -q 2>&1 | tail -4) | tee $out/${tag}_disttests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extra --no-cpu-baseline > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err
head -c 230 $out/${tag}_bench_2gpu.json; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_2gpu.err | head -5
timeout 200 python scripts/step_timeline.py --config c4 2> $out/${tag}_timeline_c4.err | c++filt > $out/${tag}_timeline_c4.txt; tail -2 $out/${tag}_timeline_c4.err; head -1 $out/${tag}_timeline_c4.txt
