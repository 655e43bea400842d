out=gpurun_out; tag=${1:-r04o}; n=${2:-2}
mkdir -p $out
echo "== dist tests"; (timeout 400 python -m pytest tests/test_gpu_dist_nccl.py -x -q 2>&1 | tail -8) | tee $out/${tag}_disttests.log
for ch in default 1 2 4; do
echo "== bench $n gpus hybrid NCCL_MAX_NCHANNELS=$ch"
if [ $ch = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > $out/${tag}_bench_${n}gpu_ch$ch.json 2> $out/${tag}_bench_${n}gpu_ch$ch.err
head -c 230 $out/${tag}_bench_${n}gpu_ch$ch.json; echo; grep -i "diverged\|Error\|DeepFlows.dist" $out/${tag}_bench_${n}gpu_ch$ch.err | head -5
done
