out=gpurun_out; tag=${1:-r04d}; n=${2:-2}
for t in peer nccl; do
DFB_TIMELINE_PREFIX=$out/${tag}_timeline_${n}gpu_$t DEEPFLOWS_DP_TRANSPORT=$t timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 scripts/step_timeline.py > $out/${tag}_timeline_${n}gpu_$t.log 2>&1
tail -3 $out/${tag}_timeline_${n}gpu_$t.log; ls $out/${tag}_timeline_${n}gpu_${t}_rank*
done
