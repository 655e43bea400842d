run() { name=$1; shift; echo "=== $name"; env "$@" timeout 32 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT scripts/dp_smoke.py > gpurun_out/dp_$name.log 2>&1; grep -E "rank 0 .*done|Timeout|NCCL version" gpurun_out/dp_$name.log | head -3 | cut -c1-100; PORT=$((PORT+3)); }
PORT=29551
for i in 1 2 3 4 5; do run bundled_$i X=1; done
echo "=== bench graph"; timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1300
