#!/bin/bash
# usage: gpu_r2_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
( timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q ) > gpurun_out/pytest_sharded_r2.log 2>&1
tail -5 gpurun_out/pytest_sharded_r2.log
fi
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_sharded_n$N.log 2>&1
grep '^{' gpurun_out/bench_sharded_n$N.log | cut -c1-1500
tail -5 gpurun_out/bench_sharded_n$N.log | cut -c1-400
