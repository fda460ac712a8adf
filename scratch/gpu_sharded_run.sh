#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_sharded.py -x -q ) > gpurun_out/pytest_sharded.log 2>&1
tail -3 gpurun_out/pytest_sharded.log
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 scratch/bench_sharded_edt.py 1024 1024 1016 > gpurun_out/sharded_n$N.log 2>&1; tail -1 gpurun_out/sharded_n$N.log
