#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 900 python -m pytest tests/test_sharded.py -x -q ) > gpurun_out/pytest_sharded.log 2>&1
tail -8 gpurun_out/pytest_sharded.log
timeout 600 python scratch/bench_sharded_edt.py 1024 1024 1016 > gpurun_out/sharded_n1.log 2>&1; tail -1 gpurun_out/sharded_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scratch/bench_sharded_edt.py 1024 1024 1016 > gpurun_out/sharded_n2.log 2>&1; tail -1 gpurun_out/sharded_n2.log
