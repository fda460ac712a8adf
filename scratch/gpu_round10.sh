#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 scratch/bench_sharded_edt.py 1024 1024 1016 > gpurun_out/sharded_n$N.log 2>&1; tail -1 gpurun_out/sharded_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 scratch/bench_sharded_edt.py 1024 1024 1016 0.00002 > gpurun_out/sharded_sparse_n$N.log 2>&1; tail -1 gpurun_out/sharded_sparse_n$N.log
