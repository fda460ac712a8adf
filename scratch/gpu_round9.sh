#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 30 --warmup 10 > gpurun_out/bench_n$N.log 2>&1
tail -1 gpurun_out/bench_n$N.log | cut -c1-1200
tail -1 gpurun_out/bench_n$N.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('sharded_batch_edt'))"
