#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --steps 50 --warmup 10 ) > gpurun_out/bench_ours.log 2>&1
tail -2 gpurun_out/bench_ours.log | cut -c1-2500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scratch/prof_run.py cfg4 22 > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 264 -c 13 -o gpurun_out/prof_full -f python scratch/prof_run.py cfg4 21 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
