#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list, ncu --set full of every engine kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_ours.log 2>&1
tail -2 gpurun_out/bench_ours.log | cut -c1-1500
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref.log 2>&1
tail -2 gpurun_out/bench_ref.log | cut -c1-800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scratch/prof_run.py cfg4 4 > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 43 -c 13 -o gpurun_out/prof_full -f python scratch/prof_run.py cfg4 4 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
