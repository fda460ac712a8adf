#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref.log 2>&1
tail -2 gpurun_out/bench_ref.log | cut -c1-600
( time timeout 900 python bench.py ) > gpurun_out/bench_ours.log 2>&1
tail -4 gpurun_out/bench_ours.log | cut -c1-3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv python scratch/prof_run.py cfg4 22 > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 344 -c 17 -o gpurun_out/prof_full -f python scratch/prof_run.py cfg4 21 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
