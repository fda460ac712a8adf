#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 344 -c 17 -o gpurun_out/prof_full -f python scratch/prof_run.py cfg4 21 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv python scratch/prof_run.py cfg4 22 > gpurun_out/ncu_launches.log 2>&1
