#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2.log
for v in 0 2 4 8; do GIE_STAGES_SCENE_ONLY=1 GIE_ZS_SYNC=$v python scratch/edt_stages.py cfg4 24 > gpurun_out/edt_stages_sync$v.log 2>&1; tail -1 gpurun_out/edt_stages_sync$v.log | cut -c1-400; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg4_r2.csv python scratch/prof_run.py cfg4 22 > gpurun_out/ncu_launches_r2.log 2>&1
tail -1 gpurun_out/ncu_launches_r2.log
