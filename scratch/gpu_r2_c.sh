#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_r2.log 2>&1
tail -15 gpurun_out/pytest_gpu_r2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_edt_(x|z)sweep" -s 20 -c 2 -o gpurun_out/prof_scene_r2 -f python scratch/prof_run.py cfg4 12 > gpurun_out/ncu_scene.log 2>&1; tail -2 gpurun_out/ncu_scene.log
