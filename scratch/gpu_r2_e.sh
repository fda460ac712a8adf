#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2e.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2e.log
GIE_STAGES_SCENE_ONLY=1 python scratch/edt_stages.py cfg4 24 > gpurun_out/edt_stages_e.log 2>&1; tail -1 gpurun_out/edt_stages_e.log
