#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded.py -m gpu -x -q -k "batch_edt or pipeline_parity or full_size_properties or emulation or staged" ) > gpurun_out/pytest_gpu_r2d.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2d.log
python scratch/edt_stages.py cfg4 24 > gpurun_out/edt_stages_d.log 2>&1; tail -1 gpurun_out/edt_stages_d.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_edt_(x|z)sweep" -s 20 -c 2 -o gpurun_out/prof_scene_r2d -f python scratch/prof_run.py cfg4 12 > gpurun_out/ncu_scene_d.log 2>&1; tail -1 gpurun_out/ncu_scene_d.log
