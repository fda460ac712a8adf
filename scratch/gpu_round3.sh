#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python scratch/e2e_probe.py prof > gpurun_out/e2e_probe.log 2>&1
cat gpurun_out/e2e_probe.log
