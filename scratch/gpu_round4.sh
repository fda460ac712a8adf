#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/wave_trace.py 2>&1 | tail -24
timeout 300 python scratch/e2e_probe.py prof 2>&1 | grep -E "prof gpu|waves|stats"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
