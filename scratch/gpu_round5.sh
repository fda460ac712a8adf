#!/bin/bash
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
for v in 16 8; do echo "== cluster $v"; GIE_WAVE_CLUSTER=$v timeout 300 python scratch/wave_trace.py 2>&1 | grep -E "^(8|32|72|112|152|192|232) |mean per"; done
timeout 300 python scratch/e2e_probe.py prof 2>&1 | grep -E "prof gpu|waves|stats"
