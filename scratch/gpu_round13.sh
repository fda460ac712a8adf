#!/bin/bash
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "pipeline_parity or golden" 2>&1 | tail -3
timeout 300 python scratch/e2e_probe.py prof 2>&1 | grep -E "prof gpu|waves|stats"
timeout 300 python scratch/wave_trace.py 2>&1 | grep -E "mean per|total ms"
