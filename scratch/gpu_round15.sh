#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | head -2
( time timeout 900 python bench.py ) > gpurun_out/bench_ours.log 2>&1
grep '^{' gpurun_out/bench_ours.log | cut -c1-300
