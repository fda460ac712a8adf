#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench_ours.log 2>&1
grep '^{' gpurun_out/bench_ours.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense-case > gpurun_out/ncu_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 344 -c 17 -o gpurun_out/prof_full -f python scratch/prof_run.py cfg4 21 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
