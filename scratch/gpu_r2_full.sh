#!/bin/bash
# round 2: the gpurun command line behind profiles/r02_* (tests, stage probe, both bench arms)
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -1 gpurun_out/smoke_r2.log
for w in 8 16 32; do GIE_ZS_WPC=$w python scratch/edt_stages.py cfg4 24 > gpurun_out/edt_stages_wpc$w.log 2>&1; tail -1 gpurun_out/edt_stages_wpc$w.log; done
( time timeout 900 python bench.py ) > gpurun_out/bench_ours_r2.log 2>&1
grep '^{' gpurun_out/bench_ours_r2.log | cut -c1-600
( time timeout 900 python bench.py --impl reference --steps 50 --warmup 5 ) > gpurun_out/bench_ref_r2.log 2>&1
grep '^{' gpurun_out/bench_ref_r2.log | cut -c1-400
