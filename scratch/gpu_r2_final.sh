#!/bin/bash
# round 2, final build: the one gpurun command behind profiles/r02_*_final.* (tests, smoke, ncu captures, both bench arms, cfg1-cfg3)
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 339 -c 20 -o gpurun_out/prof_full_final -f python scratch/prof_run.py cfg4 21 > gpurun_out/ncu_full_final.log 2>&1
tail -1 gpurun_out/ncu_full_final.log
ncu -i gpurun_out/prof_full_final.ncu-rep --page raw --csv > gpurun_out/ncu_full_final_raw.csv 2>/dev/null
for k in k_edt_xsweep k_edt_zsweep k_pc_apply k_pc_walk k_frontiers k_waves; do ncu -i gpurun_out/prof_full_final.ncu-rep --page details --kernel-name regex:$k 2>/dev/null | grep -E "Duration|Throughput|Issue|Eligible|Occupancy|Active Warps|Registers|Shared Memory|Hit Rate|Cycles Per|Section" | head -60 > gpurun_out/ncu_details_final_$k.txt; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense-case > gpurun_out/ncu_launches_bench_final.log 2>&1
python scratch/edt_stages.py cfg4 24 > gpurun_out/edt_stages_final.log 2>&1; tail -1 gpurun_out/edt_stages_final.log
( time timeout 900 python bench.py --impl reference --steps 50 --warmup 5 ) > gpurun_out/bench_ref_final.log 2>&1
grep '^{' gpurun_out/bench_ref_final.log | cut -c1-300
( time timeout 900 python bench.py ) > gpurun_out/bench_ours_final.log 2>&1
grep '^{' gpurun_out/bench_ours_final.log | cut -c1-400
for c in cfg1 cfg2 cfg3; do
( timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-dense-case ) > gpurun_out/bench_ours_final_$c.log 2>&1
( timeout 600 python bench.py --config $c --impl reference --steps 30 --warmup 5 ) > gpurun_out/bench_ref_final_$c.log 2>&1
grep '^{' gpurun_out/bench_ours_final_$c.log | cut -c1-200; grep '^{' gpurun_out/bench_ref_final_$c.log | cut -c1-200
done
