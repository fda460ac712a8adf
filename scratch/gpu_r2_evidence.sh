#!/bin/bash
# round 2 evidence: ncu full capture of one frame, launch list of bench.py, both bench arms on cfg4, cfg1-cfg3 lines
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 401 -c 20 -o gpurun_out/prof_full_r2 -f python scratch/prof_run.py cfg4 21 > gpurun_out/ncu_full_r2.log 2>&1
tail -1 gpurun_out/ncu_full_r2.log
ncu -i gpurun_out/prof_full_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_r2_raw.csv 2>/dev/null
for k in k_edt_xsweep k_edt_zsweep k_pc_apply k_waves; do ncu -i gpurun_out/prof_full_r2.ncu-rep --page details --kernel-name regex:$k 2>/dev/null | grep -E "Duration|Throughput|Issue|Eligible|Occupancy|Active Warps|Registers|Shared Memory|Hit Rate|Cycles Per|Section" | head -60 > gpurun_out/ncu_details_$k.txt; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense-case > gpurun_out/ncu_launches_bench_r2.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 50 --warmup 5 ) > gpurun_out/bench_ref_r2.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench_ours_r2.log 2>&1
grep '^{' gpurun_out/bench_ours_r2.log | cut -c1-300
for c in cfg1 cfg2 cfg3; do
( timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-dense-case ) > gpurun_out/bench_ours_$c.log 2>&1
( timeout 600 python bench.py --config $c --impl reference --steps 30 --warmup 5 ) > gpurun_out/bench_ref_$c.log 2>&1
grep '^{' gpurun_out/bench_ours_$c.log | cut -c1-200; grep '^{' gpurun_out/bench_ref_$c.log | cut -c1-300
done
