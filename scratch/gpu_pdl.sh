mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded.py -m gpu -x -q -k "batch_edt or pipeline_parity or full_size_properties or emulat or staged or pntcld or ray or sphere") > gpurun_out/pytest_subset.log 2>&1
tail -3 gpurun_out/pytest_subset.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dense-case > gpurun_out/bench_pdl.log 2>&1
grep '^{' gpurun_out/bench_pdl.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('PDL   ', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('stage_ms'))"
GIE_NO_PDL=1 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dense-case > gpurun_out/bench_nopdl.log 2>&1
grep '^{' gpurun_out/bench_nopdl.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('no PDL', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('stage_ms'))"
