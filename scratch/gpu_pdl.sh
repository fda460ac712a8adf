mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded.py -m gpu -x -q -k "batch_edt or pipeline_parity or full_size or emulat or staged or sphere or cpp_host or golden or external or pool") > gpurun_out/pytest_subset.log 2>&1
tail -3 gpurun_out/pytest_subset.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dense-case > gpurun_out/bench_cur.log 2>&1
grep '^{' gpurun_out/bench_cur.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('stage_ms'), d.get('gpu_launches'))"
