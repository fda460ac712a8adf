mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu_full.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dense-case > gpurun_out/bench_cur.log 2>&1
grep '^{' gpurun_out/bench_cur.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('stage_ms'), d.get('gpu_launches'))"
