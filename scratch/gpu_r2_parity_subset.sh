# parity subset + stage times for a build under test (about 2.5 GPU-minutes)
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded.py -m gpu -x -q -k "batch_edt or pipeline_parity or full_size_properties or emulat or staged or pntcld or ray or sphere") > gpurun_out/pytest_subset.log 2>&1
tail -3 gpurun_out/pytest_subset.log
python scratch/edt_stages.py cfg4 24 > gpurun_out/edt_stages_cur.log 2>&1
tail -1 gpurun_out/edt_stages_cur.log
