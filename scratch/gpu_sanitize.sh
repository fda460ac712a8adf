# compute-sanitizer on a small end-to-end run of every kernel (final build)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  ( time timeout 900 compute-sanitizer --tool $tool python scratch/sanitize_run.py ) > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log | head -2
  grep -c "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log
  grep "SUMMARY" gpurun_out/sanitizer_$tool.log | tail -1
done
