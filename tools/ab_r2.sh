mkdir -p gpurun_out
for t in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $t python tools/sanitize_smoke.py > gpurun_out/r02_sanitizer_$t.log 2>&1
  tail -4 gpurun_out/r02_sanitizer_$t.log
done
