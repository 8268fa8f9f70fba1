# compute-sanitizer passes over tools/sanitize_small.py (release library)
mkdir -p gpurun_out
for tool in racecheck memcheck; do
  timeout 110 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_small.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -v "^\[W\|Warning" gpurun_out/r2_sanitizer_$tool.log | tail -6
done
