#!/bin/bash
# compute-sanitizer over tools/sanitizer_run.py: memcheck with the int8 sessions, racecheck / synccheck / initcheck (EQVIO_SANITIZE_INT8 from the caller)
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool python tools/sanitizer_run.py > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$tool.log | tail -1)"
  grep -E "^N |narrow|gemm|chain" gpurun_out/san_$tool.log | tail -12
done
