#!/bin/bash
# compute-sanitizer passes over every kernel of the library (run under gpurun); logs -> gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/san_workload.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
