#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py > gpurun_out/${TAG:-r2s}_sanitize_${tool}.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG:-r2s}_sanitize_${tool}.log
  tail -4 gpurun_out/${TAG:-r2s}_sanitize_${tool}.log
done
