#!/bin/bash
# round 2, first GPU call: pipe-concurrency microbenchmark, option sweep of the round-1 kernel, timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
./build/fp64_pipes > gpurun_out/r2a_fp64_pipes.jsonl 2>&1
bash scripts/sweep.sh r2a_sweep \
  "16 5 5|" \
  "16 5 5|min_cols=1,smem_bytes=72000,threads=256" \
  "16 5 5|min_cols=1,smem_bytes=72000,threads=192" \
  "16 5 5|min_cols=1,smem_bytes=72000,threads=128" \
  "16 5 5|threads=384" \
  "16 5 5|min_cols=2,smem_bytes=100000,threads=256" \
  "18 7 7|" > /dev/null 2>&1
timeout 300 python scripts/timeline.py 16 5 5 --chart > gpurun_out/r2a_timeline_16.txt 2>&1
timeout 300 python scripts/phase_timing.py 16 5 5 > gpurun_out/r2a_phase_16.txt 2>&1
timeout 300 python scripts/whatif.py 16 5 5 > gpurun_out/r2a_whatif_16.txt 2>&1
echo done
