#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
FFSIM_B200_EXCHANGE=p2p timeout 200 compute-sanitizer --tool $tool --target-processes all --print-limit 10 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29719 scripts/dist_small_shapes.py \
  > gpurun_out/r2t_sanitize_p2p_${tool}.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_sanitize_p2p_${tool}.log
grep -E "^p2p|ERROR SUMMARY|RACECHECK SUMMARY|rc=" gpurun_out/r2t_sanitize_p2p_${tool}.log | tail -12
done
