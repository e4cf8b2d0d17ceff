#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2k_dbg.log
for mode in nccl p2p; do
FFSIM_B200_EXCHANGE=$mode timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 scripts/r2_dbg_small.py >> gpurun_out/r2k_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_dbg.log
done
grep -E "^(nccl|p2p|rc)" gpurun_out/r2k_dbg.log
