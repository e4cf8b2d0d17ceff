#!/bin/bash
# round 2, 8-GPU call: 8-rank sharded tests, the contract bench (both exchanges), C4 (LUCJ n_reps=3, norb=20 (8,8), 254 GB)
TAG=${1:-r2q}; N=8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
FFSIM_B200_WATCHDOG=100 timeout 300 python -m pytest tests/test_gpu_distributed.py -x -q -k "world_size_n and 8" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
for mode in nccl p2p; do
  FFSIM_B200_EXCHANGE=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${mode}.json 2> gpurun_out/${TAG}_bench_${mode}.err; echo "rc=$?" >> gpurun_out/${TAG}_bench_${mode}.err
done
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  scripts/bench_sharded.py --norb 20 --nelec 8 8 --n-reps 3 --steps 2 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err; echo "rc=$?" >> gpurun_out/${TAG}_c4.err
tail -2 gpurun_out/${TAG}_pytest.log; tail -c 600 gpurun_out/${TAG}_c4.json
