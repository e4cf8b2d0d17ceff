#!/bin/bash
# multi-GPU check: sharded tests at N ranks (both exchanges) + the contract bench at N ranks (default exchange);
# C4 (254 GB, 8 ranks): torchrun --nproc-per-node 8 scripts/bench_sharded.py --norb 20 --nelec 8 8 --n-reps 3 --steps 2
N=${1:-2}; TAG=${2:-r3m}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
FFSIM_B200_WATCHDOG=120 timeout 500 python -m pytest tests/test_gpu_distributed.py -x -q -k "world_size_n and $N" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "rc=$?" >> gpurun_out/${TAG}_bench_n${N}.err
tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_bench_n${N}.err; cut -c1-400 gpurun_out/${TAG}_bench_n${N}.json
