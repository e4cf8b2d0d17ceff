#!/bin/bash
# Multi-GPU round (run under gpurun --gpus N): NCCL sharded tests, sharded LUCJ bench, replica bench.
N=${1:-2}; TAG=${2:-r1_multi}; NORB=${3:-16}; NA=${4:-8}; NB=${5:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${TAG}_gpus.txt
nvidia-smi topo -m >> gpurun_out/${TAG}_gpus.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29720 \
  scripts/bench_sharded.py --norb $NORB --nelec $NA $NB --n-reps 3 --steps 3 > gpurun_out/${TAG}_sharded.json 2> gpurun_out/${TAG}_sharded.err
cat gpurun_out/${TAG}_sharded.json; tail -5 gpurun_out/${TAG}_sharded.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-600; tail -3 gpurun_out/${TAG}_bench.err
