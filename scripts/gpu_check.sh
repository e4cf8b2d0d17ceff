#!/bin/bash
# round 2 GPU check: host facts, GPU test-suite, the contract bench at N=1
mkdir -p gpurun_out
TAG=${1:-r2i}
{ free -g; nproc; cat /sys/fs/cgroup/memory.max 2>/dev/null; nvidia-smi -L; } > gpurun_out/${TAG}_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_pytest.log
