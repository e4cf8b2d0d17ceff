#!/bin/bash
# session-4 call A: parity tests, bench (both arms), phase/what-if breakdown of the fused kernel
set -u
TAG=${1:-s4a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python scripts/phase_timing.py 16 5 5 > gpurun_out/${TAG}_phase_16.txt 2>&1
timeout 200 python scripts/phase_timing.py 18 7 7 > gpurun_out/${TAG}_phase_18.txt 2>&1
timeout 200 python scripts/whatif.py 16 5 5 > gpurun_out/${TAG}_whatif_16.txt 2>&1
timeout 200 python scripts/whatif.py 18 7 7 > gpurun_out/${TAG}_whatif_18.txt 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_phase_*.txt gpurun_out/${TAG}_whatif_*.txt
