#!/bin/bash
# kernel iteration check: parity subset + per-side timings at C2 and C3 (+ optional extra option sets)
TAG=${1:-r2m}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
bash scripts/sweep.sh ${TAG}_sweep "16 5 5|" "18 7 7|" "12 6 6|" "16 8 8|" "$@" > /dev/null 2>&1
{ nvidia-smi topo -m; lscpu | head -25; cat /sys/devices/system/node/online; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done; } > gpurun_out/${TAG}_topo.txt 2>&1
tail -2 gpurun_out/${TAG}_pytest.log; cut -c1-400 gpurun_out/${TAG}_sweep.jsonl
