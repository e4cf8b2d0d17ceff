#!/bin/bash
# Option sweep of the fused Givens kernel (developer tool): one JSON line per configuration.
#   scripts/sweep.sh <tag> "<norb> <na> <nb>|<opts>" ...      (a leading "T" runs the GPU parity tests first)
TAG=$1; shift
OUT=gpurun_out/${TAG}.jsonl
mkdir -p gpurun_out; : > $OUT
for spec in "$@"; do
  if [ "$spec" = "T" ]; then
    timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $OUT
  elif [ "${spec:0:1}" = "W" ]; then   # what-if knobs: "W 16 5 5|opts"
    s=${spec:2}; shape=${s%%|*}; opts=${s#*|}
    [ -f build/dbg/libffsim_b200.so ] || bash scripts/build_dbg.sh > /dev/null 2>&1
    echo "whatif $shape $opts" >> $OUT; timeout 300 python scripts/whatif.py $shape "$opts" >> $OUT 2>&1
  else
    shape=${spec%%|*}; opts=${spec#*|}
    set -- $shape
    timeout 600 python scripts/quick_bench.py --norb $1 --nelec $2 $3 --only-rot --opts "$opts" >> $OUT 2>&1
  fi
done
cat $OUT
