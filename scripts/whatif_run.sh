#!/bin/bash
# what-if knob timings of the fused kernel at the bench shapes (developer tool; needs build/dbg from scripts/build_dbg.sh)
TAG=${1:-whatif}
OUT=gpurun_out/${TAG}.txt; : > $OUT
for side in alpha beta; do
  echo "C3 $side knobs 0,8,24,40,56,4,2" >> $OUT
  FFB_SIDE=$side FFB_KNOBS=0,8,24,40,56,4,2 timeout 600 python scripts/whatif.py 18 7 7 >> $OUT 2>&1
done
echo "C2 alpha knobs 0,8,24,40,56,4,2" >> $OUT
FFB_KNOBS=0,8,24,40,56,4,2 timeout 300 python scripts/whatif.py 16 5 5 >> $OUT 2>&1
tail -c 1500 $OUT
