#!/bin/bash
# developer A/B: the same timing script against two builds of the library (build/old = an earlier commit)
OUT=gpurun_out/${1:-ab}.jsonl; : > $OUT
for shape in "12 6 6" "14 7 7" "10 5 5" "16 8 8"; do
  set -- $shape
  for lib in build/old/libffsim_b200.so ffsim_b200/lib/libffsim_b200.so; do
    echo "$lib $shape" >> $OUT
    FFSIM_B200_LIB=$PWD/$lib timeout 300 python scripts/quick_bench.py --norb $1 --nelec $2 $3 --only-rot 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['plan'][:70], round(d['orbital_rotation_ms'],4), round(d['alpha_only_ms'],4), round(d['beta_only_ms'],4))
    except Exception: print(l[:200])
" >> $OUT
  done
done
cat $OUT
