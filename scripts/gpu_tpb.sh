#!/bin/bash
# developer experiment: fused kernel compiled for other CTA sizes (registers per thread = 64K / size)
OUT=gpurun_out/${1:-tpb}.jsonl; : > $OUT
run() { FFSIM_B200_LIB=$PWD/build/$1/libffsim_b200.so timeout 300 python scripts/quick_bench.py --norb $2 --nelec $3 $4 --only-rot --opts "$5" >> $OUT 2>&1; }
run t1024 16 5 5 "sub_window=5,threads=1024"
run t768 16 5 5 "sub_window=5,threads=768"
run t1024 18 7 7 "sub_window=5,threads=1024"
run t768 18 7 7 "sub_window=5,threads=768"
run t1024 16 5 5 "sub_window=4,threads=1024"
python - <<'PY'
import json,sys
for l in open("gpurun_out/s4o.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['opts'], d['plan'][:60], round(d['alpha_only_ms'],3), round(d['beta_only_ms'],3))
PY
