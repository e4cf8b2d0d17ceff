#!/bin/bash
# One GPU-box call (1 GPU): parity tests, smoke, both bench arms, ncu launch lists and full ncu captures of the
# top kernels at both bench configurations (c3 = norb 18 (7,7), 16.2 GB; c2 = norb 16 (5,5), 305 MB).
set -u
TAG=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt; nproc >> gpurun_out/${TAG}_gpu.txt; free -g >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}c3_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}c3_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 3 > gpurun_out/${TAG}c2_bench.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --config c2 --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}c2_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
for cfg in c3 c2; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}${cfg}_launches.csv \
    python bench.py --config $cfg --no-c2 --steps 2 --warmup 3 > gpurun_out/${TAG}${cfg}_ncu_bench.log 2>&1
done
timeout 900 ncu --set full --clock-control none -k regex:'fused_pass_kernel|diag_kernel|transpose_kernel' -s 9 -c 9 -f -o gpurun_out/${TAG}c3_prof \
  python scripts/profile_target.py --norb 18 --nelec 7 7 --reps 2 > gpurun_out/${TAG}c3_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fused_pass_kernel|diag_kernel' -s 6 -c 3 -f -o gpurun_out/${TAG}c2_prof \
  python scripts/profile_target.py --reps 6 > gpurun_out/${TAG}c2_ncu_full.log 2>&1
# the reports are large (gpurun merges at most 64 MiB back): extract the metric tables here, keep only the small report
for cfg in c3 c2; do
  ncu -i gpurun_out/${TAG}${cfg}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}${cfg}_raw.csv 2> /dev/null
done
ncu -i gpurun_out/${TAG}c2_prof.ncu-rep --page source --csv > gpurun_out/${TAG}c2_source.csv 2> /dev/null
rm -f gpurun_out/${TAG}c3_prof.ncu-rep
[ $(stat -c %s gpurun_out/${TAG}c2_prof.ncu-rep 2>/dev/null || echo 0) -gt 30000000 ] && rm -f gpurun_out/${TAG}c2_prof.ncu-rep
du -sh gpurun_out; tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; cut -c1-300 gpurun_out/${TAG}c3_bench.json; tail -5 gpurun_out/${TAG}_bench.err
