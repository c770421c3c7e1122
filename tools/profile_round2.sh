#!/bin/sh
# Round-2 profile collection (run under gpurun, ONE GPU): every capture only after the same command has
# exited 0 without ncu.  Writes into gpurun_out/; the summaries are copied to profiles/ by hand.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
# 1. launch list of the bench command (cold-cache, serialised per-launch times)
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --legs '' > $O/r2_prof_bench.json 2> $O/r2_prof_bench.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --legs '' > $O/r2_prof_ncu1.log 2>&1
# 2. DRAM traffic of the two scan launches AT BENCH SIZE (148 x 2^20 points)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:scan_fast -c 2 --csv --log-file $O/r2_traffic.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --legs '' > $O/r2_prof_ncu2.log 2>&1
# 3. the DFMA peak microbenchmark under ncu: is the FP64 pipe saturated?
timeout 100 python tools/peak_only.py > $O/r2_peak.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none -k regex:dfma -c 2 -f -o $O/r2_dfma python tools/peak_only.py > $O/r2_prof_ncu3.log 2>&1
# 4. full set on the fused log-likelihood scan (148 x 16384 points)
timeout 100 python tools/run_timing.py > $O/r2_run_timing.log 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_fast -c 1 -f -o $O/r2_scan_base python tools/run_timing.py > $O/r2_prof_ncu4.log 2>&1
ls -la $O/*.ncu-rep $O/r2_launches.csv $O/r2_traffic.csv
