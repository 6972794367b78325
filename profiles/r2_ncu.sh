#!/bin/bash
# Round-2 ncu captures (one B200):
#   gpurun --timeout 1500 -- 'bash profiles/r2_ncu.sh [rows] [kernel regex]'
# 1. launch list of two steps of the default bench workload (cold-cache, serialised: use the shares)
# 2. ncu --set full (+ source) of one launch of each dominant kernel
rows=${1:-100000000}
pat=${2:-'k_ingest|k_merge_filter|k_encode_v4'}
mkdir -p gpurun_out
B="python bench.py --rows $rows --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$pat" -c 3 -f -o gpurun_out/r2_prof $B > gpurun_out/r2_prof.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/r2_prof.log
ls -la gpurun_out
