#!/bin/bash
# usage: scripts/profile.sh <tag> [n]
#   1. ncu launch list of `python bench.py` (gpu__time_duration only, no clock control)
#   2. ncu --set full capture of one beam-search launch of the same command (DRAM traffic -> k1_traffic)
#   3. ncu --set full capture of the tcgen05 flat candidate pass (scripts/bench_configs.py flat)
TAG=${1:-r01}; N=${2:-1000000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --n $N --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:beam_search -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --n $N --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
SDB_FLAT_SKIP_EXACT=1 ncu --set full --clock-control none --import-source on -k regex:tc5_filter -s 2 -c 1 -f -o gpurun_out/prof_tc5_$TAG \
    python scripts/bench_configs.py flat > gpurun_out/ncu_tc5_$TAG.log 2>&1
