#!/bin/bash
# usage: scripts_profile.sh <tag> [n]   — ncu launch list + full capture of the beam search kernel
TAG=${1:-r01}; N=${2:-1000000}
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --n $N --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:beam_search -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --n $N --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
