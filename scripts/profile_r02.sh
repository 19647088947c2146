#!/bin/bash
# Round-2 evidence run on one B200 (profiles/README.md): the default bench line, per-config bench
# lines, the ncu launch list of the default bench command, ncu --set full captures of the
# beam-search kernel for C2 (DRAM traffic -> profiles/k1_traffic.json), the hamming search and the
# PQ search, the flat search timing.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1_r02.json 2> gpurun_out/bench_n1_r02.err
for w in c3 c4 c4cos c5b; do
  python bench.py --workload $w --extra none --steps 10 > gpurun_out/bench_${w}_r02.json 2> gpurun_out/bench_${w}_r02.err
done
python bench.py --workload c4 --points 10000000 --extra none --steps 5 --no-probe > gpurun_out/bench_c4_10m_r02.json 2> gpurun_out/bench_c4_10m_r02.err
python bench.py --workload c5a --steps 3 > gpurun_out/bench_c5a_r02.json 2> gpurun_out/bench_c5a_r02.err
SDB_FLAT_SKIP_EXACT=1 python scripts/bench_configs.py flat > gpurun_out/flat_r02.json 2> /dev/null
python scripts/flat_timeline.py --reps 7 --timeline --tag two-pass > gpurun_out/r02_flat_timeline.json 2> /dev/null
SDB_FLAT_LEVELS=1 SDB_FLAT_NO_CENTER=1 python scripts/flat_timeline.py --reps 7 --tag "level scheme, no centring" > gpurun_out/r02_flat_levels.json 2> /dev/null
ncu --set full --clock-control none --import-source on -k regex:'tc5_filter|kth_thresh|rescore_warp' --launch-skip 8 -c 4 -f -o gpurun_out/prof_flat_r02 \
    python scripts/flat_timeline.py --reps 1 > gpurun_out/ncu_flat.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r02.log 2>&1
SDB_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:beam_search -c 1 -f \
    -o gpurun_out/prof_k1_r02 python bench.py --graph gpu --extra none --steps 1 --no-probe --no-cpu-baseline > gpurun_out/ncu_k1_r02.log 2>&1
SDB_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:beam_search -c 1 -f \
    -o gpurun_out/prof_k3fly_r02 python bench.py --workload c4 --points 300000 --extra none --steps 1 --no-probe > gpurun_out/ncu_k3fly_r02.log 2>&1
SDB_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:beam_search -c 1 -f \
    -o gpurun_out/prof_k2_r02 python bench.py --workload c5b --points 2000000 --extra none --steps 1 --no-probe > gpurun_out/ncu_k2_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
