mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_flat_tc.py tests/test_distance_family.py tests/test_full_size.py -q -m gpu -x 2>&1 | tail -5
run() { python scripts/flat_timeline.py --tag "$1" 2>> gpurun_out/flat_tl.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], round(d['pageable_ms'],3), round(d['pinned_ms'],3), [round(x,3) for x in d['pinned_ms_all']], d['candidates_last_level'], d['overflowed'])"; }
run default
for d in 12 16 24 32; do SDB_FLAT_SAMPLE_DIV=$d run div$d; done
SDB_FLAT_LEVELS=1 run levels
