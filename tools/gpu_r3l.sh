#!/bin/bash
# final default bench line with the pre-allocated e2e pipeline
OUT=gpurun_out; mkdir -p $OUT
EFFCONF_BENCH_VERBOSE=1 timeout 900 python bench.py > $OUT/r3l_bench_default.json 2> $OUT/r3l_bench_default.err; echo "bench default exit $?"
python -c "
import json; d=json.load(open('$OUT/r3l_bench_default.json'))
print('train ms', d['ms_per_step'], d['step_ms_min_med_max'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_repetitions'], 'fwd ms', d['forward']['ms_per_step'], d['forward']['e2e']['ms_per_step_repetitions'])
print('cpu', d['cpu_baseline']['value'], d['clocks'], 'launches', d['gpu_launches'])"
