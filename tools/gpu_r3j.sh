#!/bin/bash
# final default bench line (e2e: best of two repetitions) + grid-size study of the LayerNorm backward
OUT=gpurun_out; mkdir -p $OUT
EFFCONF_BENCH_VERBOSE=1 timeout 900 python bench.py > $OUT/r3j_bench_default.json 2> $OUT/r3j_bench_default.err; echo "bench default exit $?"
python -c "
import json; d=json.load(open('$OUT/r3j_bench_default.json'))
print('train ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_repetitions'], 'fwd ms', d['forward']['ms_per_step'], d['forward']['e2e']['ms_per_step_repetitions'])
print('cpu', d['cpu_baseline']['value'], 'eager', json.dumps(d['torch_eager_b200'])[:400])"
for c in 148 296 444; do
EFFCONF_LNBWD_CTAS=$c timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > $OUT/r3j_bench_lnbwd$c.json 2> $OUT/r3j_bench_lnbwd$c.err; echo "bench lnbwd ctas=$c exit $?"; python -c "
import json; d=json.load(open('$OUT/r3j_bench_lnbwd$c.json')); print('  ms', round(d['ms_per_step'],3), d['step_ms_min_med_max'], 'e2e', d['e2e']['ms_per_step_repetitions']); [print('   ', o) for o in d['operators'] if 'layernorm_bwd' in o['op']]"
done
