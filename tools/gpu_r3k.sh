#!/bin/bash
# e2e diagnostic: per-step host trace with K = 40 / 30 steps, PDL on the training kernels on / off
OUT=gpurun_out; mkdir -p $OUT
for cfg in "40 1" "30 1" "40 0"; do
set -- $cfg
EFFCONF_E2E_TRACE=1 EFFCONF_PDL_TRAIN=$2 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps $1 > $OUT/r3k_bench_k$1_pdl$2.json 2> $OUT/r3k_bench_k$1_pdl$2.err; echo "bench steps=$1 pdl=$2 exit $?"; python -c "
import json; d=json.load(open('$OUT/r3k_bench_k$1_pdl$2.json')); print('  ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['ms_per_step_repetitions'], d['clocks'])"
grep "e2e trace" $OUT/r3k_bench_k$1_pdl$2.err | cut -c1-1500
done
