#!/bin/bash
# PDL on the training kernels (A/B), tree-merged statistics kernels: full GPU suite + headline benches
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r3g_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3g_pytest.log
for pdl in 1 0; do
EFFCONF_PDL_TRAIN=$pdl timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > $OUT/r3g_bench_pdl$pdl.json 2> $OUT/r3g_bench_pdl$pdl.err; echo "bench pdl_train=$pdl exit $?"; python -c "
import json; d=json.load(open('$OUT/r3g_bench_pdl$pdl.json')); print('  ms', round(d['ms_per_step'],3), d['step_ms_min_med_max'], 'launches', d['launches_per_step'], 'loss', d['loss_first_last'])"
done
timeout 300 python bench.py --no-extras --no-cpu-baseline --mode forward --steps 50 > $OUT/r3g_bench_fwd.json 2> $OUT/r3g_bench_fwd.err; echo "bench fwd exit $?"; python -c "
import json; d=json.load(open('$OUT/r3g_bench_fwd.json')); print('  fwd ms', round(d['ms_per_step'],3))"
