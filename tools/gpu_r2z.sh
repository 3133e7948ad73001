#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r2z_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r2z_pytest.log
for f in w1,res,dz,ln w1,res,dz,ln,lnf; do
EFFCONF_TRAIN_FUSE=$f timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > $OUT/r2z_bench_$f.json 2> $OUT/r2z_bench_$f.err; echo "fuse=$f exit $?"; python -c "
import json; d=json.load(open('$OUT/r2z_bench_$f.json')); print('  ms', round(d['ms_per_step'],3), d['step_ms_min_med_max'], 'launches', d['launches_per_step'])"
done
