#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_dropin.py -m gpu -q > $OUT/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/r2j_pytest.log
for f in w1,res,dz w1,res,dz,ln; do
  EFFCONF_TRAIN_FUSE=$f timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > $OUT/r2j_bench_$f.json 2> $OUT/r2j_bench_$f.err; echo "fuse=$f exit $?"; python -c "
import json; d=json.load(open('$OUT/r2j_bench_$f.json')); print('  ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step']); [print('   ', o) for o in d['operators'][:8]]"
done
