#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -m gpu -q -x > $OUT/r2i_pytest.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/r2i_pytest.log
for f in none w1 res dz res,dz w1,res,dz; do
  EFFCONF_TRAIN_FUSE=$f timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > $OUT/r2i_bench_$f.json 2> $OUT/r2i_bench_$f.err; echo "fuse=$f exit $?"; python -c "
import json; d=json.load(open('$OUT/r2i_bench_$f.json')); print('  ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step'])"
done
python -c "
import json; d=json.load(open('$OUT/r2i_bench_none.json')); [print(o) for o in d['operators'][:22]]"
