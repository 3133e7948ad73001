#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r2r_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r2r_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > $OUT/r2r_bench.json 2> $OUT/r2r_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r2r_bench.json')); print('  ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step']); [print('   ', o) for o in d['operators'][:5]]"
timeout 300 python bench.py --no-extras --no-cpu-baseline --mode forward --steps 50 > $OUT/r2r_bench_fwd.json 2> $OUT/r2r_bench_fwd.err; echo "bench fwd exit $?"; python -c "
import json; d=json.load(open('$OUT/r2r_bench_fwd.json')); print('  fwd ms', round(d['ms_per_step'],3))"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv \
    --log-file $OUT/r2r_launches_warm_train_bf16x2.csv python tools/ncu_train_target.py --precision bf16x2 > $OUT/r2r_ncu_warm.log 2>&1
echo "ncu warm train exit $?"
