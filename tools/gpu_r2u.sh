#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r2u_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r2u_pytest.log
for i in 1 2; do
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > $OUT/r2u_bench$i.json 2> $OUT/r2u_bench$i.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r2u_bench$i.json')); print('  ms', round(d['ms_per_step'],3), d['step_ms_min_med_max'], 'launches', d['launches_per_step'])"
done
