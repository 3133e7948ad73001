#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r2t_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r2t_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > $OUT/r2t_bench.json 2> $OUT/r2t_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r2t_bench.json')); print('  ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step']); [print('   ', o) for o in d['operators'][:30]]"
