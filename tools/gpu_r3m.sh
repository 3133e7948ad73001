#!/bin/bash
# last verification of the round: GPU suite (incl. the batch prefetcher test), smoke, headline
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r3m_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r3m_smoke.log 2>&1; echo "smoke exit $?"; grep smoke: $OUT/r3m_smoke.log
timeout 300 python bench.py --no-extras --no-cpu-baseline > $OUT/r3m_bench.json 2> $OUT/r3m_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r3m_bench.json')); print('  ms', round(d['ms_per_step'],3), d['step_ms_min_med_max'], 'e2e', d['e2e']['ms_per_step_repetitions'], 'launches', d['launches_per_step'])"
