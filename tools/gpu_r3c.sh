#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r3c_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r3c_smoke.log 2>&1; echo "smoke exit $?"; grep smoke: $OUT/r3c_smoke.log
EFFCONF_BENCH_VERBOSE=1 timeout 900 python bench.py > $OUT/r3c_bench_default.json 2> $OUT/r3c_bench_default.err; echo "bench default exit $?"; grep "bench rank" $OUT/r3c_bench_default.err | tail -9
python -c "
import json; d=json.load(open('$OUT/r3c_bench_default.json'))
print('train ms', d['ms_per_step'], 'fwd ms', d['forward']['ms_per_step'], 'launches', d['launches_per_step'])
print(json.dumps(d.get('kernel_time'))[:1200])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r3c_bench_reference.json 2> $OUT/r3c_bench_reference.err; echo "reference arm exit $?"; head -c 700 $OUT/r3c_bench_reference.json
