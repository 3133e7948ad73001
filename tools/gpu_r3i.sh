#!/bin/bash
# end-of-round evidence: GPU suite, smoke, default bench line, reference arm, warm ncu launch lists (train + forward)
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r3i_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r3i_smoke.log 2>&1; echo "smoke exit $?"; grep smoke: $OUT/r3i_smoke.log
EFFCONF_BENCH_VERBOSE=1 timeout 900 python bench.py > $OUT/r3i_bench_default.json 2> $OUT/r3i_bench_default.err; echo "bench default exit $?"
python -c "
import json; d=json.load(open('$OUT/r3i_bench_default.json'))
print('train ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'fwd ms', d['forward']['ms_per_step'], 'launches', d['launches_per_step'])
print('roofline', d['roofline']); print('clocks', d['clocks']); print(json.dumps(d.get('kernel_time'))[:900])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r3i_bench_reference.json 2> $OUT/r3i_bench_reference.err; echo "reference arm exit $?"; head -c 300 $OUT/r3i_bench_reference.json; echo
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv \
    --log-file $OUT/r3i_train_launches_warm_bf16x2.csv python tools/ncu_train_target.py --precision bf16x2 > $OUT/r3i_ncu_warm_train.log 2>&1; echo "ncu warm train exit $?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 2000 --csv \
    --log-file $OUT/r3i_fwd_launches_warm_bf16x2.csv python tools/ncu_target.py --precision bf16x2 > $OUT/r3i_ncu_warm_fwd.log 2>&1; echo "ncu warm fwd exit $?"
ls -la $OUT | tail -12
