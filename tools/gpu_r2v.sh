#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_transducer.py -m gpu -q -s > $OUT/r2v_pytest_rnnt.log 2>&1; echo "pytest exit $?"; grep -n "joint logits\|per-utterance\|passed\|failed\|FAILED\|^E  " $OUT/r2v_pytest_rnnt.log | head -30
EFFCONF_BENCH_VERBOSE=1 timeout 900 python bench.py > $OUT/r2v_bench_default.json 2> $OUT/r2v_bench_default.err; echo "bench default exit $?"; grep "bench rank" $OUT/r2v_bench_default.err | tail -9
python -c "
import json; d=json.load(open('$OUT/r2v_bench_default.json'))
print('train ms', d['ms_per_step'], 'fwd ms', d['forward']['ms_per_step'])
print(json.dumps(d.get('transducer_joint_forward'))[:600])
print(json.dumps(d['configs'].get('bf16_mode'))[:1500])"
