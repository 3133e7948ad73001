#!/bin/bash
# device front end tests + timings, ncu full-set capture of the fused subsampling kernel and the log-mel kernel, default bench refresh
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_frontend.py -m gpu -q -s > $OUT/r3e_pytest_frontend.log 2>&1; echo "pytest frontend exit $?"; grep -n "log-mel\|audio ->\|mean masked\|passed\|failed\|FAILED\|^E  " $OUT/r3e_pytest_frontend.log | head -40
timeout 300 python tools/ncu_frontend_target.py > $OUT/r3e_frontend_timing.log 2>&1; echo "frontend timing exit $?"; tail -3 $OUT/r3e_frontend_timing.log
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r3e_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3e_pytest.log
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none -k 'regex:subsample_linear_fused' -c 1 -o /tmp/r3e_full_front -f \
    python tools/ncu_target.py --precision bf16x2 > $OUT/r3e_ncu_front.log 2>&1
echo "ncu front exit $?"
ncu -i /tmp/r3e_full_front.ncu-rep --page raw --csv > $OUT/r3e_full_front_bf16x2_raw.csv 2>/dev/null
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:logmel_kernel|specaugment_kernel' -s 6 -c 2 -o /tmp/r3e_full_logmel -f \
    python tools/ncu_frontend_target.py > $OUT/r3e_ncu_logmel.log 2>&1
echo "ncu logmel exit $?"
ncu -i /tmp/r3e_full_logmel.ncu-rep --page raw --csv > $OUT/r3e_full_logmel_raw.csv 2>/dev/null
timeout 900 python bench.py > $OUT/r3e_bench_default.json 2> $OUT/r3e_bench_default.err; echo "bench default exit $?"; python -c "
import json; d=json.load(open('$OUT/r3e_bench_default.json')); print('  ms', round(d['ms_per_step'],3), 'value', round(d['value']), 'fwd', d.get('forward',{}).get('ms_per_step'))"
ls -la $OUT | tail -12
