#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s > $OUT/r2g_pytest_dropin.log 2>&1; echo "pytest exit $?"; grep -n "GradScaler\|drop-in\|passed\|failed\|FAILED" $OUT/r2g_pytest_dropin.log | tail
EFFCONF_BENCH_VERBOSE=1 timeout 900 python bench.py > $OUT/r2g_bench_default.json 2> $OUT/r2g_bench_default.err; echo "bench default exit $?"; head -c 400 $OUT/r2g_bench_default.json; echo; grep "bench rank" $OUT/r2g_bench_default.err | tail -12
