#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "gemm_train" > $OUT/r2h_pytest_gemm_train.log 2>&1; echo "pytest gemm_train exit $?"; tail -15 $OUT/r2h_pytest_gemm_train.log
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_dropin.py tests/test_gpu_backward.py -m gpu -q -s > $OUT/r2h_pytest_train.log 2>&1; echo "pytest train exit $?"; grep -n "^\[\|passed\|failed\|FAILED\|Error" $OUT/r2h_pytest_train.log | tail -20
timeout 300 python bench.py --no-extras --no-cpu-baseline > $OUT/r2h_bench_train.json 2> $OUT/r2h_bench_train.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r2h_bench_train.json')); print(d['ms_per_step'], d['launches_per_step']); [print(o) for o in d['operators'][:14]]"
