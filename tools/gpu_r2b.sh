#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1000 python -m pytest tests -m gpu -q -x -k "ops or backward" > $OUT/r2b_pytest_ops.log 2>&1; echo "pytest ops exit $?" | tee -a $OUT/r2b_pytest_ops.log; tail -5 $OUT/r2b_pytest_ops.log
timeout 1000 python -m pytest tests -m gpu -q -s -k "not ops and not backward" > $OUT/r2b_pytest_rest.log 2>&1; echo "pytest rest exit $?" | tee -a $OUT/r2b_pytest_rest.log; tail -15 $OUT/r2b_pytest_rest.log
for pr in bf16x2 bf16; do
  timeout 300 python bench.py --precision $pr --no-cpu-baseline > $OUT/r2b_bench_train_$pr.json 2> $OUT/r2b_bench_train_$pr.err; echo "bench train $pr exit $?"; head -c 300 $OUT/r2b_bench_train_$pr.json; echo
done
timeout 300 python bench.py --mode forward --precision bf16x2 --no-cpu-baseline > $OUT/r2b_bench_fwd_bf16x2.json 2> $OUT/r2b_bench_fwd_bf16x2.err; echo "bench fwd exit $?"; head -c 300 $OUT/r2b_bench_fwd_bf16x2.json; echo
