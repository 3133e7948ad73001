#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/r2d_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2d_pytest.log; grep -n "^\[\|drop-in\|passed\|failed\|FAILED" $OUT/r2d_pytest.log | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2d_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/r2d_smoke.log
for pr in bf16x2 bf16; do
  timeout 300 python bench.py --precision $pr --no-cpu-baseline > $OUT/r2d_bench_train_$pr.json 2> $OUT/r2d_bench_train_$pr.err; echo "bench train $pr exit $?"; head -c 300 $OUT/r2d_bench_train_$pr.json; echo
done
timeout 300 python bench.py --mode forward --precision bf16x2 --no-cpu-baseline > $OUT/r2d_bench_fwd_bf16x2.json 2> $OUT/r2d_bench_fwd_bf16x2.err; echo "bench fwd exit $?"; head -c 300 $OUT/r2d_bench_fwd_bf16x2.json; echo
