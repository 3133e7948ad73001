#!/bin/bash
# round 2, visit A: parity of the split (bf16x2) mode + first bench lines.  gpurun --timeout 1700 -- 'bash tools/gpu_r2a.sh'
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/r2a_gpu.txt 2>&1
timeout 1000 python -m pytest tests -m gpu -q -x -k "ops or backward" > $OUT/r2a_pytest_ops.log 2>&1; echo "pytest ops exit $?" | tee -a $OUT/r2a_pytest_ops.log; tail -5 $OUT/r2a_pytest_ops.log
timeout 1000 python -m pytest tests -m gpu -q -s -k "not ops and not backward" > $OUT/r2a_pytest_rest.log 2>&1; echo "pytest rest exit $?" | tee -a $OUT/r2a_pytest_rest.log; tail -15 $OUT/r2a_pytest_rest.log
for pr in bf16x2 bf16; do
  timeout 300 python bench.py --precision $pr --no-cpu-baseline > $OUT/r2a_bench_train_$pr.json 2> $OUT/r2a_bench_train_$pr.err; echo "bench train $pr exit $?"; head -c 300 $OUT/r2a_bench_train_$pr.json; echo
  timeout 300 python bench.py --mode forward --precision $pr --no-cpu-baseline > $OUT/r2a_bench_fwd_$pr.json 2> $OUT/r2a_bench_fwd_$pr.err; echo "bench fwd $pr exit $?"; head -c 300 $OUT/r2a_bench_fwd_$pr.json; echo
done
