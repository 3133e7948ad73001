#!/bin/bash
# round 2, visit E (re-entry): state of the tree on the B200.  gpurun --timeout 1500 -- 'bash tools/gpu_r2e.sh'
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/r2e_gpu.txt 2>&1
timeout 1000 python -m pytest tests -m gpu -q -s > $OUT/r2e_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2e_pytest.log; grep -n "^\[\|drop-in\|passed\|failed\|FAILED\|Error" $OUT/r2e_pytest.log | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2e_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/r2e_smoke.log
timeout 400 python bench.py > $OUT/r2e_bench_default.json 2> $OUT/r2e_bench_default.err; echo "bench default exit $?"; head -c 600 $OUT/r2e_bench_default.json; echo
for pr in bf16 tf32; do
  timeout 300 python bench.py --precision $pr --no-cpu-baseline > $OUT/r2e_bench_train_$pr.json 2> $OUT/r2e_bench_train_$pr.err; echo "bench train $pr exit $?"; head -c 300 $OUT/r2e_bench_train_$pr.json; echo
done
for pr in bf16x2 bf16 tf32; do
timeout 300 python bench.py --mode forward --precision $pr --no-cpu-baseline > $OUT/r2e_bench_fwd_$pr.json 2> $OUT/r2e_bench_fwd_$pr.err; echo "bench fwd exit $?"; head -c 300 $OUT/r2e_bench_fwd_$pr.json; echo
done
