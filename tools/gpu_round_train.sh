#!/bin/bash
# One GPU-box visit for the training step: parity tests, bench lines (train + forward), ncu launch list and a full-set capture of the
# dominant tensor-core kernels of one training step; everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_train.sh [tag] [stages]'
TAG=${1:-r1t}
STAGES=${2:-test,bench,launches,full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [[ $STAGES == *test* ]]; then
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -6 $OUT/${TAG}_pytest.log
fi
if [[ $STAGES == *bench* ]]; then
  timeout 400 python bench.py > $OUT/${TAG}_bench_train_bf16.json 2> $OUT/${TAG}_bench_train_bf16.err
  echo "bench train bf16 exit $?"; head -c 500 $OUT/${TAG}_bench_train_bf16.json; echo
  timeout 300 python bench.py --mode forward --no-cpu-baseline > $OUT/${TAG}_bench_forward_bf16.json 2> $OUT/${TAG}_bench_forward_bf16.err
  echo "bench forward bf16 exit $?"; head -c 300 $OUT/${TAG}_bench_forward_bf16.json; echo
fi
if [[ $STAGES == *launches* ]]; then
  timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv \
    --log-file $OUT/${TAG}_train_launches_warm_bf16.csv python tools/ncu_train_target.py > $OUT/${TAG}_ncu_warm.log 2>&1
  echo "ncu warm exit $?"
fi
if [[ $STAGES == *full* ]]; then
  timeout 500 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k 'regex:wgrad_tc_kernel|gemm_tc_kernel|bgemm_kernel|tc_rows_kernel' -s 40 -c 14 -o $OUT/${TAG}_full_train_bf16 -f \
    python tools/ncu_train_target.py > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit $?"
  ncu -i $OUT/${TAG}_full_train_bf16.ncu-rep --page raw --csv > $OUT/${TAG}_full_train_bf16_raw.csv 2>/dev/null
fi
ls -la $OUT | tail -12
