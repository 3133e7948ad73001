#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch lists and a full-set capture; everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
# Every stage runs under its own timeout so that a hang costs one stage, not the box.
TAG=${1:-r1}
STAGES=${2:-test,bench,launches,full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [[ $STAGES == *test* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
if [[ $STAGES == *bench* ]]; then
  timeout 400 python bench.py > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err
  echo "bench bf16 exit $?"; head -c 600 $OUT/${TAG}_bench_bf16.json; echo
  timeout 400 python bench.py --precision tf32 > $OUT/${TAG}_bench_tf32.json 2> $OUT/${TAG}_bench_tf32.err
  echo "bench tf32 exit $?"; head -c 400 $OUT/${TAG}_bench_tf32.json; echo
fi
if [[ $STAGES == *launches* ]]; then
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_coldcache_bf16.csv python tools/ncu_target.py > $OUT/${TAG}_ncu_cold.log 2>&1
  echo "ncu cold exit $?"
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_warmcache_bf16.csv python tools/ncu_target.py > $OUT/${TAG}_ncu_warm.log 2>&1
  echo "ncu warm exit $?"
fi
if [[ $STAGES == *full* ]]; then
  # front end + every kernel of block 0 (the 15 positional GEMMs on the side stream come first in launch order)
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k 'regex:gemm_tc|ffn_fused|relpos_attn|dwconv|subsample' -s 15 -c 10 -o $OUT/${TAG}_full_block0_bf16 -f \
    python tools/ncu_target.py > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit $?"
  ncu -i $OUT/${TAG}_full_block0_bf16.ncu-rep --page raw --csv > $OUT/${TAG}_full_block0_bf16_raw.csv 2>/dev/null
fi
ls -la $OUT | tail -20
