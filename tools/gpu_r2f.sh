#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/debug_autocast.py > $OUT/r2f_debug_autocast.log 2>&1; echo "debug exit $?"; grep -v Warning $OUT/r2f_debug_autocast.log | tail -40
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv \
    --log-file $OUT/r2f_launches_warm_train_bf16x2.csv python tools/ncu_train_target.py --precision bf16x2 > $OUT/r2f_ncu_warm.log 2>&1
echo "ncu warm train exit $?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv \
    --log-file $OUT/r2f_launches_warm_fwd_bf16x2.csv python tools/ncu_target.py --precision bf16x2 > $OUT/r2f_ncu_warm_fwd.log 2>&1
echo "ncu warm fwd exit $?"
