#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1000 python -m pytest tests -m gpu -q -k "ops or backward or training" > $OUT/r2c_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2c_pytest.log; tail -8 $OUT/r2c_pytest.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv \
    --log-file $OUT/r2c_launches_warm_train_bf16x2.csv python tools/ncu_train_target.py --precision bf16x2 > $OUT/r2c_ncu_warm.log 2>&1
echo "ncu warm exit $?"
