#!/bin/bash
# ncu full-set capture of the final log-mel (Stockham) and SpecAugment kernels
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:logmel_kernel|specaugment_kernel' -s 6 -c 2 -o /tmp/r3n_full_logmel -f \
    python tools/ncu_frontend_target.py > $OUT/r3n_ncu_logmel.log 2>&1
echo "ncu logmel exit $?"
ncu -i /tmp/r3n_full_logmel.ncu-rep --page raw --csv > $OUT/r3n_full_logmel_raw.csv 2>/dev/null
timeout 200 python tools/ncu_frontend_target.py > $OUT/r3n_frontend_timing.log 2>&1; tail -3 $OUT/r3n_frontend_timing.log
