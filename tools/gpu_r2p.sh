#!/bin/bash
# ncu full-set captures of the dominant kernels (training step and forward, default bf16x2 mode) + in-kernel GEMM timeline
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/gemm_timeline.py bf16x2 > $OUT/r2p_gemm_timeline_bf16x2.txt 2>&1; echo "timeline exit $?"; cat $OUT/r2p_gemm_timeline_bf16x2.txt | tail -12
timeout 700 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k 'regex:gemm_tc_kernel|wgrad_tc_kernel|bgemm_kernel|tc_rows_kernel|layernorm_bwd_kernel|dwconv_run_kernel|dwconv_bwd_weight' -s 60 -c 30 -o $OUT/r2p_full_train_bf16x2 -f \
    python tools/ncu_train_target.py --precision bf16x2 > $OUT/r2p_ncu_full_train.log 2>&1
echo "ncu full train exit $?"
ncu -i $OUT/r2p_full_train_bf16x2.ncu-rep --page raw --csv > $OUT/r2p_full_train_bf16x2_raw.csv 2>/dev/null
timeout 500 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k 'regex:gemm_tc|relpos_attn|dwconv_bn|subsample' -s 15 -c 14 -o $OUT/r2p_full_fwd_bf16x2 -f \
    python tools/ncu_target.py --precision bf16x2 > $OUT/r2p_ncu_full_fwd.log 2>&1
echo "ncu full fwd exit $?"
ncu -i $OUT/r2p_full_fwd_bf16x2.ncu-rep --page raw --csv > $OUT/r2p_full_fwd_bf16x2_raw.csv 2>/dev/null
ls -la $OUT | grep r2p
