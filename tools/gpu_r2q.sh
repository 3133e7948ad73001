#!/bin/bash
# ncu full-set captures (reports stay in /tmp on the box: only the CSV pages come back) + tests / bench of the vectorised pack kernels
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_encoder.py -m gpu -q -k "attention or training_step or transducer_small or graph_replay" > $OUT/r2q_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r2q_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > $OUT/r2q_bench.json 2> $OUT/r2q_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r2q_bench.json')); print('  ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step']); [print('   ', o) for o in d['operators'][:5]]"
timeout 700 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k 'regex:gemm_tc_kernel|wgrad_tc_kernel|bgemm_kernel|tc_rows_kernel|layernorm_bwd_kernel|dwconv_run_kernel' -s 70 -c 20 -o /tmp/r2q_full_train_bf16x2 -f \
    python tools/ncu_train_target.py --precision bf16x2 > $OUT/r2q_ncu_full_train.log 2>&1
echo "ncu full train exit $?"
ncu -i /tmp/r2q_full_train_bf16x2.ncu-rep --page raw --csv > $OUT/r2q_full_train_bf16x2_raw.csv 2>/dev/null
timeout 500 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k 'regex:gemm_tc|relpos_attn|dwconv_bn|subsample' -s 15 -c 13 -o /tmp/r2q_full_fwd_bf16x2 -f \
    python tools/ncu_target.py --precision bf16x2 > $OUT/r2q_ncu_full_fwd.log 2>&1
echo "ncu full fwd exit $?"
ncu -i /tmp/r2q_full_fwd_bf16x2.ncu-rep --page raw --csv > $OUT/r2q_full_fwd_bf16x2_raw.csv 2>/dev/null
ls -la $OUT
