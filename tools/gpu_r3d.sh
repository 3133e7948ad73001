#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -s -k "fused_front" > $OUT/r3d_pytest_front.log 2>&1; echo "pytest front exit $?"; grep -n "fused vs\|passed\|failed\|FAILED\|^E  " $OUT/r3d_pytest_front.log | head -30
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r3d_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3d_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --mode forward --steps 50 > $OUT/r3d_bench_fwd.json 2> $OUT/r3d_bench_fwd.err; echo "bench fwd exit $?"; python -c "
import json; d=json.load(open('$OUT/r3d_bench_fwd.json')); print('  fwd ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step']); [print('   ', k) for k in d['kernels'][:4]]"
timeout 300 python bench.py --no-extras --no-cpu-baseline --mode forward --steps 50 --precision bf16 > $OUT/r3d_bench_fwd_bf16.json 2> $OUT/r3d_bench_fwd_bf16.err; echo "bench fwd bf16 exit $?"; python -c "
import json; d=json.load(open('$OUT/r3d_bench_fwd_bf16.json')); print('  fwd bf16 ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step']); [print('   ', k) for k in d['kernels'][:3]]"
