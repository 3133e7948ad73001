#!/bin/bash
# 2 GPUs: overlapped gradient buckets (2-GPU parity tests, headline with / without the overlap) + the Stockham log-mel kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_frontend.py -m gpu -q -s > $OUT/r3f_pytest_frontend.log 2>&1; echo "pytest frontend exit $?"; grep -n "log-mel\|passed\|failed\|FAILED\|^E  " $OUT/r3f_pytest_frontend.log | head -20
timeout 200 python tools/ncu_frontend_target.py > $OUT/r3f_frontend_timing.log 2>&1; echo "frontend timing exit $?"; tail -3 $OUT/r3f_frontend_timing.log
timeout 600 python -m pytest tests -m gpu -q -s -k "two_gpu or distribute_strategy" > $OUT/r3f_pytest_2gpu.log 2>&1; echo "pytest 2gpu exit $?"; grep -n "2-rank\|2 GPUs\|passed\|failed\|FAILED\|skipped\|unavailable\|^E  " $OUT/r3f_pytest_2gpu.log | tail
for ov in 1 0; do
EFFCONF_BUCKET_OVERLAP=$ov timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$ov bench.py --gpus 2 --no-extras --no-cpu-baseline --steps 30 > $OUT/r3f_bench_2gpu_ov$ov.json 2> $OUT/r3f_bench_2gpu_ov$ov.err; echo "bench 2gpu overlap=$ov exit $?"
python -c "
import json
d=json.loads([l for l in open('$OUT/r3f_bench_2gpu_ov$ov.json') if l.startswith('{')][-1])
print('  ms', round(d['ms_per_step'],3), d.get('sync_bn_exchange'), 'exposed', round(d['communication']['exposed_ms_per_step'],3), 'local', round(d['communication']['ms_per_step_no_collectives'],3), d['loss_first_last'])"
tail -3 $OUT/r3f_bench_2gpu_ov$ov.err
done
