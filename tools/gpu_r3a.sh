#!/bin/bash
# 2 GPUs: the FULL default bench line under torchrun (extras with the peer-memory exchange, orderly teardown), then the 2-GPU tests
OUT=gpurun_out; mkdir -p $OUT
START=$(date +%s)
EFFCONF_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > $OUT/r3a_bench_2gpu.json 2> $OUT/r3a_bench_2gpu.err; echo "bench 2gpu exit $? after $(( $(date +%s) - START )) s"
grep "bench rank 0\|teardown" $OUT/r3a_bench_2gpu.err | tail -10
python -c "
import json
d=json.loads([l for l in open('$OUT/r3a_bench_2gpu.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'value', d['value'], d.get('sync_bn_exchange'), d.get('communication'))
for k in ('forward','sweep','modes','ragged'):
    v=d.get(k); print(k, 'ERROR '+str(v.get('error')) if isinstance(v,dict) and 'error' in v else 'ok')"
timeout 900 python -m pytest tests -m gpu -q -k "two_gpu or distribute_strategy" > $OUT/r3a_pytest_2gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3a_pytest_2gpu.log
