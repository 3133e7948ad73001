#!/bin/bash
# 2 GPUs: symmetric-memory probe, peer-memory SyncBatchNorm exchange in the 2-GPU tests, bench at N = 2 with and without it
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/probe_symm.py > $OUT/r2s_probe.log 2>&1; echo "probe exit $?"; grep -n "SYMM\|rendezvous\|peer\|Error\|error" $OUT/r2s_probe.log | head -12
timeout 900 python -m pytest tests -m gpu -q -s -k "two_gpu or distribute_strategy" > $OUT/r2s_pytest_2gpu.log 2>&1; echo "pytest exit $?"; grep -n "2-rank\|2 GPUs\|passed\|failed\|FAILED\|skipped\|unavailable" $OUT/r2s_pytest_2gpu.log | tail
for p2p in 1 0; do
EFFCONF_P2P_BN=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$p2p bench.py --gpus 2 --no-extras --steps 30 > $OUT/r2s_bench_2gpu_p2p$p2p.json 2> $OUT/r2s_bench_2gpu_p2p$p2p.err; echo "bench 2gpu p2p=$p2p exit $?"
python -c "
import json
d=json.loads([l for l in open('$OUT/r2s_bench_2gpu_p2p$p2p.json') if l.startswith('{')][-1])
print('  ms', round(d['ms_per_step'],3), d.get('sync_bn_exchange'), 'timeouts', d.get('sync_bn_exchange_timeouts'), 'exposed', round(d['communication']['exposed_ms_per_step'],3), 'local', round(d['communication']['ms_per_step_no_collectives'],3))"
done
