#!/bin/bash
# 8 GPUs: the headline step with the peer-memory SyncBatchNorm exchange (and, for comparison, NCCL), no extras
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
for p2p in 1 0; do
EFFCONF_P2P_BN=$p2p timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2963$p2p bench.py --gpus 8 --no-extras --steps 30 > $OUT/r2y_bench_8gpu_p2p$p2p.json 2> $OUT/r2y_bench_8gpu_p2p$p2p.err; echo "bench 8gpu p2p=$p2p exit $?"
python -c "
import json
d=json.loads([l for l in open('$OUT/r2y_bench_8gpu_p2p$p2p.json') if l.startswith('{')][-1])
print('  ms', round(d['ms_per_step'],3), 'value', round(d['value']), d.get('sync_bn_exchange'), 'timeouts', d.get('sync_bn_exchange_timeouts'), 'exposed', round(d['communication']['exposed_ms_per_step'],3), 'local', round(d['communication']['ms_per_step_no_collectives'],3))" || tail -5 $OUT/r2y_bench_8gpu_p2p$p2p.err
done
