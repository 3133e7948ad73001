#!/bin/bash
# 2-GPU box: the whole GPU suite (2-GPU tests included) after the PDL conversion, headline at N = 1 and N = 2
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/r3h_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r3h_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > $OUT/r3h_bench_1gpu.json 2> $OUT/r3h_bench_1gpu.err; echo "bench 1gpu exit $?"; python -c "
import json; d=json.load(open('$OUT/r3h_bench_1gpu.json')); print('  ms', round(d['ms_per_step'],3), d['step_ms_min_med_max'], 'launches', d['launches_per_step'], 'loss', d['loss_first_last']); [print('   ', o) for o in d['operators'][:14]]"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --no-extras --no-cpu-baseline --steps 30 > $OUT/r3h_bench_2gpu.json 2> $OUT/r3h_bench_2gpu.err; echo "bench 2gpu exit $?"
python -c "
import json
d=json.loads([l for l in open('$OUT/r3h_bench_2gpu.json') if l.startswith('{')][-1])
print('  ms', round(d['ms_per_step'],3), d.get('sync_bn_exchange'), 'timeouts', d.get('sync_bn_exchange_timeouts'), 'exposed', round(d['communication']['exposed_ms_per_step'],3), 'local', round(d['communication']['ms_per_step_no_collectives'],3), d['loss_first_last'])"
