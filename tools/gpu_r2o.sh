#!/bin/bash
# 2 GPUs: the reference's distribute_strategy through the drop-in, the native data-parallel step, and the bench line at N = 2
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/r2o_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "two_gpu or distribute_strategy" > $OUT/r2o_pytest_2gpu.log 2>&1; echo "pytest exit $?"; grep -n "2-rank\|passed\|failed\|FAILED\|skipped" $OUT/r2o_pytest_2gpu.log | tail
EFFCONF_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > $OUT/r2o_bench_2gpu.json 2> $OUT/r2o_bench_2gpu.err; echo "bench 2gpu exit $?"; head -c 300 $OUT/r2o_bench_2gpu.json; echo; grep "bench rank 0" $OUT/r2o_bench_2gpu.err | tail -8
python -c "
import json
d=json.loads([l for l in open('$OUT/r2o_bench_2gpu.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'value', d['value']); print(d.get('communication')); print({k:(v if not isinstance(v,dict) else '...') for k,v in d.get('forward',{}).items() if k in ('value','ms_per_step','error')})
print([ (r['frames'], round(r['forward']['value']), round(r['train']['value'])) for r in d['sweep']['rows']] if 'rows' in d.get('sweep',{}) else d.get('sweep'))"
