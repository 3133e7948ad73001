#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/fwd_options.py > $OUT/r3b_fwd_options.txt 2>&1; echo "fwd options exit $?"; cat $OUT/r3b_fwd_options.txt | grep fuse_ln
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > $OUT/r3b_bench.json 2> $OUT/r3b_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r3b_bench.json')); print('  ms', round(d['ms_per_step'],3)); print(json.dumps(d.get('kernel_time'))[:1500]); print(d['roofline']['achieved'], d['roofline']['frac'])"
