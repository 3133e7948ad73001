#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_transducer.py -m gpu -q -s > $OUT/r2w_pytest_rnnt.log 2>&1; echo "pytest exit $?"; grep -n "joint logits\|per-utterance\|passed\|failed\|FAILED\|^E  " $OUT/r2w_pytest_rnnt.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2w_smoke.log 2>&1; echo "smoke exit $?"; grep smoke: $OUT/r2w_smoke.log
