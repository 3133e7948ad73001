#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_transducer.py -m gpu -q -s > $OUT/r2x_pytest_rnnt.log 2>&1; echo "pytest exit $?"; grep -n "joint \|RNN-T\|per-utterance\|passed\|failed\|FAILED\|^E  " $OUT/r2x_pytest_rnnt.log | head -40
