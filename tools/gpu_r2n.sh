#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/r2n_pytest.log 2>&1; echo "pytest exit $?"; grep -n "^E  \|passed\|failed\|FAILED" $OUT/r2n_pytest.log | head -20
