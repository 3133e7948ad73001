#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "ctc" > $OUT/r2l_pytest_ctc.log 2>&1; echo "pytest ctc exit $?"; tail -5 $OUT/r2l_pytest_ctc.log
timeout 600 python tools/gemm_tiles.py bf16x2 > $OUT/r2l_gemm_tiles_bf16x2.txt 2>&1; echo "tiles exit $?"; cat $OUT/r2l_gemm_tiles_bf16x2.txt
timeout 600 python tools/gemm_tiles.py bf16 > $OUT/r2l_gemm_tiles_bf16.txt 2>&1; echo "tiles exit $?"; cat $OUT/r2l_gemm_tiles_bf16.txt
