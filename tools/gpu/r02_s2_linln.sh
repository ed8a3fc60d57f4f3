#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "linear_dropout_add_ln_fused" > gpurun_out/s2_linln_test.log 2>&1; echo "test rc=$?"; tail -4 gpurun_out/s2_linln_test.log | cut -c1-300
timeout 200 python tools/mb_linear_ln.py 2>&1 | tail -6
