#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -m gpu -q -x -k "attention" > gpurun_out/s2_attn_tests.log 2>&1; echo "attention tests rc=$?"; tail -3 gpurun_out/s2_attn_tests.log | cut -c1-300
timeout 300 python tools/microbench.py attn 2>&1 | tail -7
