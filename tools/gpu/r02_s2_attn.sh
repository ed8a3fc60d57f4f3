#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -m gpu -q -k "attention" > gpurun_out/s2_attn_tests.log 2>&1; echo "attention tests rc=$?"; grep -E "^FAILED|passed|failed|AssertionError" gpurun_out/s2_attn_tests.log | cut -c1-200 | head -30
for i in 1 2 3; do timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_attention and 130" 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-160; done
timeout 300 python tools/microbench.py attn 2>&1 | tail -7
