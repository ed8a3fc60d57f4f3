#!/bin/bash
# poor man's memcheck (compute-sanitizer is closed on this pool): no caching allocator, so every tensor is its own cudaMalloc
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 CUDA_LAUNCH_BLOCKING=1 timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_decoding_gpu.py -m gpu -q -x \
  -k "not fbank and not real" > gpurun_out/s2_nocache.log 2>&1; echo "no-cache rc=$?"
grep -E "passed|failed|^FAILED|^E   |Error" gpurun_out/s2_nocache.log | cut -c1-250 | head -20
