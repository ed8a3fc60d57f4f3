#!/bin/bash
# poor man's memcheck (compute-sanitizer is closed on this pool): the GPU suite under allocator configurations in which memory
# past the end of an allocation is more often unmapped — expandable segments (whole suite), and no caching at all (every tensor
# its own cudaMalloc; CUDA graphs need the caching allocator, so graph tests are left out)
mkdir -p gpurun_out
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s4_alloc_expandable.log 2>&1; echo "expandable rc=$? $(tail -1 gpurun_out/s4_alloc_expandable.log | cut -c1-120)"
grep -E "^FAILED|^ERROR" gpurun_out/s4_alloc_expandable.log | head -10 | cut -c1-200
PYTORCH_NO_CUDA_MEMORY_CACHING=1 CUDA_LAUNCH_BLOCKING=1 timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not graph and not bench_shape and not inference_step" > gpurun_out/s4_alloc_nocache.log 2>&1; echo "no-cache rc=$? $(tail -1 gpurun_out/s4_alloc_nocache.log | cut -c1-120)"
grep -E "^FAILED|^ERROR" gpurun_out/s4_alloc_nocache.log | head -10 | cut -c1-200
