#!/bin/bash
# the GPU suite in other file orders than the alphabetical one (allocator placement and leftover global state differ)
mkdir -p gpurun_out
T=tests
timeout 900 python -m pytest $T/test_real_shapes_gpu.py $T/test_parity_gpu.py $T/test_kernels_gpu.py $T/test_fp16_gpu.py $T/test_decoding_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/s4_order_a.log 2>&1; echo "order A rc=$? $(tail -1 gpurun_out/s4_order_a.log | cut -c1-120)"
grep -E "^FAILED|^ERROR" gpurun_out/s4_order_a.log | head -10 | cut -c1-200
timeout 900 python -m pytest $T/test_decoding_gpu.py $T/test_fp16_gpu.py $T/test_real_shapes_gpu.py $T/test_kernels_gpu.py $T/test_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/s4_order_b.log 2>&1; echo "order B rc=$? $(tail -1 gpurun_out/s4_order_b.log | cut -c1-120)"
grep -E "^FAILED|^ERROR" gpurun_out/s4_order_b.log | head -10 | cut -c1-200
