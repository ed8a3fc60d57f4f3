#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention or fbank" > gpurun_out/r02_attn_tests.log 2>&1
echo "attention+fbank kernel tests rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02_attn_tests.log | cut -c1-300 | tail -15
timeout 120 python tools/fbank_one.py 20 2>&1 | tail -1
timeout 900 python -m pytest tests/test_real_shapes_gpu.py -m gpu -q -s > gpurun_out/r02_real_shapes.log 2>&1
echo "real shapes rc=$?"; grep -E "gradients,|yardstick|passed|failed|^FAILED|^E  +Assert" gpurun_out/r02_real_shapes.log | cut -c1-1200 | head -30
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_real_shapes_gpu.py > gpurun_out/r02_tests_all.log 2>&1
echo "all other gpu tests rc=$?"; tail -5 gpurun_out/r02_tests_all.log | cut -c1-300
