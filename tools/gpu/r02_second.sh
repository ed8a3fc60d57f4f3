#!/bin/bash
# fbank v4 parity + timing against v2, real-shape tests with calibrated tolerances
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "fbank" > gpurun_out/r02_fbank_tests.log 2>&1
echo "fbank tests rc=$?"; tail -15 gpurun_out/r02_fbank_tests.log
for v in v4 v2; do PB200_FBANK=$v timeout 120 python tools/fbank_one.py 20 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_real_shapes_gpu.py -m gpu -q -s > gpurun_out/r02_real_shapes.log 2>&1
echo "real shapes rc=$?"; grep -E "gradients,|activations|max-abs|passed|failed|^FAILED|^E  " gpurun_out/r02_real_shapes.log | cut -c1-400 | head -60
