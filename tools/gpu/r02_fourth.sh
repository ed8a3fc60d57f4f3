#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "fbank" > gpurun_out/r02_fbank_tests.log 2>&1
echo "fbank tests rc=$?"; grep -E "fbank raw|fbank cmvn|passed|failed" gpurun_out/r02_fbank_tests.log | tail -12
timeout 120 python tools/fbank_one.py 20 2>&1 | tail -1
timeout 900 python -m pytest tests/test_real_shapes_gpu.py -m gpu -q -s > gpurun_out/r02_real_shapes.log 2>&1
echo "real shapes rc=$?"; grep -E "gradients,|yardstick|max-abs|passed|failed|^FAILED|^E  +Assert" gpurun_out/r02_real_shapes.log | cut -c1-1500 | head -40
timeout 120 python tools/fbank_one.py 3 > gpurun_out/r02_fbank_v4.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fbank|cmvn" -s 4 -c 2 -o gpurun_out/r02_ncu_fbank_v4b -f python tools/fbank_one.py 2 > gpurun_out/r02_ncu_fbank_v4b.log 2>&1
echo "ncu rc=$?"
