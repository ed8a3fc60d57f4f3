#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "whisper_variant" > gpurun_out/s2_wsp.log 2>&1; echo "whisper rc=$?"; grep -E "^E   |^pasero_b200/|^tests/.*Error|passed|failed" gpurun_out/s2_wsp.log | cut -c1-260 | head -30
timeout 900 python -m pytest tests/test_fp16_gpu.py -m gpu -q > gpurun_out/s2_fp16.log 2>&1; echo "fp16 rc=$?"; grep -E "passed|failed|^FAILED|^E   " gpurun_out/s2_fp16.log | cut -c1-260 | head -30
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_fp16_gpu.py > gpurun_out/s2_pytest_c.log 2>&1; echo "bf16 suite rc=$?"; tail -4 gpurun_out/s2_pytest_c.log | cut -c1-300
