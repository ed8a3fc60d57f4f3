#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_real_shapes_gpu.py -m gpu -q -k "iwslt2023" > gpurun_out/s2_wide.log 2>&1; echo "new tests rc=$?"; grep -E "passed|failed|^FAILED|^E   " gpurun_out/s2_wide.log | cut -c1-400 | head -20
