#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoding_gpu.py -m gpu -q > gpurun_out/s2_dec.log 2>&1; echo "decoding rc=$?"; grep -E "passed|failed|^FAILED|^E   " gpurun_out/s2_dec.log | cut -c1-300 | head -30
