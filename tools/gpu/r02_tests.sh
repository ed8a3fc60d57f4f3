#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_2.log 2>&1
echo "gpu tests rc=$?"; grep -E "^FAILED|passed|failed|^E  +Assert" gpurun_out/r02_tests_2.log | cut -c1-700 | tail -20
