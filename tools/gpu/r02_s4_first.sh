#!/bin/bash
# session 4, call 1: the per-language adapter (`*_adapters_by`) Trainer tests, then the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "adapters_by" > gpurun_out/s4_by.log 2>&1; echo "adapters_by rc=$?"; tail -30 gpurun_out/s4_by.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s4_pytest_full.log | cut -c1-300
