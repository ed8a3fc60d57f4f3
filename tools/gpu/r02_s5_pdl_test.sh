#!/bin/bash
# session 5: the PB200_PDL switch test alone
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "pdl_switch" > gpurun_out/s5_pdl_test.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/s5_pdl_test.log | cut -c1-400
