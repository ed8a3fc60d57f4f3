#!/bin/bash
# session 5: GEMM launch-choice sweep over the step's shapes (tools/gemm_sweep.py)
mkdir -p gpurun_out
timeout 500 python tools/gemm_sweep.py > gpurun_out/s5_gemm_sweep.txt 2> gpurun_out/s5_gemm_sweep.err; echo "rc=$?"
cat gpurun_out/s5_gemm_sweep.txt | cut -c1-260; tail -5 gpurun_out/s5_gemm_sweep.err | cut -c1-200
