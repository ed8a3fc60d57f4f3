#!/bin/bash
# fused bias-gradient check: kernel test first (bounded), then the full GPU suite and a bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_bias or colsum or gemm" > gpurun_out/rs_kernels.log 2>&1
echo "kernels rc=$?" | tee -a gpurun_out/rs_kernels.log
tail -5 gpurun_out/rs_kernels.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/rs_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/rs_tests.log
tail -5 gpurun_out/rs_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/rs_bench.json 2> gpurun_out/rs_bench.err
echo "bench rc=$?"
cat gpurun_out/rs_bench.json
timeout 300 python tools/step_breakdown.py > gpurun_out/rs_breakdown.txt 2>&1
head -30 gpurun_out/rs_breakdown.txt
