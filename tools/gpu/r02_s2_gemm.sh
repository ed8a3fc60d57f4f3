#!/bin/bash
# per-shape GEMM table of the step (in situ) + the microbench of the model's shapes against cuBLAS
mkdir -p gpurun_out
BENCH_GEMM_DUMP=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/s2_gemm_dump.json 2> gpurun_out/s2_gemm_dump.txt; echo "dump rc=$?"
grep -E "TFLOP/s" gpurun_out/s2_gemm_dump.txt | head -70
timeout 600 python tools/microbench.py gemm > gpurun_out/s2_microbench_gemm.log 2>&1; echo "mb rc=$?"; tail -70 gpurun_out/s2_microbench_gemm.log
