#!/bin/bash
# session 5: per-shape in-situ GEMM table of the step (bench.py BENCH_GEMM_DUMP=1)
mkdir -p gpurun_out
BENCH_GEMM_DUMP=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/s5_dump_bench.json 2> gpurun_out/s5_gemm_dump.txt; echo "rc=$?"
cat gpurun_out/s5_gemm_dump.txt | cut -c1-170
