#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > gpurun_out/a6_kernels.log 2>&1
echo "attention tests rc=$?"; tail -3 gpurun_out/a6_kernels.log
timeout 200 python tools/microbench.py attn > gpurun_out/a6_mb.log 2>&1; cat gpurun_out/a6_mb.log | tail -8
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/a6_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/a6_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/a6_bench.json 2> gpurun_out/a6_bench.err
echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/a6_bench.json').read())
for k in ('value','ms_per_step','e2e','gemm_share_of_step','gpu_launches'):
    print(k, d[k])
print({k: d['roofline'].get(k) for k in ('achieved','frac','achieved_raw','gemm_ms_per_step','single_stream_step_ms')})
P
