#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/rs_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/rs_tests.log
BENCH_GEMM_DUMP=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/rs_bench2.json 2> gpurun_out/rs_bench2.err
echo "bench rc=$?"; grep "a_mn1" gpurun_out/rs_bench2.err | head -12
python - <<'P'
import json
d=json.loads(open('gpurun_out/rs_bench2.json').read())
for k in ('value','ms_per_step','e2e','gemm_share_of_step','gpu_launches'):
    print(k, d[k])
print({k: d['roofline'][k] for k in ('achieved','frac','achieved_raw','gemm_ms_per_step')})
P
