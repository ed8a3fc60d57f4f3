#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/ws_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/ws_tests.log
for v in 1 0; do
PB200_WGRAD_STREAM=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ws_bench$v.json 2> gpurun_out/ws_bench$v.err
echo "bench ws=$v rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/ws_bench$v.json').read())
for k in ('value','ms_per_step','e2e','gemm_share_of_step','gpu_launches'):
    print(k, d[k])
print({k: d['roofline'].get(k) for k in ('achieved','frac','achieved_raw','gemm_ms_per_step','single_stream_step_ms')})
P
done
