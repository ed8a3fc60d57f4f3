#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/p_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/p_tests.log
for v in 1 0; do
BENCH_PIPELINE=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/p_bench$v.json 2> gpurun_out/p_bench$v.err
echo "bench pipeline=$v rc=$?"; tail -3 gpurun_out/p_bench$v.err
python - <<P
import json
d=json.loads(open('gpurun_out/p_bench$v.json').read())
for k in ('value','ms_per_step','e2e','gemm_share_of_step','gpu_launches'):
    print(k, d[k])
print({k: d['roofline'].get(k) for k in ('achieved','frac','achieved_raw','gemm_ms_per_step','single_stream_step_ms')})
P
done
