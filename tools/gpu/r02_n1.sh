#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_1.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r02_tests_1.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','gemm_ms_per_step','single_stream_step_ms')})
print('fbank', d['fbank'])
print('families', json.dumps(d['families'], indent=0)[:1500])
print('eager', d['gpu_eager'])
print('cpu', d['cpu_baseline'])
print('logs', d['last_step_logs'])
"
