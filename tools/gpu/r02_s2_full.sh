#!/bin/bash
# full GPU suites (bfloat16 + float16 + decoding) and the N = 1 bench of the current tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s2_pytest_full.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/s2_bench_n1_b.json 2> gpurun_out/s2_bench_n1_b.err; echo "bench rc=$?"; tail -2 gpurun_out/s2_bench_n1_b.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_n1_b.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
r=d['roofline']; print('gemm frac', round(r['frac'],4), 'with fused', round(r.get('frac_with_fused_ln',0),4), 'gemm ms', round(r['gemm_ms_per_step'],3), 'single-stream', round(r['single_stream_step_ms'],3))
print({k: round(v['ms_per_step'],3) for k,v in d['families'].items()})
print('eager', d.get('gpu_eager'))
print('cpu', d.get('cpu_baseline'))
PY
