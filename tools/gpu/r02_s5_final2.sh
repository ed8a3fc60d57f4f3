#!/bin/bash
# session 5, last tree at N = 1 (every SIMT kernel of the step PDL-launched): smoke, full GPU suite, bench
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke2.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s5_smoke2.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s5_pytest_final2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s5_pytest_final2.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/s5_bench_n1_final2.json 2> gpurun_out/s5_bench_n1_final2.err; echo "bench rc=$?"; tail -2 gpurun_out/s5_bench_n1_final2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s5_bench_n1_final2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','gemm_ms_per_step','single_stream_step_ms')})
for k in ('layernorm','loss','conv_glue','embedding'): print(k, d['families'][k])
print(d['last_step_logs'])
PY
