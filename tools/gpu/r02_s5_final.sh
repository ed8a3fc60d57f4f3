#!/bin/bash
# session 5, final tree at N = 1: smoke, full GPU suite, bench, LayerNorm microbench (training-shape kernel vs the general one),
# ncu --set full of the training-shape LayerNorm backward kernel
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s5_smoke.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s5_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s5_pytest_final.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/s5_bench_n1_final.json 2> gpurun_out/s5_bench_n1_final.err; echo "bench rc=$?"; tail -2 gpurun_out/s5_bench_n1_final.err | cut -c1-300
timeout 200 python tools/ln_one.py > gpurun_out/s5_ln_microbench.log 2>&1; echo "ln_one rc=$?"; cat gpurun_out/s5_ln_microbench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_fast -s 6 -c 1 -o gpurun_out/s5_prof_ln_bwd_fast -f python tools/ln_one.py > gpurun_out/s5_ncu_ln.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s5_bench_n1_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','gemm_ms_per_step','single_stream_step_ms')})
print('layernorm', d['families']['layernorm']); print('loss', d['families']['loss'])
PY
