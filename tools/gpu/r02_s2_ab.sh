#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s2_pytest_b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s2_pytest_b.log | cut -c1-300
for f in 1 0; do
PB200_FUSE_LINLN=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/s2_bench_fuse$f.json 2> gpurun_out/s2_bench_fuse$f.err; echo "bench fuse=$f rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/s2_bench_fuse$f.json').read().strip().splitlines()[-1])
print('fuse=$f value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gemm frac', round(d['roofline']['frac'],4), 'gemm ms', round(d['roofline']['gemm_ms_per_step'],3), 'single-stream', round(d['roofline']['single_stream_step_ms'],3))
print({k: round(v['ms_per_step'],3) for k,v in d['families'].items()})
print(d.get('last_step_logs'))
PY
done
