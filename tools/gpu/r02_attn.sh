#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > gpurun_out/r02_attn_tests.log 2>&1
echo "attention tests rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02_attn_tests.log | cut -c1-300 | tail -8
timeout 200 python tools/microbench.py attn 2>&1 | tail -8
timeout 600 python -m pytest tests/test_real_shapes_gpu.py tests/test_parity_gpu.py -m gpu -q > gpurun_out/r02_model_tests.log 2>&1
echo "model tests rc=$?"; tail -3 gpurun_out/r02_model_tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n1b.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1b.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in d['families'].items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
"
