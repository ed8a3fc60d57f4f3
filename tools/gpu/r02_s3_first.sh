#!/bin/bash
# session 3, call 1: the recipe test, the other BASELINE workloads at N = 1, the small-GEMM floor
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_real_shapes_gpu.py -m gpu -q -k "iwslt2023" > gpurun_out/s3_recipe.log 2>&1; echo "recipe rc=$?"; tail -3 gpurun_out/s3_recipe.log | cut -c1-300
timeout 400 python bench.py --workload whisper_small --steps 10 --warmup 3 > gpurun_out/s3_bench_whisper_small.json 2> gpurun_out/s3_bench_whisper_small.err; echo "whisper rc=$?"; tail -2 gpurun_out/s3_bench_whisper_small.err | cut -c1-300
timeout 400 python bench.py --workload big_adapters --steps 10 --warmup 3 > gpurun_out/s3_bench_big_adapters.json 2> gpurun_out/s3_bench_big_adapters.err; echo "big rc=$?"; tail -2 gpurun_out/s3_bench_big_adapters.err | cut -c1-300
python - <<'PY'
import json
for w in ('whisper_small', 'big_adapters'):
    try:
        d = json.loads(open(f'gpurun_out/s3_bench_{w}.json').read().strip().splitlines()[-1])
        print(w, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'gemm frac', round(d['roofline']['frac'], 4))
        print({k: round(v['ms_per_step'], 3) for k, v in d['families'].items()})
    except Exception as e:
        print(w, 'no line', e)
PY
timeout 300 python tools/mb_small.py > gpurun_out/s3_mb_small.log 2>&1; echo "mb_small rc=$?"; cat gpurun_out/s3_mb_small.log | tail -40
