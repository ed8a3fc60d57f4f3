#!/bin/bash
# session 2 of round 2: state of the tree — smoke, full GPU suite, N=1 bench, reference arm, in-situ breakdown, ncu launch list with DRAM bytes
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s2_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s2_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/s2_bench_n1.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2_bench_ref.json 2> gpurun_out/s2_bench_ref.err; echo "ref rc=$?"
PB200_WGRAD_STREAM=0 timeout 300 python tools/step_breakdown.py > gpurun_out/s2_step_breakdown.txt 2>&1; echo "breakdown rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2800 -c 1400 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','gemm_ms_per_step','single_stream_step_ms')})
print('fbank', {k: d['fbank'].get(k) for k in ('frames_per_s','ms_per_step','frac')})
print('families', json.dumps(d.get('families'))[:1800])
PY
