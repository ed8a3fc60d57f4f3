#!/bin/bash
# session 4, state of the final tree at N = 1: smoke, full GPU suite, bench, in-situ breakdown, ncu launch list with DRAM bytes
# (-> traffic per family), memcheck of the SIMT kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s4_smoke.log | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s4_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/s4_bench_n1_final.json 2> gpurun_out/s4_bench_n1_final.err; echo "bench rc=$?"; tail -2 gpurun_out/s4_bench_n1_final.err | cut -c1-300
PB200_WGRAD_STREAM=0 timeout 300 python tools/step_breakdown.py > gpurun_out/s4_step_breakdown.txt 2>&1; echo "breakdown rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/s4_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2800 -c 1400 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/s4_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_traffic.py gpurun_out/s4_launches.csv gemm:254 gpurun_out/s4_traffic.json > gpurun_out/s4_traffic.log 2>&1; echo "traffic rc=$?"; tail -12 gpurun_out/s4_traffic.log | cut -c1-200
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "dropout_add_ln or colsum or embed or glu or gelu or ls_ce or adam or fbank_vs_oracle or collate" > gpurun_out/s4_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/s4_memcheck.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4_bench_n1_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','gemm_ms_per_step','single_stream_step_ms','frac_from_ncu_launch_list')})
print('fbank', {k: d['fbank'].get(k) for k in ('frames_per_s','ms_per_step','frac')})
PY
