timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s7_pytest.log
BENCH_GEMM_DUMP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; echo "bench rc=$?"; grep "n=" gpurun_out/s7_bench.err | head -12; grep -v "n=" gpurun_out/s7_bench.err | tail -5
python - <<'P'
import json
d=json.loads(open('gpurun_out/s7_bench.json').read())
for k in ('value','ms_per_step','e2e','roofline','gemm_share_of_step','gpu_launches'):
    print(k, d[k])
P
