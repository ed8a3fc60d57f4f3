BENCH_GEMM_DUMP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; echo "bench rc=$?"; tail -40 gpurun_out/s6_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/s6_bench.json').read())
for k in ('value','ms_per_step','e2e','roofline','gemm_share_of_step','gpu_launches','clocks'):
    print(k, d[k])
P
