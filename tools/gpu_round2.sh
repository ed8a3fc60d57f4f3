timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s4_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/s4_bench.json; tail -3 gpurun_out/s4_bench.err
