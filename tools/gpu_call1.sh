set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1800 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2_ncu.log 2>&1; echo "ncu rc=$?"
