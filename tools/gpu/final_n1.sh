set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
PB200_WGRAD_STREAM=0 timeout 300 python tools/step_breakdown.py > gpurun_out/step_breakdown.txt 2>&1; echo "breakdown rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 1400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu.log 2>&1; echo "ncu rc=$?"
