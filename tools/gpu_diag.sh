timeout 600 python tools/e2e_diag.py > gpurun_out/e2e_diag.log 2>&1; echo "diag rc=$?"; tail -12 gpurun_out/e2e_diag.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s5_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1600 --csv --log-file gpurun_out/s5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s5_ncu.log 2>&1; echo "ncu rc=$?"
