for mode in tc tc_ss hmma; do
  PB200_ATTN=$mode timeout 300 python tools/microbench.py attn > gpurun_out/attn_mb_$mode.log 2>&1; echo "mb $mode rc=$?"; grep attn gpurun_out/attn_mb_$mode.log
done
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"; cat gpurun_out/s3_bench.json | cut -c1-400
