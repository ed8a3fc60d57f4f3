for mode in tc tc_ss; do
  PB200_ATTN=$mode timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k attention > gpurun_out/attn_$mode.log 2>&1; echo "mode $mode rc=$?"; tail -4 gpurun_out/attn_$mode.log
  PB200_ATTN=$mode timeout 300 python tools/microbench.py attn > gpurun_out/attn_mb_$mode.log 2>&1; echo "mb $mode rc=$?"; grep attn gpurun_out/attn_mb_$mode.log
done
