timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k attention > gpurun_out/attn_tc.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/attn_tc.log
for v in 1 0; do echo "PB200_ATTN_BWD_TMA=$v"; PB200_ATTN_BWD_TMA=$v timeout 300 python tools/microbench.py attn 2>&1 | grep "attn bwd"; done
