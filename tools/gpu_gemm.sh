timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > gpurun_out/gemm_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/gemm_tests.log
timeout 600 python tools/microbench.py gemm > gpurun_out/gemm_mb.log 2>&1; echo "mb rc=$?"; grep -v "^$" gpurun_out/gemm_mb.log | tail -70
