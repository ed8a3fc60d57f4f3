timeout 300 python tools/gemm_one.py fwd 16000 1536 512 > gpurun_out/g_fwd_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_fwd_r1b -f python tools/gemm_one.py fwd 16000 1536 512 > gpurun_out/ncu_gemm_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 python tools/gemm_one.py wgrad 16000 512 512 > gpurun_out/g_wgrad_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_wgrad_r1b -f python tools/gemm_one.py wgrad 16000 512 512 > gpurun_out/ncu_gemm_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
