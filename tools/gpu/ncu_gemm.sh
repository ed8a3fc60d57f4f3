timeout 300 python tools/gemm_one.py fwd 16000 1536 512 > gpurun_out/g_fwd_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_pair -f python tools/gemm_one.py fwd 16000 1536 512 > gpurun_out/ncu_gemm_pair.log 2>&1; echo "ncu rc=$?"
