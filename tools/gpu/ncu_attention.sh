timeout 300 python tools/microbench.py attn > gpurun_out/ncu_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -o gpurun_out/prof_attn_fwd -f python tools/microbench.py attn > gpurun_out/ncu_attn_fwd.log 2>&1; echo "ncu fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 2 -c 1 -o gpurun_out/prof_attn_bwd -f python tools/microbench.py attn > gpurun_out/ncu_attn_bwd.log 2>&1; echo "ncu bwd rc=$?"
