timeout 300 python tools/ln_one.py > gpurun_out/ln_plain.log 2>&1 && cat gpurun_out/ln_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:dropout_add_ln_bwd -s 6 -c 1 -o gpurun_out/prof_ln_bwd -f python tools/ln_one.py > gpurun_out/ncu_ln.log 2>&1; echo "ncu rc=$?"
