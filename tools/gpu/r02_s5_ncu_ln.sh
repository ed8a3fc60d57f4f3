#!/bin/bash
# session 5: ncu --set full of the training-shape LayerNorm backward kernel (after the same command ran clean without ncu)
mkdir -p gpurun_out /tmp/rep
timeout 200 python tools/ln_one.py > gpurun_out/s5_ln_plain.log 2>&1; echo "ln plain rc=$?"; cat gpurun_out/s5_ln_plain.log
N="ncu --set full --clock-control none --import-source on"
name=ln_bwd_fast
timeout 400 $N -k regex:ln_bwd_fast -s 6 -c 1 -o /tmp/rep/$name -f python tools/ln_one.py > gpurun_out/ncu_$name.log 2>&1; echo "ncu rc=$?"
bash tools/ncu_summary.sh /tmp/rep/$name.ncu-rep > gpurun_out/r02_ncu_$name.txt 2>&1
ncu -i /tmp/rep/$name.ncu-rep --page source --csv > /tmp/rep/$name.src.csv 2>/dev/null; echo '# top SASS lines by stall samples' >> gpurun_out/r02_ncu_$name.txt; python tools/ncu_hot.py /tmp/rep/$name.src.csv 40 >> gpurun_out/r02_ncu_$name.txt 2>&1
ncu -i /tmp/rep/$name.ncu-rep --page details 2>/dev/null | grep -iE "stall|warp cycles per|eligible|issued warp|active warps" | head -30 >> gpurun_out/r02_ncu_$name.txt
cp /tmp/rep/$name.ncu-rep gpurun_out/ 2>/dev/null
head -60 gpurun_out/r02_ncu_$name.txt
