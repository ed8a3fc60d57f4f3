#!/bin/bash
mkdir -p gpurun_out /tmp/rep
timeout 120 python tools/linln_one.py 16000 512 0.1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ln -s 3 -c 1 -o /tmp/rep/linln -f python tools/linln_one.py 16000 512 0.1 > gpurun_out/ncu_linln.log 2>&1; echo "ncu rc=$?"
bash tools/ncu_summary.sh /tmp/rep/linln.ncu-rep > gpurun_out/r02_ncu_linln.txt 2>&1
ncu -i /tmp/rep/linln.ncu-rep --page source --csv > /tmp/rep/linln.src.csv 2>/dev/null
python tools/ncu_hot.py /tmp/rep/linln.src.csv 60 >> gpurun_out/r02_ncu_linln.txt 2>&1
cp /tmp/rep/linln.src.csv gpurun_out/linln_src.csv
ls -la /tmp/rep
