#!/bin/bash
# session 5: in-situ breakdown of the last tree
mkdir -p gpurun_out
PB200_WGRAD_STREAM=0 timeout 200 python tools/step_breakdown.py > gpurun_out/s5_step_breakdown_last.txt 2>&1; echo "rc=$?"; head -26 gpurun_out/s5_step_breakdown_last.txt
