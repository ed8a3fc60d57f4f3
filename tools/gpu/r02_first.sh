#!/bin/bash
# round 2, first call: the whole -m gpu suite (old + real-shape tests) and an ncu --set full capture of the fbank kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_0.log 2>&1
echo "tests rc=$?"; tail -40 gpurun_out/r02_tests_0.log
for v in v2 v3; do
  PB200_FBANK=$v timeout 120 python tools/fbank_one.py > gpurun_out/r02_fbank_$v.log 2>&1 &&
  PB200_FBANK=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank -s 2 -c 1 \
      -o gpurun_out/r02_ncu_fbank_$v -f python tools/fbank_one.py 2 > gpurun_out/r02_ncu_fbank_$v.log 2>&1
  echo "fbank $v rc=$?"; cat gpurun_out/r02_fbank_$v.log
done
