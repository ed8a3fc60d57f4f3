#!/bin/bash
# fbank v3 (3 CTAs per SM) against v2: parity tests, then the kernel's time on the config-2 batch (64 x 10 s)
mkdir -p gpurun_out
for v in v2 v3; do
  PB200_FBANK=$v timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -m gpu -k "fbank" > gpurun_out/fbank_$v.log 2>&1
  echo "fbank $v tests rc=$?"; tail -3 gpurun_out/fbank_$v.log
  PB200_FBANK=$v timeout 120 python - <<'P'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from pasero_b200 import frontend
B, N = 64, 160000
pcm = (torch.rand(B * N, device='cuda') - 0.5)
plan = frontend.FbankPlan([N] * B, 'cuda')
for cmvn in (False, True):
    for _ in range(3):
        frontend.fbank(pcm, plan, cmvn=cmvn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        frontend.fbank(pcm, plan, cmvn=cmvn)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{os.environ['PB200_FBANK']} cmvn={cmvn}: {ms * 1e3:.1f} us, {plan.total_frames / ms / 1e6:.2f} G frames/s, "
          f"{961.3 * plan.total_frames / ms / 1e6:.0f} GB/s algorithmic")
P
done
