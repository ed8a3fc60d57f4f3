"""The fbank(+CMVN) kernel alone on the config-2 batch (64 x 10 s): a few launches with their device time. Short on
purpose: this is the command `ncu --set full` wraps (tools/gpu/ncu_fbank.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasero_b200 import frontend  # noqa: E402

B, N = 64, 160000
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pcm = (torch.rand(B * N, device='cuda') - 0.5)
plan = frontend.FbankPlan([N] * B, 'cuda')
rows, front = plan.max_frames + 6, 2
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(2):
    frontend.fbank(pcm, plan, cmvn=True, padded_frames=rows, front_pad=front)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    frontend.fbank(pcm, plan, cmvn=True, padded_frames=rows, front_pad=front)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
nbytes = 4 * B * N + 320 * plan.total_frames
print(f"fbank+cmvn: {ms * 1e3:.1f} us (eager launch pair, L2 flushed), "
      f"{plan.total_frames / ms / 1e6:.3f} G frames/s, {nbytes / ms / 1e6:.0f} GB/s algorithmic")
