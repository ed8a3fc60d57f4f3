"""Runs the fused Linear + dropout + residual + LayerNorm kernel a few times (the command ncu wraps): M K p."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasero_b200 import kernels as k
M, K, p = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
BF16 = torch.bfloat16
a = torch.randn(M, K, device='cuda').to(BF16)
w = (torch.randn(512, K, device='cuda') * K ** -0.5).to(BF16)
b = torch.randn(512, device='cuda').to(BF16)
r = torch.randn(M, 512, device='cuda').to(BF16)
g, be = torch.ones(512, device='cuda').to(BF16), torch.zeros(512, device='cuda').to(BF16)
keep = torch.empty(M * 64, dtype=torch.uint8, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(5):
    flush.zero_()
    y = k.linear_dropout_add_ln_fwd(a, w, b, r, g, be, 1e-5, p, 1, 1, keep_out=keep if p > 0 else None)[0]
torch.cuda.synchronize()
print('ok', float(y.float().abs().mean()))
