"""Runs one GEMM shape a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasero_b200 import kernels as k
M, N, K = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (16000, 2048, 512))]
tn = int(sys.argv[4]) if len(sys.argv) > 4 else 256
x = torch.randn(M, K, device='cuda').to(torch.bfloat16)
w = torch.randn(N, K, device='cuda').to(torch.bfloat16)
b = torch.randn(N, device='cuda').to(torch.bfloat16)
for _ in range(6):
    y = k.linear_fwd(x, w, b, act=1, tile_n=tn)
torch.cuda.synchronize()
print('ok', float(y.float().abs().mean()))
