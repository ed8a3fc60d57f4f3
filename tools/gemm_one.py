"""Runs one GEMM shape a few times (for ncu): mode fwd|dgrad|wgrad|wgrad32 (fp32 reduce-add sink)|wgrad32b (+ fused bias gradient)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasero_b200 import kernels as k
mode = sys.argv[1]
M, N, K = [int(x) for x in sys.argv[2:5]]
x = torch.randn(M, K, device='cuda').to(torch.bfloat16)
w = torch.randn(N, K, device='cuda').to(torch.bfloat16)
b = torch.randn(N, device='cuda').to(torch.bfloat16)
dy = torch.randn(M, N, device='cuda').to(torch.bfloat16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
g32 = torch.zeros(N, K, dtype=torch.float32, device='cuda')
b32 = torch.zeros(N, dtype=torch.float32, device='cuda')
for _ in range(6):
    flush.zero_()
    if mode == 'fwd':
        y = k.linear_fwd(x, w, b, act=1)
    elif mode == 'dgrad':
        y = k.linear_dgrad(dy, w)
    elif mode == 'wgrad32':
        y = k.linear_wgrad(dy, x, out=g32, accumulate=True)
    elif mode == 'wgrad32b':
        y = k.linear_wgrad(dy, x, out=g32, accumulate=True, bias_grad32=b32)
    else:
        y = k.linear_wgrad(dy, x)
torch.cuda.synchronize()
print('ok', float(y.float().abs().mean()))
