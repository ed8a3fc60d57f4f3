"""Cost of the +residual epilogue (EPI 5) of the dgrad GEMMs: dx[M, D] = dy[M, N] @ w[N, D] (+ residual[M, D])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pasero_b200 import kernels as k
from tools.microbench import timeit

BF16 = torch.bfloat16
for M, N, D in [(16000, 1536, 512), (16000, 1024, 512), (16000, 512, 512), (3072, 1536, 512), (3072, 512, 512)]:
    dy = torch.randn(M, N, device='cuda').to(BF16)
    w = torch.randn(N, D, device='cuda').to(BF16)
    res = torch.randn(M, D, device='cuda').to(BF16)
    fl = 2.0 * M * N * D
    for name, fn in [('plain', lambda: k.linear_dgrad(dy, w)), ('residual', lambda: k.linear_dgrad(dy, w, residual=res))]:
        for flush in (True, False):
            t = timeit(fn, flush=flush)
            print(f'dgrad {M}x{D}x{N} {name:9s} flush={flush}: {t * 1e3:7.1f} us {fl / t / 1e9:7.1f} TFLOP/s', flush=True)
    x = torch.randn(M, D, device='cuda').to(BF16)
    t = timeit(lambda: torch.add(x, res), flush=True)
    print(f'torch add {M}x{D}: {t * 1e3:.1f} us')
