import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasero_b200 import kernels as k
from tools.microbench import timeit
BF16 = torch.bfloat16
dev = 'cuda'
M, N, K = 16000, 512, 2048     # fc2: dy [M, N], w2 [N, K]; dh [M, K]
dy = torch.randn(M, N, device=dev).to(BF16)
w2 = torch.randn(N, K, device=dev).to(BF16)
h = torch.randn(M, K, device=dev).to(BF16)
bits = torch.randint(-2**31, 2**31 - 1, (K // 32, M), device=dev, dtype=torch.int32)
fl = 2.0 * M * N * K
for name, fn in [('plain', lambda: k.linear_dgrad(dy, w2)), ('relu_mask', lambda: k.linear_dgrad(dy, w2, relu_mask=h)),
                 ('relu_bits', lambda: k.linear_dgrad(dy, w2, relu_bits=bits)), ('residual', lambda: k.linear_dgrad(dy, w2, residual=h))]:
    for flush in (True, False):
        t = timeit(fn, flush=flush)
        print(f'fc2 dgrad {name:10s} flush={flush}: {t * 1e3:7.1f} us {fl / t / 1e9:7.1f} TFLOP/s')
x = torch.randn(M, N, device=dev).to(BF16)
w1 = torch.randn(K, N, device=dev).to(BF16)
b1 = torch.randn(K, device=dev).to(BF16)
ob = torch.empty(K // 32, M, dtype=torch.int32, device=dev)
for name, fn in [('bias+relu', lambda: k.linear_fwd(x, w1, b1, act=1)), ('bias+relu+bits', lambda: k.linear_fwd(x, w1, b1, act=1, relu_bits_out=ob)),
                 ('bias', lambda: k.linear_fwd(x, w1, b1))]:
    for flush in (True, False):
        t = timeit(fn, flush=flush)
        print(f'fc1 fwd {name:15s} flush={flush}: {t * 1e3:7.1f} us {fl / t / 1e9:7.1f} TFLOP/s')
