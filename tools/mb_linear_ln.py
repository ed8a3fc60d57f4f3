"""Fused Linear + dropout + residual + LayerNorm (csrc/gemm_ln.cu) against the two kernels it replaces, at the encoder's shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasero_b200 import kernels as k  # noqa: E402
from tools.microbench import timeit  # noqa: E402

BF16 = torch.bfloat16
dev = 'cuda'
D = 512
for M, K in [(16000, 512), (16000, 2048), (8192, 512)]:
    a = torch.randn(M, K, device=dev).to(BF16)
    w = (torch.randn(D, K, device=dev) * K ** -0.5).to(BF16)
    b = torch.randn(D, device=dev).to(BF16)
    r = torch.randn(M, D, device=dev).to(BF16)
    g, be = torch.ones(D, device=dev).to(BF16), torch.zeros(D, device=dev).to(BF16)
    keep = torch.empty(M * D // 8, dtype=torch.uint8, device=dev)
    for p in (0.1, 0.0):
        kp = keep if p > 0 else None
        t_f = timeit(lambda: k.linear_dropout_add_ln_fwd(a, w, b, r, g, be, 1e-5, p, 1, 1, keep_out=kp))
        t_g = timeit(lambda: k.linear_fwd(a, w, b))
        y = k.linear_fwd(a, w, b)
        t_l = timeit(lambda: k.dropout_add_ln_fwd(y, r, g, be, 1e-5, p, 1, 1, keep_out=kp))

        def both():
            yy = k.linear_fwd(a, w, b)
            k.dropout_add_ln_fwd(yy, r, g, be, 1e-5, p, 1, 1, keep_out=kp)
        t_b = timeit(both)
        fl = 2.0 * M * D * K
        print(f'{M}x{D}x{K} p={p}: fused {t_f * 1e3:6.1f} us ({fl / t_f / 1e9:6.0f} TFLOP/s)   gemm {t_g * 1e3:6.1f} + ln {t_l * 1e3:6.1f} '
              f'= back to back {t_b * 1e3:6.1f} us')
