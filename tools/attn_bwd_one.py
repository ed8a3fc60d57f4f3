"""One attention backward launch at the encoder's shape (64 x 8 heads x 250 x 250); with a -DPB_ATTN_TIMING build the kernel
prints CTA 0's per-iteration clock64() timeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasero_b200 import kernels as k
B, H, Lq, Lk = 64, 8, 250, 250
q = torch.randn(B * Lq, H * 64, device='cuda').to(torch.bfloat16)
kv = torch.randn(B * Lk, 2 * H * 64, device='cuda').to(torch.bfloat16)
kl = torch.full((B,), Lk, device='cuda', dtype=torch.int32)
o, lse = k.attention_fwd(q, kv[:, :H * 64], kv[:, H * 64:], B, H, Lq, Lk, kl, False, 0.125)
do = torch.randn_like(o)
dq, dkv = torch.empty_like(q), torch.empty_like(kv)
for _ in range(2):
    k.attention_bwd(do, q, kv[:, :H * 64], kv[:, H * 64:], o, lse, B, H, Lq, Lk, kl, False, 0.125, dq, dkv[:, :H * 64], dkv[:, H * 64:])
    torch.cuda.synchronize()
    print('---')
