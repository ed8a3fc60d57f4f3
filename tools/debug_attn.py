"""Largest attention-backward errors against the fp32 torch reference for one shape: B H Lq Lk causal(0/1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tests.test_kernels_gpu as tk
B, H, Lq, Lk, causal = [int(x) for x in sys.argv[1:6]]
k = tk._k()
D = H * 64
q2, kv2 = tk._rand(B * Lq, D, seed=1), tk._rand(B * Lk, 2 * D, seed=2)
k2, v2 = kv2[:, :D], kv2[:, D:]
o, lse = k.attention_fwd(q2, k2, v2, B, H, Lq, Lk, None, bool(causal), 0.125)
heads = lambda t, L: t.float().reshape(B, L, H, 64).transpose(1, 2).detach().requires_grad_(True)
qr, kr, vr = heads(q2, Lq), heads(k2, Lk), heads(v2, Lk)
ref = tk._attn_ref(qr, kr, vr, None, bool(causal), 0.125)
do = tk._rand(B * Lq, D, seed=3)
ref.backward(do.float().view(B, Lq, H, 64).transpose(1, 2))
dq, dkv = torch.empty_like(q2), torch.empty_like(kv2)
k.attention_bwd(do, q2, k2, v2, o, lse, B, H, Lq, Lk, None, bool(causal), 0.125, dq, dkv[:, :D], dkv[:, D:])
for name, got, want, L in (('dq', dq, qr.grad, Lq), ('dk', dkv[:, :D], kr.grad, Lk), ('dv', dkv[:, D:], vr.grad, Lk)):
    g = got.float().reshape(B, L, H, 64).transpose(1, 2)
    err = (g - want).abs()
    top = err.flatten().topk(6)
    print(name, 'max err', err.max().item(), 'mean err', err.mean().item())
    for v, idx in zip(top.values.tolist(), top.indices.tolist()):
        b, h, r, d = idx // (H * L * 64), (idx // (L * 64)) % H, (idx // 64) % L, idx % 64
        print(f'   err {v:.4f} at b={b} h={h} row={r} d={d}: got {g[b, h, r, d].item():.4f} want {want[b, h, r, d].item():.4f}')
