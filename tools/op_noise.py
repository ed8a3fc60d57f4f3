"""Per-op bf16 noise at BASELINE shapes: each hand-written kernel and the stock PyTorch bf16 op are both compared with an
fp32 evaluation on the same bf16 inputs. Prints max-abs error / max |truth| for ours and for PyTorch. (Diagnostic behind
the calibrated tolerances of tests/test_real_shapes_gpu.py.)"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasero_b200 import kernels as K  # noqa: E402

BF16 = torch.bfloat16
torch.manual_seed(0)


def err(a, t):
    return ((a.float() - t.float()).abs().max() / t.float().abs().max().clamp_min(1e-20)).item()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device='cuda') * scale).to(BF16)


def attention(B, H, Lq, Lk, causal, masked, qscale=1.0):
    D = H * 64
    q2, kv2 = rnd(B * Lq, D, scale=qscale), rnd(B * Lk, 2 * D, scale=qscale)
    k2, v2 = kv2[:, :D], kv2[:, D:]
    kv_len = None
    if masked:
        kv_len = torch.randint(max(1, Lk // 2), Lk + 1, (B,), device='cuda', dtype=torch.int32)
        kv_len[0] = Lk
    scale = 1 / 8
    do = rnd(B * Lq, D)

    def ref(dtype):
        def heads(t, L):
            return t.to(dtype).reshape(B, L, H, 64).transpose(1, 2).detach().requires_grad_(True)
        q, k, v = heads(q2, Lq), heads(k2, Lk), heads(v2, Lk)
        s = torch.matmul(q, k.transpose(2, 3)).float() * scale
        if causal:
            s = s.masked_fill(torch.ones(Lq, Lk, dtype=torch.bool, device='cuda').triu(1 + Lk - Lq), float('-inf'))
        elif kv_len is not None:
            s = s.masked_fill((torch.arange(Lk, device='cuda')[None] >= kv_len[:, None])[:, None, None], float('-inf'))
        p = torch.softmax(s, -1).to(dtype)
        o = torch.matmul(p, v)
        o.backward(do.to(dtype).view(B, Lq, H, 64).transpose(1, 2))
        return o, q.grad, k.grad, v.grad
    t = ref(torch.float32)
    r = ref(BF16)
    o, lse = K.attention_fwd(q2, k2, v2, B, H, Lq, Lk, kv_len, causal, scale)
    dq, dkv = torch.empty_like(q2), torch.empty_like(kv2)
    K.attention_bwd(do, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale, dq, dkv[:, :D], dkv[:, D:])
    ours = (o.view(B, Lq, H, 64).transpose(1, 2), dq.view(B, Lq, H, 64).transpose(1, 2),
            dkv[:, :D].reshape(B, Lk, H, 64).transpose(1, 2), dkv[:, D:].reshape(B, Lk, H, 64).transpose(1, 2))
    names = ('o', 'dq', 'dk', 'dv')
    print(f'attention B={B} H={H} Lq={Lq} Lk={Lk} causal={causal} masked={masked} |q|={qscale}: ' +
          '  '.join(f'{n}: ours {err(a, tt):.4f} torch {err(b, tt):.4f}' for n, a, b, tt in zip(names, ours, r, t)))


def linear(M, N, Kd):
    x, w, b, dy = rnd(M, Kd), rnd(N, Kd, scale=Kd ** -0.5), rnd(N, scale=0.1), rnd(M, N)
    t = (F.linear(x.float(), w.float(), b.float()), dy.float() @ w.float(), dy.float().t() @ x.float())
    r = (F.linear(x, w, b), dy @ w, dy.t() @ x)
    ours = (K.linear_fwd(x, w, b), K.linear_dgrad(dy, w), K.linear_wgrad(dy, x))
    print(f'linear {M}x{N}x{Kd}: ' + '  '.join(f'{n}: ours {err(a, tt):.4f} torch {err(bb, tt):.4f}'
                                              for n, a, bb, tt in zip(('y', 'dx', 'dw'), ours, r, t)))


def layernorm(M, D):
    x, res, dy = rnd(M, D), rnd(M, D), rnd(M, D)
    g, b = (1 + 0.1 * torch.randn(D, device='cuda')).to(BF16), rnd(D, scale=0.1)

    def ref(dtype):
        xs = (x.to(dtype) + res.to(dtype)).detach().requires_grad_(True)
        gg, bb = g.to(dtype).requires_grad_(True), b.to(dtype).requires_grad_(True)
        y = F.layer_norm(xs, (D,), gg, bb, 1e-5)
        y.backward(dy.to(dtype))
        return y, xs.grad, gg.grad, bb.grad
    t, r = ref(torch.float32), ref(BF16)
    y, s, mean, rstd = K.dropout_add_ln_fwd(x, res, g, b, 1e-5, 0.0, 0, 0)
    dg, db = torch.empty(D, dtype=BF16, device='cuda'), torch.empty(D, dtype=BF16, device='cuda')
    dx, _ = K.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, 0.0, 0, 0, True, dg, db)
    print(f'layernorm {M}x{D}: ' + '  '.join(f'{n}: ours {err(a, tt):.4f} torch {err(bb, tt):.4f}'
                                            for n, a, bb, tt in zip(('y', 'dx', 'dgamma', 'dbeta'), (y, dx, dg, db), r, t)))


if __name__ == '__main__':
    for args in [(4, 8, 250, 250, False, True), (4, 8, 275, 275, False, True), (4, 8, 48, 48, True, False), (4, 8, 48, 250, False, True),
                 (2, 12, 1500, 1500, False, True), (2, 16, 32, 40, False, True), (4, 8, 250, 250, False, True, 3.0)]:
        attention(*args)
    for args in [(1000, 1536, 512), (1000, 2048, 512), (1000, 512, 2048), (192, 8032, 512)]:
        linear(*args)
    layernorm(1000, 512)
    layernorm(3000, 768)
