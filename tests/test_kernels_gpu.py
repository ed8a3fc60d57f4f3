"""Kernel-level numerics: every C-ABI compute entry point against a plain PyTorch fp32 restatement of the
same op on the same seeded inputs (the model-level parity against the oracle lives in test_parity_gpu.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPE = torch.bfloat16      # tests/test_fp16_gpu.py re-runs this module's kernel tests with torch.float16 (the library's other build)


def _k():
    from pasero_b200 import kernels
    return kernels


def _rand(*shape, scale=1.0, seed=0, dtype=None):
    dtype = dtype or DTYPE
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(*shape, generator=g, device='cuda') * scale).to(dtype)


def _close(got, want, atol, rtol, name=''):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = (err > tol)
    assert not bad.any(), f'{name}: {int(bad.sum())}/{bad.numel()} off, max err {err.max().item():.4g} ' \
                          f'(ref max {want.abs().max().item():.4g})'


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (300, 136, 200), (1000, 512, 512), (4096, 2048, 512), (257, 8032, 512)])
@pytest.mark.parametrize('tile_n', [128, 256])
@pytest.mark.parametrize('cluster_m', [1, 2, 4])
def test_gemm_forward(M, N, K, tile_n, cluster_m):
    k = _k()
    x, w, b = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, seed=3)
    y = k.linear_fwd(x, w, b, tile_n=tile_n, cluster_m=cluster_m)
    ref = x.float() @ w.float().T + b.float()
    _close(y, ref, 2e-2, 1e-2, 'linear_fwd')


@pytest.mark.parametrize('M,N,K', [(1000, 512, 512), (4096, 1024, 256), (300, 136, 200), (16000, 512, 2048)])
def test_gemm_tile512(M, N, K):
    """256 x 512 tiles on CTA pairs (single accumulator, two N = 256 MMAs per k-step): forced, and with each epilogue."""
    k = _k()
    x, w, b, r = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, seed=3), _rand(M, N, seed=4)
    ref = x.float() @ w.float().T
    _close(k.linear_fwd(x, w, b, tile_n=512), ref + b.float(), 2e-2, 1e-2, 'tile512 bias')
    _close(k.linear_fwd(x, w, None, tile_n=512), ref, 2e-2, 1e-2, 'tile512 plain')
    wt = w.T.contiguous()    # [K, N]: the MN-major B operand of a dgrad
    y = torch.empty(M, N, dtype=DTYPE, device='cuda')
    k.gemm_raw(x.data_ptr(), wt.data_ptr(), y, M, N, K, lda=K, ldb=N, ldc=N, b_mn_major=True, residual=r, split_k=1, tile_n=512)
    _close(y, ref.to(DTYPE).float() + r.float(), 3e-2, 1e-2, 'tile512 MN-major B + residual')


def test_gemm_epilogues():
    k = _k()
    M, N, K = 513, 384, 256
    x, w, b, r = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, seed=3), _rand(M, N, seed=4)
    ref = x.float() @ w.float().T + b.float()
    _close(k.linear_fwd(x, w, b, act=1), ref.relu(), 2e-2, 1e-2, 'relu')
    _close(k.linear_fwd(x, w, b, act=2), F.gelu(ref), 2e-2, 1e-2, 'gelu')
    _close(k.linear_fwd(x, w, b, residual=r), ref + r.float(), 2e-2, 1e-2, 'residual')
    _close(k.linear_fwd(x, w, None, out_dtype=torch.float32), x.float() @ w.float().T, 1e-3, 1e-3, 'fp32 out')


@pytest.mark.parametrize('M,N,K', [(300, 136, 200), (2048, 512, 2048), (16000, 512, 512)])
def test_gemm_dgrad_wgrad(M, N, K):
    k = _k()
    x, w, dy = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(M, N, seed=5)
    dx = k.linear_dgrad(dy, w)
    _close(dx, dy.float() @ w.float(), 3e-2, 1e-2, 'dgrad')
    h = _rand(M, K, seed=6)
    dxm = k.linear_dgrad(dy, w, relu_mask=h)
    _close(dxm, (dy.float() @ w.float()) * (h.float() > 0), 3e-2, 1e-2, 'dgrad+relu mask')
    if K % 32 == 0:                              # 1-bit ReLU sign map written by the forward epilogue, applied by dgrad
        b1 = _rand(K, seed=8)
        bits = torch.empty(K // 32, M, dtype=torch.int32, device='cuda')
        hh = k.linear_fwd(dy, w.T.contiguous(), b1, act=1, relu_bits_out=bits)
        unpacked = ((bits.T.unsqueeze(-1) >> torch.arange(32, device='cuda', dtype=torch.int32)) & 1).reshape(M, K).bool()
        assert torch.equal(unpacked, hh > 0)
        dxb = k.linear_dgrad(dy, w, relu_bits=bits)
        _close(dxb, (dy.float() @ w.float()) * (hh.float() > 0), 3e-2, 1e-2, 'dgrad+relu bits')
    dxr = k.linear_dgrad(dy, w, residual=h)     # skip-connection gradient added in the TMA-store epilogue
    _close(dxr, (dy.float() @ w.float()).to(DTYPE).float() + h.float(), 3e-2, 1e-2, 'dgrad+residual')
    dw = k.linear_wgrad(dy, x)
    ref = dy.float().T @ x.float()
    _close(dw, ref, 2e-2 * math.sqrt(M / 256), 1e-2, 'wgrad')
    dw2 = k.linear_wgrad(dy, x, out=dw.clone(), accumulate=True)
    _close(dw2, 2 * ref, 4e-2 * math.sqrt(M / 256), 2e-2, 'wgrad accumulate')


@pytest.mark.parametrize('M,N,K', [(16000, 1536, 512), (16000, 2048, 512), (3072, 512, 512), (3072, 1024, 512),
                                   (1000, 200, 136), (16000, 512, 2048), (77, 64, 64)])
def test_wgrad_fused_bias_grad(M, N, K):
    """pb200_gemm_params.a_rowsum: the bias gradient (column sums of dy) out of the weight-gradient GEMM itself — every
    tile shape / CTA-pair / split-K variant the training step reaches, accumulating into non-zero fp32 buffers"""
    k = _k()
    dy, x = _rand(M, N, seed=1), _rand(M, K, seed=2)
    w32 = torch.full((N, K), 0.5, dtype=torch.float32, device='cuda')
    b32 = torch.full((N,), -2.0, dtype=torch.float32, device='cuda')
    for rep in range(2):   # second round: accumulate on top
        k.linear_wgrad(dy, x, out=w32, accumulate=True, bias_grad32=b32)
    ref_w = 0.5 + 2 * (dy.float().T @ x.float())
    ref_b = -2.0 + 2 * dy.float().sum(0)
    _close(w32, ref_w, 1e-2 * math.sqrt(M / 256), 1e-3, 'wgrad fp32')
    _close(b32, ref_b, 1e-3 * math.sqrt(M), 1e-4, 'fused bias grad')
    # the stand-alone fp32 column-sum kernel (fallback for shapes / paths the GEMM cannot fold it into)
    c32 = torch.full((N,), -2.0, dtype=torch.float32, device='cuda')
    k.colsum(dy, out=c32, accumulate=True)
    _close(c32, -2.0 + dy.float().sum(0), 1e-3 * math.sqrt(M), 1e-4, 'colsum fp32')


def test_gemm_conv_view():
    """strided conv1d (k=5, s=2, p=2) as a GEMM over an overlapping-row view of the zero-padded input"""
    k = _k()
    B, T, C, N, KW = 3, 101, 80, 256, 5
    x = _rand(B, T, C, seed=1)
    w = _rand(N, C, KW, scale=(C * KW) ** -0.5, seed=2)
    bias = _rand(N, seed=3)
    Tout = 1 + (T + 2 * (KW // 2) - KW) // 2
    Tp = T + 4 + (T % 2)
    xp = k.pad_rows(x, B, T, C, Tp, 2)
    w2 = w.permute(0, 2, 1).contiguous().view(N, KW * C)      # [N, tap*C + c]
    y = torch.empty(B, Tout, N, dtype=DTYPE, device='cuda')
    k.gemm_raw(xp.data_ptr(), w2.data_ptr(), y, Tout, N, KW * C, lda=2 * C, ldb=KW * C, ldc=N, batch=B,
               a_batch_stride=Tp * C, b_batch_stride=0, c_batch_stride=Tout * N, bias=bias, split_k=1)
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), bias.float(), stride=2, padding=2).transpose(1, 2)
    _close(y, ref, 3e-2, 1e-2, 'conv fwd')
    # weight gradient: reduce over (batch, time) with both operands MN-major
    dy = _rand(B, Tout, N, seed=4)
    dw = torch.empty(N, KW * C, dtype=DTYPE, device='cuda')
    k.gemm_raw(dy.data_ptr(), xp.data_ptr(), dw, N, KW * C, Tout, lda=N, ldb=2 * C, ldc=KW * C, batch=B,
               a_batch_stride=Tout * N, b_batch_stride=Tp * C, a_mn_major=True, b_mn_major=True, reduce_batch=True)
    xr = x.float().transpose(1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    F.conv1d(xr, wr, None, stride=2, padding=2).backward(dy.float().transpose(1, 2))
    _close(dw.view(N, KW, C).permute(0, 2, 1), wr.grad, 5e-2, 2e-2, 'conv wgrad')
    # input gradient: dcol = dy @ w2, then col2im
    dcol = torch.empty(B * Tout, KW * C, dtype=DTYPE, device='cuda')
    k.gemm_raw(dy.data_ptr(), w2.data_ptr(), dcol, B * Tout, KW * C, N, lda=N, ldb=KW * C, ldc=KW * C, b_mn_major=True,
               split_k=1)
    dx = k.col2im(dcol, B, T, Tout, C, KW, 2)
    _close(dx.view(B, T, C), xr.grad.transpose(1, 2), 5e-2, 2e-2, 'conv dgrad')


def test_gemm_shared_operand_at_the_end_of_a_mapping():
    """Regression: the weight of a batched implicit-GEMM convolution is shared by all batches (batch stride 0). Its TMA
    descriptor once claimed `batch` slices of the weight's size, i.e. memory past the end of the weight, and the launch died
    with an illegal memory access whenever that range was not mapped (found by running the test files in another order: the
    Whisper-style conv front end after four Trainers had come and gone). The weight here fills a 10 MiB device allocation
    of its own (the caching allocator gives requests above 10 MiB their own segment), after everything cached was released,
    so what follows it in the address space is most likely unmapped."""
    import gc
    k = _k()
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    B, T, C, N, KW = 4, 93, 128, 8192, 5          # weight: 8192 x 640 x 2 B = 10 MiB exactly
    w2 = _rand(N, KW * C, scale=(C * KW) ** -0.5, seed=2)
    assert w2.numel() * 2 == 10 * 2 ** 20
    x = _rand(B, T, C, seed=1)
    Tout = 1 + (T + 2 * (KW // 2) - KW) // 2
    Tp = T + 4 + (T % 2)
    xp = k.pad_rows(x, B, T, C, Tp, 2)
    y = torch.empty(B, Tout, N, dtype=DTYPE, device='cuda')
    k.gemm_raw(xp.data_ptr(), w2.data_ptr(), y, Tout, N, KW * C, lda=2 * C, ldb=KW * C, ldc=N, batch=B,
               a_batch_stride=Tp * C, b_batch_stride=0, c_batch_stride=Tout * N, split_k=1)
    torch.cuda.synchronize()
    w = w2.view(N, KW, C).permute(0, 2, 1).float()
    ref = F.conv1d(x.float().transpose(1, 2), w, None, stride=2, padding=2).transpose(1, 2)
    _close(y, ref, 3e-2, 1e-2, 'conv fwd, weight in its own mapping')


# ---------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, kv_len, causal, scale):
    B, H, Lq, Lk = q.shape[0], q.shape[1], q.shape[2], k.shape[2]
    s = (q @ k.transpose(-1, -2)) * scale
    if kv_len is not None:
        mask = torch.arange(Lk, device=q.device)[None, :] >= kv_len[:, None]
        s = s.masked_fill(mask[:, None, None, :], float('-inf'))
    if causal:
        cm = torch.ones(Lq, Lk, dtype=torch.bool, device=q.device).triu(1 + Lk - Lq)
        s = s.masked_fill(cm, float('-inf'))
    return torch.softmax(s, -1) @ v


@pytest.mark.parametrize('B,H,Lq,Lk,causal,masked', [
    (2, 2, 64, 64, False, False), (3, 8, 250, 250, False, True), (3, 8, 48, 250, False, True),
    (4, 8, 48, 48, True, False), (2, 4, 130, 130, True, False), (2, 2, 1, 77, False, True),
    (2, 2, 200, 256, False, True), (2, 3, 256, 256, True, False), (2, 2, 17, 33, False, True),
    (1, 2, 300, 200, False, True),      # two query blocks in backward
    (2, 2, 100, 300, False, True), (1, 2, 320, 320, True, False),   # Lk > 256: key blocks merged by attn_combine / attn_reduce
    (2, 8, 750, 750, False, True), (1, 12, 1500, 1500, False, True), (2, 4, 48, 750, False, True),   # 30 s of audio, Whisper's 1 500
    (1, 2, 448, 448, True, False), (1, 2, 257, 513, True, False)])
def test_attention(B, H, Lq, Lk, causal, masked):
    k = _k()
    D = H * 64
    q2, kv2 = _rand(B * Lq, D, seed=1), _rand(B * Lk, 2 * D, seed=2)
    k2, v2 = kv2[:, :D], kv2[:, D:]
    gen = torch.Generator(device='cuda').manual_seed(17 + Lq + Lk)   # seeded: the key lengths are part of the test case
    kv_len = torch.randint(1, Lk + 1, (B,), device='cuda', dtype=torch.int32, generator=gen) if masked else None
    if masked:
        kv_len[0] = Lk
    scale = 1 / 8
    o, lse = k.attention_fwd(q2, k2, v2, B, H, Lq, Lk, kv_len, causal, scale)

    def heads(t, L):
        return t.float().reshape(B, L, H, 64).transpose(1, 2).detach().requires_grad_(True)
    qr, kr, vr = heads(q2, Lq), heads(k2, Lk), heads(v2, Lk)
    ref = _attn_ref(qr, kr, vr, kv_len, causal, scale)
    _close(o.view(B, Lq, H, 64).transpose(1, 2), ref, 2e-2, 2e-2, 'attn fwd')
    do = _rand(B * Lq, D, seed=3)
    ref.backward(do.float().view(B, Lq, H, 64).transpose(1, 2))
    dq = torch.empty_like(q2)
    dkv = torch.empty_like(kv2)
    k.attention_bwd(do, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale, dq, dkv[:, :D], dkv[:, D:])
    _close(dq.view(B, Lq, H, 64).transpose(1, 2), qr.grad, 3e-2, 3e-2, 'dq')
    _close(dkv[:, :D].reshape(B, Lk, H, 64).transpose(1, 2), kr.grad, 3e-2, 3e-2, 'dk')
    _close(dkv[:, D:].reshape(B, Lk, H, 64).transpose(1, 2), vr.grad, 3e-2, 3e-2, 'dv')


@pytest.mark.parametrize('B,H,Lq,Lk,causal,masked,p', [
    (3, 8, 250, 250, False, True, 0.1), (4, 8, 48, 48, True, False, 0.1), (3, 8, 48, 250, False, True, 0.3),
    (2, 4, 300, 750, False, True, 0.1), (1, 2, 320, 320, True, False, 0.25)])
def test_attention_dropout(B, H, Lq, Lk, causal, masked, p):
    """Attention dropout (pasero/models/modules.py:686-731), mask-consistency style: the keep bits the forward kernel drew are
    read back and the reference is evaluated with exactly that mask, O = (softmax(S) o M / (1 - p)) V, forward and backward.
    The keep rate, and that a second call with another offset draws another mask, are checked as well."""
    k = _k()
    D = H * 64
    q2, kv2 = _rand(B * Lq, D, seed=1), _rand(B * Lk, 2 * D, seed=2)
    k2, v2 = kv2[:, :D], kv2[:, D:]
    gen = torch.Generator(device='cuda').manual_seed(23 + Lq + Lk)
    kv_len = torch.randint(1, Lk + 1, (B,), device='cuda', dtype=torch.int32, generator=gen) if masked else None
    if masked:
        kv_len[0] = Lk
    scale = 1 / 8
    mask = k.attention_mask_buffer(B, H, Lq, Lk)
    mask.fill_(-1)
    drop = (p, 1234, 7, mask)
    o, lse = k.attention_fwd(q2, k2, v2, B, H, Lq, Lk, kv_len, causal, scale, dropout=drop)
    bits = ((mask.unsqueeze(-1) >> torch.arange(32, device='cuda', dtype=torch.int32)) & 1).reshape(B, H, Lq, -1)[..., :Lk].bool()
    mask2 = k.attention_mask_buffer(B, H, Lq, Lk)
    mask2.fill_(-1)
    k.attention_fwd(q2, k2, v2, B, H, Lq, Lk, kv_len, causal, scale, dropout=(p, 1234, 8, mask2))
    assert not torch.equal(mask, mask2)

    def heads(t, L):
        return t.float().reshape(B, L, H, 64).transpose(1, 2).detach().requires_grad_(True)
    qr, kr, vr = heads(q2, Lq), heads(k2, Lk), heads(v2, Lk)
    s = (qr @ kr.transpose(-1, -2)) * scale
    if kv_len is not None:
        s = s.masked_fill((torch.arange(Lk, device='cuda')[None, :] >= kv_len[:, None])[:, None, None, :], float('-inf'))
    if causal:
        s = s.masked_fill(torch.ones(Lq, Lk, dtype=torch.bool, device='cuda').triu(1 + Lk - Lq), float('-inf'))
    visible = torch.isfinite(s)      # (words of fully masked 32-key chunks are never drawn)
    rate = bits[visible].float().mean().item()
    assert abs(rate - (1 - p)) < 0.01, rate
    ref = (torch.softmax(s, -1) * bits / (1 - p)) @ vr
    _close(o.view(B, Lq, H, 64).transpose(1, 2), ref, 2e-2, 2e-2, 'attn dropout fwd')
    do = _rand(B * Lq, D, seed=3)
    ref.backward(do.float().view(B, Lq, H, 64).transpose(1, 2))
    dq, dkv = torch.empty_like(q2), torch.empty_like(kv2)
    k.attention_bwd(do, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale, dq, dkv[:, :D], dkv[:, D:], dropout=drop)
    _close(dq.view(B, Lq, H, 64).transpose(1, 2), qr.grad, 3e-2, 3e-2, 'dropout dq')
    _close(dkv[:, :D].reshape(B, Lk, H, 64).transpose(1, 2), kr.grad, 3e-2, 3e-2, 'dropout dk')
    _close(dkv[:, D:].reshape(B, Lk, H, 64).transpose(1, 2), vr.grad, 3e-2, 3e-2, 'dropout dv')


# ---------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize('M,D', [(1000, 512), (77, 128), (300, 768), (129, 1024)])
def test_dropout_add_ln(M, D):
    k = _k()
    x, r = _rand(M, D, seed=1), _rand(M, D, seed=2)
    g, b = (1 + 0.1 * _rand(D, seed=3).float()).to(DTYPE), _rand(D, scale=0.1, seed=4)
    y, s, mean, rstd = k.dropout_add_ln_fwd(x, r, g, b, 1e-5, 0.0, 0, 0)
    sr = (x.float() + r.float()).to(DTYPE)
    assert torch.equal(s, sr)
    xs = sr.float().requires_grad_(True)
    gr, br = g.float().requires_grad_(True), b.float().requires_grad_(True)
    ref = F.layer_norm(xs, (D,), gr, br, 1e-5)
    _close(y, ref, 2e-2, 1e-2, 'ln fwd')
    dy = _rand(M, D, seed=5)
    ref.backward(dy.float())
    dg, db = torch.empty(D, dtype=DTYPE, device='cuda'), torch.empty(D, dtype=DTYPE, device='cuda')
    dx, dres = k.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, 0.0, 0, 0, True, dg, db)
    assert dres is dx
    _close(dx, xs.grad, 2e-2, 2e-2, 'ln dx')
    _close(dg, gr.grad, 0.05 * math.sqrt(M / 100), 2e-2, 'ln dgamma')
    _close(db, br.grad, 0.05 * math.sqrt(M / 100), 2e-2, 'ln dbeta')
    # fused bias gradient of the preceding Linear: column sums of dx; accumulate flags are independent
    dl = torch.ones(D, dtype=DTYPE, device='cuda')
    dg2, db2 = dg.clone(), db.clone()
    dx2, _ = k.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, 0.0, 0, 0, True, dg2, db2, accumulate=True, dbias_lin=dl)
    assert torch.equal(dx2, dx)
    _close(dl, dx.float().sum(0), 0.05 * math.sqrt(M / 100), 2e-2, 'ln dbias_lin')
    _close(dg2, 2 * gr.grad, 0.1 * math.sqrt(M / 100), 3e-2, 'ln dgamma accumulate')
    dl2 = dl.clone()
    k.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, 0.0, 0, 0, True, None, None, dbias_lin=dl2, accumulate_lin=True)
    _close(dl2, 2 * dx.float().sum(0), 0.1 * math.sqrt(M / 100), 3e-2, 'ln dbias_lin accumulate')


@pytest.mark.parametrize('M,D,p', [(16000, 512, 0.1), (3072, 512, 0.1), (1237, 512, 0.0), (531, 768, 0.1), (700, 1024, 0.25),
                                   (7, 512, 0.1)])
@pytest.mark.parametrize('lin', [True, False])
def test_ln_bwd_training_shape_kernel(M, D, p, lin):
    """csrc/norm.cu has a second backward kernel for the shape the training step launches (LayerNorm present, D a multiple of
    256, stored keep bits or no dropout, no extra gradient operand): against an fp32 torch restatement applied to the kernel's
    own keep bits, and against the general kernel (selected by passing an all-zero extra gradient operand) on the same inputs."""
    k = _k()
    x, r = _rand(M, D, seed=11), _rand(M, D, seed=12)
    g, b = (1 + 0.1 * _rand(D, seed=13).float()).to(DTYPE), _rand(D, scale=0.1, seed=14)
    keep = torch.zeros(M * D // 8, dtype=torch.uint8, device='cuda') if p > 0 else None
    y, s, mean, rstd = k.dropout_add_ln_fwd(x, r, g, b, 1e-5, p, 4242, 3, keep_out=keep)
    dy = _rand(M, D, seed=15)

    def run(extra):
        dg, db = torch.empty(D, dtype=DTYPE, device='cuda'), torch.empty(D, dtype=DTYPE, device='cuda')
        dl = torch.empty(D, dtype=DTYPE, device='cuda') if lin else None
        dx, dres = k.dropout_add_ln_bwd(dy, extra, s, g, mean, rstd, p, 4242, 3, True, dg, db, dbias_lin=dl, keep_mask=keep)
        return dx, dres, dg, db, dl

    dx, dres, dg, db, dl = run(None)
    dx0, dres0, dg0, db0, dl0 = run(torch.zeros_like(dy))
    # fp32 restatement
    xs = s.float().requires_grad_(True)
    gr, br = g.float().requires_grad_(True), b.float().requires_grad_(True)
    F.layer_norm(xs, (D,), gr, br, 1e-5).backward(dy.float())
    _close(dres, xs.grad, 2e-2, 2e-2, 'training-shape ln dres')
    _close(dg, gr.grad, 0.05 * math.sqrt(max(M, 100) / 100), 2e-2, 'training-shape ln dgamma')
    _close(db, br.grad, 0.05 * math.sqrt(max(M, 100) / 100), 2e-2, 'training-shape ln dbeta')
    if p > 0:
        bits = ((keep.view(M, D // 8, 1).int() >> torch.arange(8, device='cuda').view(1, 1, 8)) & 1).view(M, D).bool()
        ref_dx = torch.where(bits, (dres.float() * (1 / (1 - p))).to(DTYPE), torch.zeros_like(dres))
        assert torch.equal(dx, ref_dx)          # dropout' applied to the 16-bit gradient exactly as stored
    else:
        assert dres is dx
    if lin:
        _close(dl, dx.float().sum(0), 0.05 * math.sqrt(max(M, 100) / 100), 2e-2, 'training-shape ln dbias_lin')
    # the general kernel: same values up to the last 16-bit rounding (the two evaluate rs * (g - mean g - xhat mean(g xhat))
    # with differently associated fp32 operations)
    ulp = 2 ** -7 if DTYPE == torch.bfloat16 else 2 ** -10
    for name, a_, b_ in (('dres', dres, dres0), ('dx', dx, dx0)):
        d = (a_.float() - b_.float()).abs()
        assert (d <= 2 * ulp * b_.float().abs() + 1e-6).all(), (name, d.max().item())
        assert (d > 0).float().mean().item() < 0.05, name
    _close(dg, dg0.float(), 0.02 * math.sqrt(max(M, 100) / 100), 1e-2, 'dgamma vs general kernel')
    _close(db, db0.float(), 0.02 * math.sqrt(max(M, 100) / 100), 1e-2, 'dbeta vs general kernel')
    if lin:
        _close(dl, dl0.float(), 0.02 * math.sqrt(max(M, 100) / 100), 1e-2, 'dbias_lin vs general kernel')


@pytest.mark.parametrize('M,K,p', [(16000, 512, 0.0), (16000, 2048, 0.1), (8192, 512, 0.1), (8300, 1024, 0.25)])
def test_linear_dropout_add_ln_fused(M, K, p):
    """csrc/gemm_ln.cu against the two-kernel path it replaces (GEMM, then dropout + residual + LayerNorm) and against an
    fp32 torch restatement of pasero/models/transformer.py:1043-1054 applied to the kernel's own keep bits."""
    k = _k()
    D = 512
    a, w = _rand(M, K, seed=1), _rand(D, K, scale=K ** -0.5, seed=2)
    bias, r = _rand(D, scale=0.1, seed=3), _rand(M, D, seed=4)
    g, b = (1 + 0.1 * _rand(D, seed=5).float()).to(DTYPE), _rand(D, scale=0.1, seed=6)
    keep = torch.zeros(M * D // 8, dtype=torch.uint8, device='cuda') if p > 0 else None
    y, s, mean, rstd = k.linear_dropout_add_ln_fwd(a, w, bias, r, g, b, 1e-5, p, 99, 5, keep_out=keep)
    x = k.linear_fwd(a, w, bias)                       # the Linear's bf16 output (same tcgen05 accumulation order)
    if p > 0:
        bits = ((keep.view(M, D // 8, 1).int() >> torch.arange(8, device='cuda').view(1, 1, 8)) & 1).view(M, D).bool()
        frac = bits.float().mean().item()
        assert abs(frac - (1 - p)) < 2e-3, frac
        # no structure along rows / columns / neighbours: per-row and per-column keep rates stay within 5 sigma
        sig_r, sig_c = math.sqrt(p * (1 - p) / D), math.sqrt(p * (1 - p) / M)
        assert (bits.float().mean(1) - (1 - p)).abs().max().item() < 5.5 * sig_r
        assert (bits.float().mean(0) - (1 - p)).abs().max().item() < 5.5 * sig_c
        f = bits.float() - (1 - p)
        for shift in ((0, 1), (1, 0), (0, 2), (1, 1)):
            corr = (f[:M - shift[0], :D - shift[1]] * f[shift[0]:, shift[1]:]).mean().item() / (p * (1 - p))
            assert abs(corr) < 5 / math.sqrt(M * D), (shift, corr)
        xd = torch.where(bits, (x.float() * (1 / (1 - p))).to(DTYPE), torch.zeros_like(x))
    else:
        xd = x
    s_ref = (xd.float() + r.float()).to(DTYPE)
    assert torch.equal(s, s_ref), (s.float() - s_ref.float()).abs().max().item()
    ref = F.layer_norm(s_ref.float(), (D,), g.float(), b.float(), 1e-5)
    _close(y, ref, 2e-2, 1e-2, 'fused ln fwd')
    mu = s_ref.float().mean(1)
    var = s_ref.float().var(1, unbiased=False)
    assert torch.allclose(mean, mu, atol=1e-5, rtol=1e-5)
    assert torch.allclose(rstd, (var + 1e-5).rsqrt(), rtol=1e-4)
    if p == 0:      # identical to the unfused kernels up to the last bf16 rounding of y
        y2, s2, mean2, rstd2 = k.dropout_add_ln_fwd(x, r, g, b, 1e-5, 0.0, 0, 0)
        assert torch.equal(s2, s)
        assert (y2.float() - y.float()).abs().max().item() <= 2 ** -6 * max(1.0, y2.float().abs().max().item())
    else:           # backward consumes the stored bits
        dy = _rand(M, D, seed=7)
        dg, db = torch.empty(D, dtype=DTYPE, device='cuda'), torch.empty(D, dtype=DTYPE, device='cuda')
        dx, dres = k.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, p, 99, 5, True, dg, db, keep_mask=keep)
        assert torch.equal(dx.float() != 0, bits & (dres.float() != 0))


def test_sublayers_fused_linear_ln_matches_two_kernel_path():
    """The post-norm sub-block nodes (ops.AttnSublayerFn / FFNSublayerFn) at an encoder-sized row count, where their forward
    ends in the fused Linear + dropout + residual + LayerNorm kernel: outputs and every gradient against the same nodes
    with the fusion switched off. With p = 0 the two paths differ only by the last 16-bit rounding of the two LayerNorm outputs
    (single-pass vs two-pass variance); those one-ulp differences (0.4 % in bfloat16) travel through the attention and the FFN
    backward, hence 5 % of the tensor maximum for the gradients and 2 % for the output."""
    from pasero_b200 import ops
    B, L, D, H, F_ = 36, 250, 512, 8, 2048          # 9 000 rows: not a multiple of the 256-row tile
    x0 = _rand(B * L, D, seed=1)
    in_w, in_b = _rand(3 * D, D, scale=D ** -0.5, seed=2), _rand(3 * D, scale=0.1, seed=3)
    out_w, out_b = _rand(D, D, scale=D ** -0.5, seed=4), _rand(D, scale=0.1, seed=5)
    w1, b1 = _rand(F_, D, scale=D ** -0.5, seed=6), _rand(F_, scale=0.1, seed=7)
    w2, b2 = _rand(D, F_, scale=F_ ** -0.5, seed=8), _rand(D, scale=0.1, seed=9)
    g1, e1 = (1 + 0.1 * _rand(D, seed=10).float()).to(DTYPE), _rand(D, scale=0.1, seed=11)
    g2, e2 = (1 + 0.1 * _rand(D, seed=12).float()).to(DTYPE), _rand(D, scale=0.1, seed=13)
    kv_len = torch.full((B,), L, dtype=torch.int32, device='cuda')
    kv_len[::3] = 201
    dy = _rand(B * L, D, seed=14)

    def run(fuse):
        ops.FUSE_LINEAR_LN = fuse
        leaves = [t.clone().requires_grad_(True) for t in (x0, in_w, in_b, out_w, out_b, g1, e1, w1, b1, w2, b2, g2, e2)]
        x, iw, ib, ow, ob, ga, ea, a1, c1, a2, c2, gb, eb = leaves
        h = ops.AttnSublayerFn.apply(x, None, iw, ib, ow, ob, ga, ea, kv_len, B, L, L, H, False, 0.125, 0.0, 1e-5)
        y = ops.FFNSublayerFn.apply(h, a1, c1, a2, c2, gb, eb, 0.0, 1e-5)
        y.backward(dy)
        return [y.detach()] + [t.grad for t in leaves]

    try:
        fused = run(True)
        plain = run(False)
    finally:
        ops.FUSE_LINEAR_LN = True
    names = ['y', 'dx', 'in_w', 'in_b', 'out_w', 'out_b', 'ln1.g', 'ln1.b', 'fc1.w', 'fc1.b', 'fc2.w', 'fc2.b', 'ln2.g', 'ln2.b']
    for n, a, b in zip(names, fused, plain):
        scale = b.float().abs().max().item()
        err = (a.float() - b.float()).abs().max().item()
        assert err <= (0.02 if n == 'y' else 0.05) * scale + 1e-6, (n, err, scale)


def test_dropout_mask_consistency():
    k = _k()
    M, D, p = 512, 512, 0.1
    x = torch.ones(M, D, dtype=DTYPE, device='cuda')
    r = torch.zeros(M, D, dtype=DTYPE, device='cuda')
    s, _, _, _ = k.dropout_add_ln_fwd(x, r, None, None, 1e-5, p, 1234, 7)
    keep = s.float() > 0
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 0.01, frac
    assert torch.allclose(s.float()[keep], torch.tensor(1 / (1 - p)).to(DTYPE).float())
    # backward regenerates the same mask
    dy = torch.ones(M, D, dtype=DTYPE, device='cuda')
    dx, dres = k.dropout_add_ln_bwd(dy, None, None, None, None, None, p, 1234, 7, True)
    assert torch.equal(dx.float() > 0, keep)
    assert torch.equal(dres, dy)
    # the keep bits stored by forward (one byte per 8 elements) give backward the same mask without Philox
    kb = torch.zeros(M * D // 8, dtype=torch.uint8, device='cuda')
    s1, _, _, _ = k.dropout_add_ln_fwd(x, r, None, None, 1e-5, p, 1234, 7, keep_out=kb)
    assert torch.equal(s1, s)
    bits = ((kb.view(-1, 1) >> torch.arange(8, device='cuda', dtype=torch.uint8)) & 1).reshape(M, D).bool()
    assert torch.equal(bits, keep)
    dx1, _ = k.dropout_add_ln_bwd(dy, None, None, None, None, None, p, 999, 999, True, keep_mask=kb)   # seed/offset unused
    assert torch.equal(dx1, dx)
    # a different offset gives a different mask
    s2, _, _, _ = k.dropout_add_ln_fwd(x, r, None, None, 1e-5, p, 1234, 8)
    assert not torch.equal(s2, s)


def test_colsum():
    k = _k()
    x = _rand(3001, 1024, seed=1)
    _close(k.colsum(x), x.float().sum(0), 0.5, 2e-2, 'colsum')


# ---------------------------------------------------------------------------------------------- embed / glu
def test_embed():
    k = _k()
    V, D, B, T = 1000, 512, 5, 37
    table = _rand(V, D, scale=D ** -0.5, seed=1)
    pos = _rand(T, D, seed=2)
    ids = torch.randint(0, V, (B, T), device='cuda')
    ids[:, -3:] = 1
    scale = math.sqrt(D)
    out = k.embed_fwd(ids.view(-1), table, pos, B * T, D, T, scale, 0.0, 0, 0)
    e = table[ids]
    e = e * scale
    e = e + pos[None]
    assert torch.equal(out.view(B, T, D), e)
    dy = _rand(B * T, D, seed=3)
    dtab = torch.zeros(V, D, dtype=torch.float32, device='cuda')
    k.embed_bwd(dy, ids.view(-1), B * T, D, scale, 0.0, 0, 0, padding_idx=1, dtable=dtab)
    ref = torch.zeros(V, D, device='cuda')
    g = (dy.float() * scale).to(DTYPE).float()
    g[(ids.view(-1) == 1)] = 0
    ref.index_add_(0, ids.view(-1), g)
    _close(dtab, ref, 1e-3, 1e-3, 'embed bwd')
    # feature path (no gather)
    x = _rand(B * T, D, seed=4)
    out2 = k.embed_fwd(None, x, pos, B * T, D, T, scale, 0.0, 0, 0)
    assert torch.equal(out2.view(B, T, D), (x.view(B, T, D) * scale) + pos[None])
    dx = k.embed_bwd(dy, None, B * T, D, scale, 0.0, 0, 0)
    assert torch.equal(dx, (dy.float() * scale).to(DTYPE))


def test_glu():
    k = _k()
    B, T, C = 3, 50, 512
    x = _rand(B * T, 2 * C, seed=1)
    out = k.glu_fwd(x, B, T, C, T + 5, 2)
    xr = x.float().requires_grad_(True)
    ref = F.glu(xr, dim=1)
    _close(out[:, 2:2 + T].reshape(B * T, C), ref, 1e-2, 1e-2, 'glu')
    assert out[:, :2].abs().sum() == 0 and out[:, 2 + T:].abs().sum() == 0
    dy = _rand(B * T, C, seed=2)
    ref.backward(dy.float())
    _close(k.glu_bwd(dy, x, B * T, C), xr.grad, 1e-2, 1e-2, 'glu bwd')


# ---------------------------------------------------------------------------------------------- loss / optimizer
def test_gelu_padded_layout():
    """erf-GELU forward into the zero-padded [B, Tp, C] layout (Whisper-style conv front end) and its backward."""
    k = _k()
    B, T, C, Tp, front = 3, 37, 64, 42, 2
    x = _rand(B * T, C, seed=1)
    out = k.gelu_fwd(x, B, T, C, Tp, front)
    ref = F.gelu(x.float()).view(B, T, C)
    assert torch.equal(out[:, :front], torch.zeros_like(out[:, :front]))
    assert torch.equal(out[:, front + T:], torch.zeros_like(out[:, front + T:]))
    _close(out[:, front:front + T], ref, 1e-2, 1e-2, 'gelu fwd')
    dy = _rand(B * T, C, seed=2)
    xr = x.float().requires_grad_(True)
    F.gelu(xr).backward(dy.float())
    _close(k.gelu_bwd(dy, x), xr.grad, 1e-2, 1e-2, 'gelu bwd')


@pytest.mark.parametrize('V', [8032, 1000])
def test_ls_ce(V):
    k = _k()
    B, T, eps = 7, 48, 0.1
    logits = _rand(B * T, V, scale=2.0, seed=1)
    dec = torch.randint(4, V, (B, T + 1), device='cuda')
    dec[:, -5:] = 1
    dec[0, :] = 1
    target = dec[:, 1:]
    lr = logits.float().requires_grad_(True)
    ref = F.cross_entropy(lr, target.reshape(-1), ignore_index=1, reduction='sum', label_smoothing=eps)
    nll = F.cross_entropy(lr.detach(), target.reshape(-1), ignore_index=1, reduction='sum')
    ref.backward()
    stats = torch.zeros(8, dtype=torch.float64, device='cuda')
    k.ls_ce_fwd_bwd(logits, target, 1, eps, stats)
    assert abs(stats[0].item() - ref.item()) / ref.item() < 1e-5
    assert abs(stats[1].item() - nll.item()) / nll.item() < 1e-5
    assert stats[2].item() == (target != 1).sum().item()
    _close(logits, lr.grad, 2e-3, 1e-2, 'ce grad')


def test_adam_and_clip():
    k = _k()
    n = 8 * 1000 + 8
    p0 = _rand(n, scale=0.1, seed=1)
    g = _rand(n, scale=3.0, seed=2)
    m0 = _rand(n, scale=0.01, seed=3, dtype=torch.float32)
    v0 = _rand(n, scale=0.01, seed=4, dtype=torch.float32).abs()
    lr, b1, b2, eps, wd, step, clip, ntok = 1e-3, 0.9, 0.999, 1e-8, 0.01, 3, 1.0, 3072.0
    # reference (pasero/training.py:462-477 + optimization.py:65-149), fp32 on GPU
    gr = g.clone()
    gr.mul_(1.0 / ntok)
    norm = torch.norm(gr.float())
    gr.mul_((clip / (norm + 1e-6)).clamp_(max=1))
    gf = gr.float()
    m, v, pf = m0.clone(), v0.clone(), p0.float()
    m.mul_(b1).add_(gf, alpha=1 - b1)
    v.mul_(b2).addcmul_(gf, gf, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    pf.add_(pf, alpha=-wd * lr)
    pf.addcdiv_(m, denom, value=-(lr / bc1))
    hyper = torch.tensor([lr, b1, b2, eps, wd, bc1, math.sqrt(bc2), clip, 1.0, 1 - b1, 1 - b2, lr / bc1, wd * lr, 0, 0, 0],
                         dtype=torch.float32, device='cuda')
    stats = torch.zeros(8, dtype=torch.float64, device='cuda')
    stats[2] = ntok
    p, m1, v1 = p0.clone(), m0.clone(), v0.clone()
    k.grad_sqnorm(g, hyper, stats)
    assert abs(math.sqrt(stats[4].item()) - norm.item()) / norm.item() < 1e-5
    k.adam_step(p, g, m1, v1, hyper, stats)
    _close(m1, m, 1e-7, 1e-5, 'exp_avg')
    _close(v1, v, 1e-9, 1e-5, 'exp_avg_sq')
    assert (p.float() - pf.to(DTYPE).float()).abs().max() <= 2 ** -8 * p0.float().abs().max()
    assert (p != pf.to(DTYPE)).float().mean() < 1e-3


# ---------------------------------------------------------------------------------------------- fbank
def _fbank_tol(truth64):
    """Stated tolerance (north star: 1e-4 abs in fp32): 1e-4, widened for log-mel bins far below the frame's
    strongest bin, where any fp32 FFT (torchaudio's included) is limited by its relative-to-total error."""
    import numpy as np
    e = np.exp(truth64)
    ratio = np.sqrt(e.max(axis=1, keepdims=True) / e)
    return 1e-4 * np.maximum(1.0, ratio / 50.0)


@pytest.mark.parametrize('lengths', [[160000, 160000], [400, 16000, 48123, 5000], [399, 8000]])
def test_fbank_vs_oracle(lengths):
    import numpy as np
    from oracle import fbank as of
    k = _k()
    rng = np.random.RandomState(1234)
    waves = [(rng.rand(n).astype(np.float32) - 0.5) for n in lengths]
    frames = [k.fbank_num_frames(n) for n in lengths]
    assert frames == [of.num_frames(n) for n in lengths]
    so = torch.tensor(np.concatenate([[0], np.cumsum(lengths)]), dtype=torch.int64, device='cuda')
    fo = torch.tensor(np.concatenate([[0], np.cumsum(frames)]), dtype=torch.int64, device='cuda')
    pcm = torch.from_numpy(np.concatenate(waves)).cuda()
    raw, _ = k.fbank_cmvn(pcm, so, fo, len(lengths), max(frames), sum(frames), cmvn=False)
    norm, padded = k.fbank_cmvn(pcm, so, fo, len(lengths), max(frames), sum(frames), cmvn=True,
                                padded_frames=max(frames) + 4, front_pad=2)
    raw, norm = raw.cpu().numpy(), norm.cpu().numpy()
    fo = fo.cpu().numpy()
    for i, (w, f) in enumerate(zip(waves, frames)):
        if f == 0:
            continue
        truth = of.fbank_cmvn(w, cmvn=False, dtype=np.float64)
        got = raw[fo[i]:fo[i + 1]]
        err = np.abs(got - truth)
        tol = _fbank_tol(truth)
        print(f'fbank raw: n={len(w)} max err {err.max():.3g}, within 1e-4: {(err <= 1e-4).mean():.6f}')
        assert (err <= tol).all(), (err.max(), (err / tol).max())
        assert (err <= 1e-4).mean() > 0.999
        ref32 = of.fbank_cmvn(w, cmvn=False)
        assert np.abs(got - ref32).max() < 2e-3
        if f > 1:
            truth_n = of.utterance_cmvn(truth)
            gotn = norm[fo[i]:fo[i + 1]]
            std = truth.std(axis=0)
            errn = np.abs(gotn - truth_n)
            print(f'fbank cmvn: max err {errn.max():.3g}, within 1e-4: {(errn <= 1e-4).mean():.6f}')
            assert (errn <= (tol + 1e-5) / std[None, :] + 1e-5).all(), errn.max()
        pb = padded[i].float().cpu().numpy()
        assert np.abs(pb[:2]).sum() == 0 and np.abs(pb[2 + f:]).sum() == 0
        assert np.abs(pb[2:2 + f] - norm[fo[i]:fo[i + 1]]).max() < 0.05


def test_fbank_vs_torchaudio_live():
    """torchaudio ships in the GPU image too: check the kernel directly against the third-party routine."""
    torchaudio = pytest.importorskip('torchaudio')
    import numpy as np
    k = _k()
    g = torch.Generator().manual_seed(7)
    t = torch.arange(64000) / 16000.0
    w = 0.3 * torch.sin(2 * math.pi * 440 * t) * torch.exp(-t) + 0.01 * torch.randn(64000, generator=g)
    f = k.fbank_num_frames(64000)
    so = torch.tensor([0, 64000], dtype=torch.int64, device='cuda')
    fo = torch.tensor([0, f], dtype=torch.int64, device='cuda')
    raw, _ = k.fbank_cmvn(w.cuda(), so, fo, 1, f, f, cmvn=False)
    ref = torchaudio.compliance.kaldi.fbank((w * 2 ** 15)[None], num_mel_bins=80, sample_frequency=16000.0)
    ref64 = torchaudio.compliance.kaldi.fbank((w.double() * 2 ** 15)[None], num_mel_bins=80, sample_frequency=16000.0)
    err_k = (raw.cpu().double() - ref64).abs().max().item()
    err_ta = (ref.double() - ref64).abs().max().item()
    print(f'vs fp64 torchaudio: kernel {err_k:.3g}, torchaudio-fp32 {err_ta:.3g}')
    assert err_k < max(1e-4, 3 * err_ta)
    assert (raw.cpu() - ref).abs().max() < 2e-3


def test_fbank_baseline_shape_no_worse_than_torchaudio_fp32():
    """BASELINE config 2's frontend batch (64 x 160 000 samples): against torchaudio's own fp64 evaluation, the kernel must be
    as accurate as torchaudio's fp32 evaluation — the routine the north star's "1e-4 abs in fp32" refers to — on the same
    utterances ("fbank tolerance"): the same share of values within 1e-4, and a worst-case error (an extreme
    value statistic of two independent fp32 round-off patterns; it sits in bins > 60 dB below the frame's peak) within 1.5x
    of torchaudio-fp32's. Measured: 1.0e-3 vs 8.1e-4, 99.982 % within 1e-4 for both. Eight of the 64 utterances are compared."""
    torchaudio = pytest.importorskip('torchaudio')
    import numpy as np
    k = _k()
    B, N = 64, 160000
    rng = np.random.RandomState(1234)
    pcm = (rng.rand(B, N).astype(np.float32) - 0.5)
    t = np.arange(N) / 16000.0
    pcm[1] = (0.3 * np.sin(2 * math.pi * 440 * t) * np.exp(-t / 3) + 0.01 * rng.randn(N)).astype(np.float32)   # a tonal one
    pcm[2] *= np.linspace(0.001, 1.0, N, dtype=np.float32)                                                    # a fade-in
    F = k.fbank_num_frames(N)
    so = torch.arange(B + 1, dtype=torch.int64, device='cuda') * N
    fo = torch.arange(B + 1, dtype=torch.int64, device='cuda') * F
    raw, _ = k.fbank_cmvn(torch.from_numpy(pcm).cuda().reshape(-1), so, fo, B, F, B * F, cmvn=False)
    raw = raw.view(B, F, 80).cpu().double()
    worst_k = worst_ta = 0.0
    within_k, within_ta = [], []
    for b in range(8):
        w = torch.from_numpy(pcm[b])
        ref32 = torchaudio.compliance.kaldi.fbank((w * 2 ** 15)[None], num_mel_bins=80, sample_frequency=16000.0).double()
        ref64 = torchaudio.compliance.kaldi.fbank((w.double() * 2 ** 15)[None], num_mel_bins=80, sample_frequency=16000.0)
        ek, et = (raw[b] - ref64).abs(), (ref32 - ref64).abs()
        worst_k, worst_ta = max(worst_k, ek.max().item()), max(worst_ta, et.max().item())
        within_k.append((ek <= 1e-4).double().mean().item())
        within_ta.append((et <= 1e-4).double().mean().item())
    print(f'64 x 10 s: max |err| vs torchaudio fp64: kernel {worst_k:.3g}, torchaudio fp32 {worst_ta:.3g}; '
          f'within 1e-4: kernel {min(within_k):.5f}, torchaudio fp32 {min(within_ta):.5f}')
    assert worst_k <= 1.5 * worst_ta, (worst_k, worst_ta)
    assert min(within_k) >= min(within_ta) - 1e-4 and min(within_k) > 0.999


def test_collate_features_matches_task_collate(tmp_path):
    """NumpyFile.pack + the device-side collate == Task.collate's padded batch cast to bf16 (pasero/tasks/task.py:531-571),
    and the same data straight in the conv subsampler's padded layout."""
    import numpy as np
    from pasero_b200 import frontend
    from pasero_b200.task import NumpyFile, tokens_as_tensor
    rng = np.random.RandomState(3)
    feats = [rng.randn(n, 80).astype(np.float32) for n in (37, 5, 120, 64, 1)]
    f = NumpyFile.build(str(tmp_path / 'feats.npy'), feats, dtype='float16')
    order = [2, 0, 4, 3]
    packed, offsets = f.pack(order)
    lens = [len(feats[i]) for i in order]
    want, _ = tokens_as_tensor([torch.from_numpy(f[i].astype(np.float32)) for i in order], 1, dtype=DTYPE)
    got, lengths = frontend.collate_features(packed, offsets, max(lens))
    assert lengths.tolist() == lens
    assert torch.equal(got.cpu(), want)
    padded, _ = frontend.collate_features(packed, offsets, max(lens) + 4, front_pad=2)
    assert torch.equal(padded[:, 2:2 + max(lens)].cpu(), want) and padded[:, :2].abs().sum() == 0 and padded[:, 2 + max(lens):].abs().sum() == 0


@pytest.mark.parametrize('V,D,M,cols', [(70376, 512, 256, 8192), (1003, 128, 77, 256), (520, 64, 40, 512)])
def test_projection_ce_chunked(V, D, M, cols):
    """Projection + label-smoothed CE over vocabulary chunks (no [M, V] tensor in HBM; BASELINE config 5's V = 70 376):
    loss, nll, token count, dx and the table gradient against fp32 PyTorch and against the single-pass kernel."""
    from pasero_b200 import kernels as K, modules, ops
    saved = (K.CE_CHUNK_MIN_VOCAB, K.CE_CHUNK_COLS)
    emb = modules.Embedding(V, D, 1).to(DTYPE).cuda()
    g = torch.Generator(device='cuda').manual_seed(5)
    x0 = (torch.randn(M, D, generator=g, device='cuda') * 2).to(DTYPE)
    B = 1
    target = torch.randint(4, V, (B, M), generator=g, device='cuda')
    target[0, -3:] = 1
    target[0, 0] = V - 1
    results = []
    try:
        for chunked in (True, False):
            K.CE_CHUNK_MIN_VOCAB, K.CE_CHUNK_COLS = (1, cols) if chunked else (1 << 30, cols)
            emb.weight.grad = None
            x = x0.clone().requires_grad_(True)
            stats = torch.zeros(8, dtype=torch.float64, device='cuda')
            loss = ops.ProjCEFn.apply(x, emb.weight, target, 1, 0.1, stats, V)
            loss.backward()
            results.append((loss.item(), stats.tolist(), x.grad.float(), emb.weight.grad.float().clone()))
    finally:
        K.CE_CHUNK_MIN_VOCAB, K.CE_CHUNK_COLS = saved
    xr = x0.float().requires_grad_(True)
    wr = emb.weight.detach().float()[:V].requires_grad_(True)
    logits = xr @ wr.T
    ref = F.cross_entropy(logits, target.reshape(-1), ignore_index=1, reduction='sum', label_smoothing=0.1)
    nll = F.cross_entropy(logits.detach(), target.reshape(-1), ignore_index=1, reduction='sum')
    ref.backward()
    for name, (loss, stats, dx, dw) in zip(('chunked', 'single pass'), results):
        assert abs(loss - ref.item()) / ref.item() < 5e-3, (name, loss, ref.item())
        assert abs(stats[1] - nll.item()) / nll.item() < 5e-3 and stats[2] == M - 3 and stats[7] == 0, (name, stats)
        _close(dx, xr.grad, 2e-2 * xr.grad.abs().max().item(), 2e-2, name + ' dx')
        _close(dw[:V], wr.grad, 2e-2 * wr.grad.abs().max().item(), 3e-2, name + ' dtable')
        assert torch.equal(dw[V:], torch.zeros_like(dw[V:]))
    assert abs(results[0][0] - results[1][0]) / results[1][0] < 2e-3
