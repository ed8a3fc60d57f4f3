"""Per-kernel timings on the B200 (CUDA events, warm-up, L2 flush between iterations)."""
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasero_b200 import kernels as k  # noqa: E402

BF16 = torch.bfloat16
_flush = None


def timeit(fn, iters=20, warmup=5, flush=True):
    """Median device time of fn() over `iters` launches. The GPU is kept busy (a long spin kernel) while the host
    enqueues every iteration, so the events bracket the kernels themselves and not the host's launch path; a 256 MB
    write between iterations evicts the 126 MB L2."""
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.05 * 1.9e9))
    evs = []
    for _ in range(iters):
        if flush:
            _flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2]


def attn(out, dev):
    # attention
    for (Bq, H, Lq, Lk, causal) in [(64, 8, 250, 250, False), (64, 8, 48, 48, True), (64, 8, 48, 250, False)]:
        q = torch.randn(Bq * Lq, H * 64, device=dev).to(BF16)
        kv = torch.randn(Bq * Lk, 2 * H * 64, device=dev).to(BF16)
        kl = torch.full((Bq,), Lk, device=dev, dtype=torch.int32)
        fl = 4.0 * Bq * H * Lq * Lk * 64 * (0.5 if causal else 1)
        t = timeit(lambda: k.attention_fwd(q, kv[:, :H * 64], kv[:, H * 64:], Bq, H, Lq, Lk, None if causal else kl, causal, 0.125))
        out[f'attn fwd {Bq}x{H}x{Lq}x{Lk}'] = (t, fl / t / 1e9)
        o, lse = k.attention_fwd(q, kv[:, :H * 64], kv[:, H * 64:], Bq, H, Lq, Lk, None if causal else kl, causal, 0.125)
        do = torch.randn_like(o)
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        t = timeit(lambda: k.attention_bwd(do, q, kv[:, :H * 64], kv[:, H * 64:], o, lse, Bq, H, Lq, Lk, None if causal else kl,
                                           causal, 0.125, dq, dkv[:, :H * 64], dkv[:, H * 64:]))
        out[f'attn bwd {Bq}x{H}x{Lq}x{Lk}'] = (t, 2.5 * fl / t / 1e9)
    for name, (t, rate) in out.items():
        if name.startswith('attn'):
            print(f'{name:40s} {t * 1000:10.1f} us   {rate:10.1f} TFLOP/s')


def main():
    out = {}
    dev = 'cuda'
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only == 'attn':
        return attn(out, dev)
    gemm_only = only == 'gemm'
    # GEMM shapes of config 2 (M = 64 * 250 = 16000 encoder rows)
    for (M, N, K) in [(16000, 512, 512), (16000, 1536, 512), (16000, 2048, 512), (16000, 512, 2048), (31936, 1024, 400),
                      (16000, 1024, 2560), (3072, 8032, 512), (8192, 8192, 8192)]:
        x = torch.randn(M, K, device=dev).to(BF16)
        w = torch.randn(N, K, device=dev).to(BF16)
        b = torch.randn(N, device=dev).to(BF16)
        dy = torch.randn(M, N, device=dev).to(BF16)
        fl = 2.0 * M * N * K
        for tn, cl in ((128, 1), (256, 1), (256, 2), (256, 4), (512, 2)):
            t = timeit(lambda: k.linear_fwd(x, w, b, tile_n=tn, cluster_m=cl))
            out[f'fwd {M}x{N}x{K} tn{tn} cl{cl}'] = (t, fl / t / 1e9)
        t = timeit(lambda: torch.nn.functional.linear(x, w, b))
        out[f'cublas fwd {M}x{N}x{K}'] = (t, fl / t / 1e9)
        t = timeit(lambda: k.linear_dgrad(dy, w))
        out[f'dgrad {M}x{N}x{K}'] = (t, fl / t / 1e9)
        t = timeit(lambda: k.linear_wgrad(dy, x))
        out[f'wgrad {M}x{N}x{K}'] = (t, fl / t / 1e9)
    if gemm_only:
        for name, (t, rate) in out.items():
            print(f'{name:40s} {t * 1000:10.1f} us   {rate:10.1f} TFLOP/s')
        return
    # fbank
    B, N = 64, 160000
    pcm = (torch.rand(B * N, device=dev) - 0.5)
    F = k.fbank_num_frames(N)
    so = torch.arange(B + 1, device=dev, dtype=torch.int64) * N
    fo = torch.arange(B + 1, device=dev, dtype=torch.int64) * F
    t = timeit(lambda: k.fbank_cmvn(pcm, so, fo, B, F, B * F, cmvn=True, padded_frames=F + 4, front_pad=2))
    nbytes = 4 * B * N + 320 * B * F
    out['fbank+cmvn 64x160000'] = (t, nbytes / t / 1e6)
    t = timeit(lambda: k.fbank_cmvn(pcm, so, fo, B, F, B * F, cmvn=False))
    out['fbank raw 64x160000'] = (t, nbytes / t / 1e6)
    attn(out, dev)
    # LN, CE
    M, D = 16000, 512
    x, r = torch.randn(M, D, device=dev).to(BF16), torch.randn(M, D, device=dev).to(BF16)
    g, b = torch.ones(D, device=dev).to(BF16), torch.zeros(D, device=dev).to(BF16)
    t = timeit(lambda: k.dropout_add_ln_fwd(x, r, g, b, 1e-5, 0.1, 1, 1))
    out['ln fwd 16000x512'] = (t, 4 * M * D * 2 / t / 1e6)
    logits = torch.randn(3072, 8032, device=dev).to(BF16)
    tgt = torch.randint(4, 8032, (64, 48), device=dev)
    stats = torch.zeros(8, dtype=torch.float64, device=dev)
    t = timeit(lambda: k.ls_ce_fwd_bwd(logits, tgt, 1, 0.1, stats))
    out['ce 3072x8032'] = (t, 2 * 3072 * 8032 * 2 / t / 1e6)
    for name, (t, rate) in out.items():
        print(f'{name:40s} {t * 1000:10.1f} us   {rate:10.1f} {"GB/s" if ("fbank" in name or "ln " in name or "ce " in name) else "TFLOP/s"}')
    json.dump(out, open('gpurun_out/microbench.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
