"""Sweep of the GEMM launch choices (tile width, CTA-pair / single CTAs, split-K count) over the shapes of the config-2
training step, against what the library's own heuristic picks (tile_n = cluster_m = split_k = 0). Each figure is the device
time of a chain of back-to-back launches of one configuration divided by the chain length (L2-warm operands, PDL-chained like
the step), median of several chains. Prints, per shape, the heuristic's time and every override that beats it by > 3 %."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pasero_b200 import kernels as k  # noqa: E402

BF16 = torch.bfloat16
CHAIN, REPS = 10, 7


def chain_time(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(CHAIN):
            fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / CHAIN * 1e3)
    return sorted(ts)[len(ts) // 2]


def rnd(*shape):
    return torch.randn(*shape, device='cuda').to(BF16)


def make(kind, m, n, kk):
    """Returns launch(tile_n, cluster_m, split_k) for one shape as the step issues it (tags of bench.py BENCH_GEMM_DUMP)."""
    if kind in ('fwd', 'fwd_relu'):          # y[m, n] = x[m, k] w[n, k]^T + bias (+ ReLU)
        x, w, b, y = rnd(m, kk), rnd(n, kk), rnd(n), torch.empty(m, n, dtype=BF16, device='cuda')
        act = 1 if kind == 'fwd_relu' else 0
        return lambda t, c, s: k.gemm_raw(x.data_ptr(), w.data_ptr(), y, m, n, kk, lda=kk, ldb=kk, ldc=n, bias=b, act=act,
                                          split_k=1, tile_n=t, cluster_m=c, elem_dtype=BF16)
    if kind in ('dgrad', 'dgrad_res'):       # dx[m, n] = dy[m, k] w[k, n] (+ residual)
        dy, w, dx = rnd(m, kk), rnd(kk, n), torch.empty(m, n, dtype=BF16, device='cuda')
        res = rnd(m, n) if kind == 'dgrad_res' else None
        return lambda t, c, s: k.gemm_raw(dy.data_ptr(), w.data_ptr(), dx, m, n, kk, lda=kk, ldb=n, ldc=n, b_mn_major=True,
                                          residual=res, split_k=1, tile_n=t, cluster_m=c, elem_dtype=BF16)
    # wgrad: dw[m, n] (fp32, += ) = dy[k, m]^T x[k, n], bias gradient fused
    dy, x = rnd(kk, m), rnd(kk, n)
    dw, db = torch.zeros(m, n, dtype=torch.float32, device='cuda'), torch.zeros(m, dtype=torch.float32, device='cuda')
    return lambda t, c, s: k.gemm_raw(dy.data_ptr(), x.data_ptr(), dw, m, n, kk, lda=m, ldb=n, ldc=n, a_mn_major=True,
                                      b_mn_major=True, accumulate=True, a_rowsum=db, split_k=s, tile_n=t, cluster_m=c,
                                      elem_dtype=BF16)


SHAPES = [  # (kind, m, n, k, launches per step)
    ('fwd_relu', 16000, 2048, 512, 12), ('dgrad_res', 16000, 512, 2048, 12), ('dgrad', 16000, 2048, 512, 12),
    ('fwd', 16000, 1536, 512, 12), ('dgrad_res', 16000, 512, 1536, 12), ('dgrad', 16000, 512, 512, 12),
    ('fwd', 16000, 1024, 512, 6), ('dgrad', 16000, 512, 1024, 6),
    ('wgrad', 2048, 512, 16000, 12), ('wgrad', 512, 2048, 16000, 12), ('wgrad', 1536, 512, 16000, 12),
    ('wgrad', 512, 512, 16000, 12), ('wgrad', 1024, 512, 16000, 6),
    ('fwd', 3072, 512, 512, 18), ('dgrad', 3072, 512, 512, 12), ('dgrad_res', 3072, 512, 512, 6),
    ('fwd', 3072, 1536, 512, 6), ('dgrad_res', 3072, 512, 1536, 6), ('fwd_relu', 3072, 2048, 512, 6),
    ('dgrad_res', 3072, 512, 2048, 6), ('fwd', 3072, 512, 2048, 6), ('dgrad', 3072, 2048, 512, 6),
    ('wgrad', 512, 512, 3072, 18), ('wgrad', 2048, 512, 3072, 6), ('wgrad', 512, 2048, 3072, 6), ('wgrad', 1536, 512, 3072, 6),
    ('fwd', 3072, 8032, 512, 1), ('dgrad', 3072, 512, 8032, 1), ('wgrad', 8032, 512, 3072, 1),
]


def main():
    total_base = total_best = 0.0
    for kind, m, n, kk, count in SHAPES:
        launch = make(kind, m, n, kk)
        base = chain_time(lambda: launch(0, 0, 0))
        rows = []
        tiles = (128, 256) if n >= 256 else (128,)
        splits = (0,) if kind != 'wgrad' else (0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 16)
        for t in tiles:
            for c in (1, 2):
                for s in splits:
                    try:
                        rows.append((chain_time(lambda: launch(t, c, s)), t, c, s))
                    except RuntimeError as err:      # a combination the library rejects
                        rows.append((float('inf'), t, c, s))
                        sys.stderr.write(f'{kind} {m}x{n}x{kk} tile {t} cl {c} split {s}: {str(err)[:80]}\n')
        rows.sort()
        best = rows[0]
        total_base += base * count
        total_best += min(base, best[0]) * count
        flop = 2.0 * m * n * kk
        line = f'{kind:9s} {m:6d}x{n:5d}x{kk:6d} n={count:2d}  heuristic {base:7.2f} us ({flop / base / 1e6:7.1f} TFLOP/s)'
        better = [r for r in rows if r[0] < 0.97 * base][:3]
        for tt, t, c, s in better:
            line += f' | tile {t} cl {c} split {s}: {tt:6.2f} ({100 * (1 - tt / base):4.1f} %)'
        print(line, flush=True)
    print(f'sum over the step: heuristic {total_base / 1e3:.3f} ms, best per shape {total_best / 1e3:.3f} ms')


if __name__ == '__main__':
    main()
