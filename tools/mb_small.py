"""Fixed-cost floor of small GEMMs (decoder shapes) across tile / cluster choices; device time inside a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasero_b200 import kernels as k
BF16 = torch.bfloat16
dev = 'cuda'


def graph_time(fn, n=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


for (M, N, K) in [(3072, 512, 512), (3072, 1536, 512), (3072, 2048, 512), (3072, 512, 2048), (16000, 512, 512), (128, 128, 64)]:
    x = torch.randn(M, K, device=dev).to(BF16)
    w = torch.randn(N, K, device=dev).to(BF16)
    b = torch.randn(N, device=dev).to(BF16)
    dy = torch.randn(M, N, device=dev).to(BF16)
    y = torch.empty(M, N, dtype=BF16, device=dev)
    fl = 2.0 * M * N * K
    for tn, cl in ((128, 1), (128, 2), (256, 1), (256, 2)):
        t = graph_time(lambda: k.gemm_raw(x.data_ptr(), w.data_ptr(), y, M, N, K, lda=K, ldb=K, ldc=N, bias=b, split_k=1, tile_n=tn, cluster_m=cl))
        print(f'fwd {M}x{N}x{K} tn{tn} cl{cl}: {t:7.1f} us  {fl / t / 1e6:8.1f} TFLOP/s')
    t = graph_time(lambda: torch.nn.functional.linear(x, w, b))
    print(f'cublas fwd {M}x{N}x{K}: {t:7.1f} us  {fl / t / 1e6:8.1f} TFLOP/s')
    g32 = torch.zeros(N, K, dtype=torch.float32, device=dev)
    t = graph_time(lambda: k.linear_wgrad(dy, x, out=g32, accumulate=True))
    print(f'wgrad32 {N}x{K}x{M}: {t:7.1f} us  {fl / t / 1e6:8.1f} TFLOP/s')
z = torch.zeros(1024, device=dev)
t = graph_time(lambda: z.add_(1.0))
print(f'tiny elementwise kernel in a graph chain: {t:.2f} us')
