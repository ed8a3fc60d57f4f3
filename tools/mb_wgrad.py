"""Weight-gradient GEMM (fp32 reduce-add sink) with and without the fused bias gradient across split-K / cluster choices;
device time of back-to-back launches inside a CUDA graph (inputs far larger than L2 are rotated)."""
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


for (M, N, K) in [(16000, 2048, 512), (16000, 1536, 512), (16000, 1024, 512), (3072, 2048, 512), (3072, 1536, 512)]:
    dys = [torch.randn(M, N, device=dev).to(BF16) for _ in range(4)]
    xs = [torch.randn(M, K, device=dev).to(BF16) for _ in range(4)]
    g32 = torch.zeros(N, K, dtype=torch.float32, device=dev)
    b32 = torch.zeros(N, dtype=torch.float32, device=dev)
    fl = 2.0 * M * N * K
    for fused in (0, 1):
        for sk, cl in ((0, 0), (4, 2), (4, 1), (3, 2), (6, 1), (6, 2), (9, 2), (9, 1), (12, 1), (18, 1)):
            i = [0]

            def run():
                i[0] = (i[0] + 1) % 4
                k.gemm_raw(dys[i[0]].data_ptr(), xs[i[0]].data_ptr(), g32, N, K, M, lda=N, ldb=K, ldc=K, a_mn_major=True, elem_dtype=torch.bfloat16,
                           b_mn_major=True, accumulate=True, split_k=sk, cluster_m=cl, a_rowsum=b32 if fused else None)
            try:
                t = graph_time(run, n=40)
            except Exception as e:  # noqa: BLE001
                print(f'wgrad {N}x{K}x{M} fused{fused} sk{sk} cl{cl}: {e}')
                continue
            print(f'wgrad {N}x{K}x{M} fused{fused} sk{sk} cl{cl}: {t:7.1f} us  {fl / t / 1e6:8.1f} TFLOP/s')
