import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasero_b200 import kernels as k
from tools.microbench import timeit
BF16 = torch.bfloat16
dev = 'cuda'
M, D = 16000, 512
x, r = torch.randn(M, D, device=dev).to(BF16), torch.randn(M, D, device=dev).to(BF16)
g, b = torch.ones(D, device=dev).to(BF16), torch.zeros(D, device=dev).to(BF16)
kb = torch.empty(M * D // 8, dtype=torch.uint8, device=dev)
y, s, mean, rstd = k.dropout_add_ln_fwd(x, r, g, b, 1e-5, 0.1, 1, 1, keep_out=kb)
dy = torch.randn(M, D, device=dev).to(BF16)
dg, db, dl = (torch.empty(D, dtype=BF16, device=dev) for _ in range(3))
for name, fn in [('ln fwd', lambda: k.dropout_add_ln_fwd(x, r, g, b, 1e-5, 0.1, 1, 1, keep_out=kb)),
                 ('ln bwd keep', lambda: k.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, 0.1, 1, 1, True, dg, db, dbias_lin=dl, keep_mask=kb)),
                 ('ln bwd philox', lambda: k.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, 0.1, 1, 1, True, dg, db, dbias_lin=dl)),
                 ('ln bwd p=0', lambda: k.dropout_add_ln_bwd(dy, None, s, g, mean, rstd, 0.0, 1, 1, True, dg, db, dbias_lin=dl))]:
    for flush in (True, False):
        t = timeit(fn, flush=flush)
        print(f'{name:16s} flush={flush}: {t * 1e3:7.1f} us  {4 * M * D * 2 / t / 1e6:7.1f} GB/s')
