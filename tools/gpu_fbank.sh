timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -k "fbank or in_linear" 2>&1 | tail -4
cat > /tmp/fb.py <<'P'
import os, sys
sys.path.insert(0, '.')
import torch
from pasero_b200 import kernels as k
from tools.microbench import timeit
B, N = 64, 160000
pcm = (torch.rand(B * N, device='cuda') - 0.5)
F = k.fbank_num_frames(N)
so = torch.arange(B + 1, device='cuda', dtype=torch.int64) * N
fo = torch.arange(B + 1, device='cuda', dtype=torch.int64) * F
nbytes = 4 * B * N + 320 * B * F
for name, fn in [('fbank+cmvn', lambda: k.fbank_cmvn(pcm, so, fo, B, F, B * F, cmvn=True, padded_frames=F + 4, front_pad=2)),
                 ('fbank raw', lambda: k.fbank_cmvn(pcm, so, fo, B, F, B * F, cmvn=False))]:
    t = timeit(fn)
    print(f'{os.environ.get("PB200_FBANK", "v2")} {name}: {t*1e3:.1f} us  {nbytes / t / 1e6:.1f} GB/s  {B * F / t / 1e6:.2f} G frames/s')
P
for v in v2 v1; do PB200_FBANK=$v timeout 300 python /tmp/fb.py; done
