"""In-kernel timeline of the tcgen05 attention kernels (library built with PB200_NVCC_EXTRA=-DPB_ATTN_TIMING)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasero_b200 import kernels as k, _lib

B, H, Lq, Lk = 64, 8, 250, 250
if len(sys.argv) > 1:
    Lq, Lk = int(sys.argv[1]), int(sys.argv[2])
dev = 'cuda'
q = torch.randn(B * Lq, H * 64, device=dev).to(torch.bfloat16)
kv = torch.randn(B * Lk, 2 * H * 64, device=dev).to(torch.bfloat16)
kl = torch.full((B,), Lk, device=dev, dtype=torch.int32)
lib = _lib.load()
raw = ctypes.CDLL(os.path.join(os.path.dirname(_lib.__file__), 'libpasero_b200.so'))


def dump(n_ctas, names, title):
    buf = np.zeros(n_ctas * 16, dtype=np.int64)
    rc = raw.pb200_debug_attn_timing(buf.ctypes.data_as(ctypes.c_void_p), n_ctas)
    assert rc == 0
    t = buf.reshape(n_ctas, 16)
    base = t[:, 0:1]
    rel = t[:, :14] - base
    order = np.argsort(t[:, 14])
    print(f'== {title}: {n_ctas} CTAs; first start {t[:,14].min()} ns, last start +{(t[:,14].max()-t[:,14].min())/1e3:.1f} us')
    for sl, name in names.items():
        v = rel[:, sl]
        print(f'  slot {sl:2d} {name:28s} median {np.median(v):9.0f}  p10 {np.percentile(v,10):9.0f}  p90 {np.percentile(v,90):9.0f} cycles')
    # CTAs per SM and their start offsets
    sm = t[:, 15]
    first = t[order[:3]]
    for r in first:
        print('   first CTA rel:', (r[:14] - r[0]).tolist())
    starts = (t[:, 14] - t[:, 14].min()) / 1e3
    hist, _ = np.histogram(starts, bins=12)
    print('   CTA start histogram (us bins up to %.1f):' % starts.max(), hist.tolist())


for _ in range(3):
    o, lse = k.attention_fwd(q, kv[:, :H * 64], kv[:, H * 64:], B, H, Lq, Lk, kl, False, 0.125)
torch.cuda.synchronize()
n = min(4096, ((Lq + 127) // 128) * H * B)
dump(n, {1: 'mma: setup done', 2: 'mma: TMA landed', 3: 'mma: S issued', 4: 'mma: P ready', 5: 'mma: PV issued', 6: 'mma: dealloc',
         8: 'sm0: start', 9: 'sm0: S ready', 10: 'sm0: max done', 11: 'sm0: P written', 12: 'sm0: O ready', 13: 'sm0: done'}, 'forward')
do = torch.randn_like(o)
dq, dkv = torch.empty_like(q), torch.empty_like(kv)
for _ in range(3):
    k.attention_bwd(do, q, kv[:, :H * 64], kv[:, H * 64:], o, lse, B, H, Lq, Lk, kl, False, 0.125, dq, dkv[:, :H * 64], dkv[:, H * 64:])
torch.cuda.synchronize()
dump(H * B, {1: 'mma: TMA issued', 2: 'mma: TMA landed', 3: 'mma: S,dP issued (it0)', 4: 'mma: P,dS ready (it0)', 5: 'mma: dV,dK,dQ issued (it0)',
             6: 'mma: all issued', 7: 'mma: dealloc', 8: 'sm0: delta done', 9: 'sm0: S ready (it0)', 10: 'sm0: P written (it0)',
             11: 'sm0: S ready (it1)', 12: 'sm0: P written (it1)', 13: 'sm0: done'}, 'backward')
