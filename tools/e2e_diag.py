"""Where does the end-to-end step time go? (H2D bandwidth, prefetch overlap, host enqueue time)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pasero_b200 import build_model, frontend
from pasero_b200.task import Dictionary, SpeechTranslationTask
from pasero_b200.training import Trainer, DevicePrefetcher

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
cfg = bench.st_config()
task = SpeechTranslationTask(cfg, Dictionary(tokens=[f't{i}' for i in range(bench.VOCAB - 4)]))
torch.manual_seed(1234)
model = build_model(cfg, task)
trainer = Trainer(cfg, task, model)
sub = model.encoder.subsample
pcm_host, dec_host = bench.synthetic_batch(0, bench.B_PER_GPU)
pcm_host, dec_host = pcm_host.pin_memory(), dec_host.pin_memory()
plan = frontend.FbankPlan([bench.N_SAMPLES] * bench.B_PER_GPU, dev)
F = plan.max_frames
rows, front = sub.padded_rows(F, 0)


def step(pcm_flat, dec):
    _, padded = frontend.fbank(pcm_flat, plan, cmvn=True, padded_frames=rows, front_pad=front)
    batch = {'encoder_input': padded, 'encoder_input_length': plan.frames_tensor, 'decoder_input': dec,
             'padded_layout': (F,), 'meta': {}}
    return trainer.train_step_graphed(batch)


pcm_dev = pcm_host.to(dev).reshape(-1).contiguous()
dec_dev = dec_host.to(dev)
for _ in range(5):
    step(pcm_dev, dec_dev)
torch.cuda.synchronize()

# 1. H2D bandwidth
t0 = time.perf_counter()
for _ in range(10):
    x = pcm_host.to(dev, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f'H2D {pcm_host.numel() * 4 / 1e6:.1f} MB in {dt * 1e3:.2f} ms = {pcm_host.numel() * 4 / dt / 1e9:.1f} GB/s')

# 2. host enqueue time of one step (GPU idle at start)
t0 = time.perf_counter()
for _ in range(10):
    lg = step(pcm_dev, dec_dev)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue {1e3 * (t1 - t0) / 10:.2f} ms/step; wall {1e3 * (t2 - t0) / 10:.2f} ms/step (device-resident, no reads)')

# 3. device-resident + deferred read
t0 = time.perf_counter()
prev = None
for _ in range(20):
    lg = step(pcm_dev, dec_dev)
    if prev is not None:
        prev['loss']
    prev = lg
prev['loss']
torch.cuda.synchronize()
print(f'device-resident + deferred log read: {1e3 * (time.perf_counter() - t0) / 20:.2f} ms/step')

# 4. prefetcher
for name, n in (('prefetch (cold)', 10), ('prefetch (warm)', 20)):
    t0 = time.perf_counter()
    prev = None
    for pcm_d, dec_d in DevicePrefetcher(((pcm_host, dec_host) for _ in range(n)), dev):
        lg = step(pcm_d.reshape(-1), dec_d)
        if prev is not None:
            prev['loss']
        prev = lg
    prev['loss']
    torch.cuda.synchronize()
    print(f'{name}: {1e3 * (time.perf_counter() - t0) / n:.2f} ms/step')

# 5. same-stream H2D, immediate read (old e2e loop)
t0 = time.perf_counter()
for _ in range(10):
    pcm_d = pcm_host.to(dev, non_blocking=True).reshape(-1)
    dec_d = dec_host.to(dev, non_blocking=True)
    step(pcm_d, dec_d)['loss']
torch.cuda.synchronize()
print(f'same-stream copy + immediate read: {1e3 * (time.perf_counter() - t0) / 10:.2f} ms/step')
