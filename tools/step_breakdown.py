"""In-situ breakdown of the ST training step: the step's CUDA graph is re-captured with external-event brackets around
EVERY kernel wrapper (GEMM and non-GEMM), replayed, and the bracket times (minus the calibrated bracket overhead) are
summed per kernel family."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pasero_b200 import build_model, frontend, kernels as K
from pasero_b200.task import Dictionary, SpeechTranslationTask
from pasero_b200.training import Trainer

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
cfg = bench.st_config()
task = SpeechTranslationTask(cfg, Dictionary(tokens=[f't{i}' for i in range(bench.VOCAB - 4)]))
torch.manual_seed(1234)
model = build_model(cfg, task)
trainer = Trainer(cfg, task, model)
sub = model.encoder.subsample
pcm_host, dec_host = bench.synthetic_batch(0, bench.B_PER_GPU)
plan = frontend.FbankPlan([bench.N_SAMPLES] * bench.B_PER_GPU, dev)
F = plan.max_frames
rows, front = sub.padded_rows(F, 0)
pcm_dev = pcm_host.to(dev).reshape(-1).contiguous()
dec_dev = dec_host.to(dev)


def step():
    _, padded = frontend.fbank(pcm_dev, plan, cmvn=True, padded_frames=rows, front_pad=front)
    batch = {'encoder_input': padded, 'encoder_input_length': plan.frames_tensor, 'decoder_input': dec_dev,
             'padded_layout': (F,), 'meta': {}}
    return trainer.train_step_graphed(batch)


for _ in range(3):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    step()
e.record(); torch.cuda.synchronize()
print(f'un-instrumented step: {s.elapsed_time(e) / 10:.3f} ms')
overhead = bench.event_bracket_overhead(dev)
print(f'bracket overhead {overhead * 1e3:.2f} us')
trainer._graphs.clear()
K.PROFILE, K.PROFILE_TAGS[:], K.PROFILE_ALL, K.PROFILE_EXTERNAL = [], [], [], True
marks = []
orig_body = trainer._graph_body


def body(*a, **kw):
    marks.append(len(K.PROFILE_ALL))
    orig_body(*a, **kw)
    marks.append(len(K.PROFILE_ALL))


trainer._graph_body = body
step()
trainer._graph_body = orig_body
n_g = len(K.PROFILE) // 2
gem = list(zip(K.PROFILE_TAGS[-n_g:], K.PROFILE[-n_g:]))
# keep the brackets of the CAPTURED body (second call) and of the eager kernels around the graph (fbank, optimizer):
# the warm-up pass's events are never re-recorded by the replays
allk = K.PROFILE_ALL[:marks[0]] + K.PROFILE_ALL[marks[2]:]
K.PROFILE, K.PROFILE_ALL = None, None
for _ in range(3):
    step()
torch.cuda.synchronize()
# non-GEMM brackets: fbank (eager, once per call) + warm-up body + captured body + optimizer; keep those recorded by the replay
fam = collections.defaultdict(lambda: [0, 0.0])
for tag, (f, a, b) in gem:
    t = a.elapsed_time(b) - overhead
    k = 'gemm ' + ('wgrad' if 'a_mn1' in tag else ('dgrad' if 'b_mn1' in tag else 'fwd'))
    fam[k][0] += 1; fam[k][1] += t
seen = 0
for tag, a, b in allk:
    try:
        t = a.elapsed_time(b) - overhead
    except Exception:
        continue
    name = tag.split(' ')[0]
    fam[tag if name.startswith('attention') else name][0] += 1
    fam[tag if name.startswith('attention') else name][1] += t
tot = sum(v[1] for v in fam.values())
for k, (n, t) in sorted(fam.items(), key=lambda x: -x[1][1]):
    print(f'{k:46s} n={n:4d} {t * 1e3:9.1f} us  {100 * t / tot:5.1f}%   avg {t / n * 1e3:7.1f} us')
print(f'sum of brackets {tot:.3f} ms')
