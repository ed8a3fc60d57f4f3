import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden import synth
from tests.test_parity_gpu import build, to_cuda, rel_to_max, BF16
from tests.test_oracle_cpu import tiny_cfg
from oracle import model as om
cfg, task, model = build()
model.to(BF16).cuda().train()
batch = synth.model_batch()
cb = to_cuda(batch)
loss, logs = model(**cb)
model.zero_grad()
loss.backward()
sd = {k: v.clone().requires_grad_(True) for k, v in synth.tiny_state_dict().items()}
oloss, ologs = om.forward(tiny_cfg(), sd, batch['encoder_input'], batch['encoder_input_length'], batch['decoder_input'])
oloss.backward()
saved = [(p, p.data) for p in model.parameters()]
for p in model.parameters():
    p.data = p.grad
grads = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if k in sd}
for p, d in saved:
    p.data = d
for k, g in grads.items():
    r = rel_to_max(g, sd[k].grad)
    cos = torch.nn.functional.cosine_similarity(g.flatten(), sd[k].grad.flatten(), dim=0).item()
    print(f'{k:55s} rel {r:8.4f} cos {cos:8.5f} |ref| {sd[k].grad.abs().max().item():.3e} |got| {g.abs().max().item():.3e}')
