"""Child process of test_parity_gpu.py::test_pdl_switch: one forward + backward of the tiny ST model on the GPU with whatever
PB200_PDL says (the library reads the variable once per process); prints the loss and a digest of every gradient as JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.golden import synth  # noqa: E402
from tests.test_parity_gpu import BF16, build, to_cuda  # noqa: E402

cfg, task, model = build()
model.to(BF16).cuda().train()
loss, logs = model(**to_cuda(synth.model_batch()))
loss.backward()
torch.cuda.synchronize()
grads = {n: p.grad.float() for n, p in model.named_parameters() if p.grad is not None}
print(json.dumps({'loss': loss.item(), 'nll_loss': logs['nll_loss'],
                  'grads': {n: [g.abs().sum().item(), g.abs().max().item()] for n, g in grads.items()}}))
