import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tests.golden import synth
from tests.test_parity_gpu import build

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from pasero_b200.training import Trainer
batches = synth.train_batches()
cfg, task, model = build(seed=1, bucket_bytes=64 * 1024)
trainer = Trainer(cfg, task, model)
b = trainer.move_to_device(batches[rank])
b['meta'] = {}
# local grads (no sync)
trainer.stats.zero_(); trainer.direct.start_step(); trainer.buckets.start_step(sync=False)
trainer._forward_backward(dict(b)); torch.cuda.synchronize()
local_g = trainer.arena.grad.clone()
expect = local_g.clone(); dist.all_reduce(expect)
# DDP grads
trainer.stats.zero_(); trainer.direct.start_step(); trainer.buckets.start_step(sync=True)
trainer._forward_backward(dict(b)); trainer.buckets.finish(trainer.stats); torch.cuda.synchronize()
got = trainer.arena.grad.clone()
for i, name in enumerate(trainer.arena.names):
    g, e, l = trainer.arena.view(got, i).float(), trainer.arena.view(expect, i).float(), trainer.arena.view(local_g, i).float()
    if not torch.allclose(g, e, rtol=2e-2, atol=1e-3):
        print(f'rank {rank} {name}: got {g.flatten()[:3].tolist()} expect {e.flatten()[:3].tolist()} local {l.flatten()[:3].tolist()} '
              f'|got-local| {(g-l).abs().max().item():.3e} |got-expect| {(g-e).abs().max().item():.3e}', flush=True)
dist.destroy_process_group()
