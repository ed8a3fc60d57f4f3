"""Launched under torchrun with 2+ ranks on one box: checks that the data-parallel step (bucketed NCCL allreduce
overlapped with backward, and the CUDA-graph variant with the allreduce between graph segments) produces the same
update as a single process accumulating the ranks' batches as micro-batches (update_freq = world)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from tests.golden import synth
from tests.test_parity_gpu import build, by_build


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pasero_b200.training import Trainer
    batches = synth.train_batches()
    steps = 4
    # `adapters_by`: per-language adapters (encoder_adapters_by / decoder_adapters_by): the ranks of one step train different
    # language pairs, i.e. different adapter sets, and some sets are used by no rank at all
    by = 'adapters_by' in sys.argv[1:]
    if by:
        steps = 6
        batches = [synth.by_train_batch(i) for i in range(steps * world)]

        def build_fn(**kw):     # the base check's learning rate: the parameters barely move, so the last step's reduced
            return by_build(lr=synth.TRAIN['lr'], warmup=synth.TRAIN['warmup'], **kw)    # gradients stay comparable
    else:
        build_fn = build
    # `float16` as the first argument: the same check with dtype float16 + FlatFP16Adam (loss scaling, fp32 master weights)
    extra = dict(dtype='float16', memory_efficient_fp16=False, flat_fp16=True) if 'float16' in sys.argv[1:] else {}
    results = {}
    for mode in ('eager', 'graph'):
        cfg, task, model = build_fn(seed=1, bucket_bytes=(4 if by else 64) * 1024, **extra)
        trainer = Trainer(cfg, task, model)
        assert trainer.distributed and trainer.dp_size == world and len(trainer.buckets.buckets) > 3
        logs_all = []
        for s in range(steps):
            b = batches[(s * world + rank) % len(batches)]
            if mode == 'graph':
                logs = trainer.train_step_graphed(trainer.move_to_device(b))
            else:
                logs = trainer.train_step(iter([b]))
            logs_all.append(dict(logs.items()))
        results[mode] = (trainer.arena.param.clone(), logs_all, trainer.arena.grad.float().clone(), dict(trainer.group_steps))
    # reference: one process, the ranks' batches as micro-batches
    cfg, task, model = build_fn(seed=1, virtual_dp_size=world, **extra)
    ref = Trainer(cfg, task, model, data_parallel=False)
    ref.update_freq = world
    ref_logs = []
    for s in range(steps):
        mb = [batches[(s * world + r) % len(batches)] for r in range(world)]
        ref_logs.append(dict(ref.train_step(iter(mb)).items()))
    ok = True
    ref_grad = ref.arena.grad.float()
    for mode, (param, logs_all, grad, group_steps) in results.items():
        assert group_steps == ref.group_steps, (mode, group_steps, ref.group_steps)
        # the reduced gradients of the last step themselves (the parameters barely move at this learning rate)
        gdiff = ((grad - ref_grad).abs().max() / ref_grad.abs().max()).item()
        diff = (param.float() - ref.arena.param.float()).abs().max().item()
        frac = (param != ref.arena.param).float().mean().item()
        for a, b in zip(logs_all, ref_logs):
            assert a['num_tokens'] == b['num_tokens'] and a['num_lines'] == b['num_lines'], (a, b)
            assert abs(a['loss'] - b['loss']) / b['loss'] < 2e-3, (mode, a['loss'], b['loss'])
            assert abs(a['gnorm'] - b['gnorm']) / b['gnorm'] < 2e-2, (mode, a['gnorm'], b['gnorm'])
        if rank == 0:
            print(f'{mode}: max |param diff| vs single-process accumulation {diff:.3e}, differing elements {frac:.4f}, '
                  f'loss {logs_all[-1]["loss"]:.3f} vs {ref_logs[-1]["loss"]:.3f}, gnorm {logs_all[-1]["gnorm"]:.4f} vs {ref_logs[-1]["gnorm"]:.4f}')
        if rank == 0:
            print(f'{mode}: max |reduced gradient - accumulated gradient| / max |gradient| = {gdiff:.3e}')
        ok = ok and diff < 2e-2 and gdiff < 3e-2   # per-rank bf16 rounding before the sum: ~2e-2 of the largest entry, eager and graph alike
    # all ranks hold identical parameters
    p = results['eager'][0].float()
    p0 = p.clone()
    dist.broadcast(p0, 0)
    assert torch.equal(p, p0), 'ranks diverged'
    if rank == 0:
        print('DDP CHECK', 'float16' if extra else 'bfloat16', 'adapters_by ' + str(ref.group_steps) if by else '', 'OK' if ok else 'FAILED')
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
