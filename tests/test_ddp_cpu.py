"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucket partition of the gradient arena, the
ready-order driven launches, sum-allreduce semantics with the 1/num_tokens normalisation, and unused-parameter
flushing. The CUDA kernels are not involved: gradients are filled by hand."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pasero_b200.training import GradBuckets, ParamArena, adapter_group, groups_in_use


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(64, 96)
        self.b = torch.nn.Linear(96, 40)
        self.c = torch.nn.Linear(40, 8)
        self.unused = torch.nn.Parameter(torch.zeros(24))


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = Tiny().to(torch.bfloat16)
        arena = ParamArena(model, 'cpu')
        buckets = GradBuckets(arena, bucket_bytes=8 * 1024)
        # partition: contiguous, covers the arena, built back to front
        spans = sorted(buckets.buckets)
        assert spans[0][0] == 0 and spans[-1][1] == arena.numel
        assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
        assert buckets.buckets[0][1] == arena.numel and len(buckets.buckets) >= 3
        # every offset is 16-byte aligned
        assert all(o % 8 == 0 for o in arena.offsets)
        # fill rank-dependent grads, fire the hooks in backward (reverse) order, skip the unused parameter
        buckets.start_step(sync=True)
        for i in reversed(range(len(arena.params))):
            if arena.names[i] == 'unused':
                continue
            arena.view(arena.grad, i).fill_(float(rank + 1))
            buckets.param_ready(i)
        stats = torch.tensor([10.0 * (rank + 1), 5.0, 100.0 + rank, 4.0, 0, 0, float(rank == 1), 0], dtype=torch.float64)
        launched_before_finish = list(buckets.launch_order)
        buckets.finish(stats)
        expect = float(sum(r + 1 for r in range(world)))
        for i, name in enumerate(arena.names):
            g = arena.view(arena.grad, i).float()
            want = 0.0 if name == 'unused' else expect
            assert torch.all(g == want), (name, g.unique())
        assert stats[:4].tolist() == [10.0 * 3, 10.0, 201.0, 8.0]
        assert stats[6].item() == 1.0 and stats[7].item() == 0.0     # rank 1's failed micro-batch is known to every rank
        # buckets complete in reverse-layer order (the last bucket holds the last parameters)
        assert launched_before_finish == sorted(launched_before_finish)
        # graph mode: one allreduce over the whole arena between the step graph and the optimizer (allreduce_all)
        flushed = []
        buckets.pre_all = lambda: flushed.append('all')
        buckets.start_step(sync=False)
        for i in range(len(arena.params)):
            arena.view(arena.grad, i).fill_(0.0 if arena.names[i] == 'unused' else float(2 * rank + 1))
            buckets.param_ready(i)                       # disabled inside the graph: nothing is launched
        assert buckets.launch_order == []
        stats = torch.tensor([1.0 + rank, 2.0, 3.0, 4.0, 0, 0, 0, 0], dtype=torch.float64)
        buckets.allreduce_all(stats)
        expect = float(sum(2 * r + 1 for r in range(world)))
        for i, name in enumerate(arena.names):
            g = arena.view(arena.grad, i).float()
            assert torch.all(g == (0.0 if name == 'unused' else expect)), (name, g.unique())
        assert stats[:4].tolist() == [3.0, 4.0, 6.0, 8.0] and flushed == ['all']
        assert sorted(buckets.launch_order) == list(range(len(buckets.buckets)))
        buckets.pre_all = None
        # no_sync micro-batch: nothing is reduced
        buckets.start_step(sync=False)
        arena.grad.fill_(float(rank))
        for i in range(len(arena.params)):
            buckets.param_ready(i)
        assert buckets.launch_order == [] and torch.all(arena.grad.float() == rank)
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


class Adapter(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.down = torch.nn.Linear(32, 8)
        self.up = torch.nn.Linear(8, 32)


class Layer(torch.nn.Module):
    def __init__(self, names):
        super().__init__()
        self.fc = torch.nn.Linear(32, 32)
        self.adapters = torch.nn.ModuleDict({n: Adapter() for n in names})


class ByLang(torch.nn.Module):
    """Parameter names of an AdapterTransformer with `encoder_adapters_by: source_lang`, `decoder_adapters_by: target_lang`."""
    def __init__(self):
        super().__init__()
        self.encoder = torch.nn.Module()
        self.encoder.layers = torch.nn.ModuleList([Layer(['lang:de', 'lang:it']) for _ in range(2)])
        self.decoder = torch.nn.Module()
        self.decoder.layers = torch.nn.ModuleList([Layer(['lang:en', 'lang:fr']) for _ in range(2)])


def _worker_by(rank: int, world: int, port: int, q):
    """`*_adapters_by`: rank 0 trains a (de, en) batch, rank 1 an (it, en) batch. Buckets are issued in index order on every
    rank (their completion order during backward differs), gradients of adapters a rank did not use enter the sum as zeros,
    and the union of the groups in use is the same on both ranks."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = ByLang().to(torch.bfloat16)
        arena = ParamArena(model, 'cpu', group_key=adapter_group)
        keys = [g for g in arena.group_ranges if g is not None]
        assert keys == [('encoder', 'lang:de'), ('encoder', 'lang:it'), ('decoder', 'lang:en'), ('decoder', 'lang:fr')]
        # layout: the always-used parameters first, then one contiguous range per group, covering the arena
        spans = [arena.group_ranges[None]] + [arena.group_ranges[g] for g in keys]
        assert spans[0][0] == 0 and spans[-1][1] == arena.numel and all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
        for i, name in enumerate(arena.names):
            lo, hi = arena.group_ranges[adapter_group(name)]
            assert lo <= arena.offsets[i] < hi and arena.param_group[i] == adapter_group(name)
        buckets = GradBuckets(arena, bucket_bytes=2 * 1024)
        buckets.ordered = True
        assert len(buckets.buckets) >= 4
        mine = {('encoder', 'lang:de' if rank == 0 else 'lang:it'), ('decoder', 'lang:en')}
        arena.grad.zero_()          # what the Trainer does at the start of a step in this mode
        buckets.start_step(sync=True)
        order = reversed(range(len(arena.params))) if rank == 0 else range(len(arena.params))    # different completion orders
        for i in order:
            g = arena.param_group[i]
            if g is None or g in mine:
                arena.view(arena.grad, i).fill_(float(rank + 1))
                buckets.param_ready(i)
        assert buckets.launch_order == []            # nothing before finish()
        buckets.finish(None)
        assert buckets.launch_order == list(range(len(buckets.buckets)))
        want = {None: 3.0, ('encoder', 'lang:de'): 1.0, ('encoder', 'lang:it'): 2.0, ('decoder', 'lang:en'): 3.0, ('decoder', 'lang:fr'): 0.0}
        for i, name in enumerate(arena.names):
            g = arena.view(arena.grad, i).float()
            assert torch.all(g == want[arena.param_group[i]]), (name, g.unique())
        used = groups_in_use(keys, mine, True)
        assert used == [('encoder', 'lang:de'), ('encoder', 'lang:it'), ('decoder', 'lang:en')], used
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _run(worker):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: 'ok', 1: 'ok'}, results


def test_adapters_by_world2_gloo():
    _run(_worker_by)


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: 'ok', 1: 'ok'}, results
