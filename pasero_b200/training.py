"""The training step of the ST path: B200-native counterpart of pasero/training.py (Trainer.train_step :324-517,
checkpoint I/O :613-939) and pasero/optimization.py (Adam :56-149, LRScheduler :21-52, clip_grad_norm_ :390-427).

Design (B200-first, one process per GPU):
  * every parameter, gradient and Adam moment lives in ONE flat arena (bf16 params/grads, fp32 moments); the
    nn.Parameters are views, so the optimizer is two kernel launches (grad-norm, clip+Adam) instead of ~10 x 353;
  * data parallelism = bucketed bf16 allreduce (sum) of slices of the gradient arena over NCCL on a side stream,
    launched from post-accumulate hooks in reverse-layer order so it overlaps the rest of backward; the loss /
    token-count scalars ride in one fp64 allreduce; no all_gather_object, no per-step host synchronisation
    (the reference syncs >= 5 times per micro-batch);
  * the whole step (forward, backward, clip, Adam) can be captured in a CUDA graph per batch shape; step scalars
    (lr, bias corrections, dropout offsets) live in device memory and are refreshed before each replay.
"""
import math
import os
import re
import time
import traceback
from collections import defaultdict
from collections.abc import Mapping
from typing import Iterator, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import kernels as K
from . import ops
from .config import TrainingConfig

BF16 = torch.bfloat16
ALIGN = 8  # elements: every parameter starts on a 16-byte boundary (TMA / 128-bit loads)


class LRScheduler:
    """pasero/optimization.py:21-52: linear warmup init_lr -> lr, then lr * sqrt(warmup / step); warmup == 0 means
    linear decay to min_lr over max_steps. `last_epoch` counts completed scheduler steps, like torch's."""

    def __init__(self, cfg: TrainingConfig):
        assert cfg.init_lr >= 0 and cfg.warmup >= 0 and cfg.init_lr <= cfg.lr and cfg.max_steps > 0
        self.warmup, self.max_steps = cfg.warmup, cfg.max_steps
        self.max_lr, self.min_lr, self.init_lr = cfg.lr, cfg.min_lr, cfg.init_lr
        self.last_epoch = 0

    def lr_at(self, step: int) -> float:
        if self.warmup == 0:
            lr = self.max_lr - step * (self.max_lr - self.min_lr) / self.max_steps
        elif step < self.warmup:
            lr = self.init_lr + step * (self.max_lr - self.init_lr) / self.warmup
        else:
            lr = self.max_lr * math.sqrt(self.warmup / step)
        return max(lr, self.min_lr)

    def get_last_lr(self) -> list:
        return [self.lr_at(self.last_epoch)]

    def step(self) -> None:
        self.last_epoch += 1

    def state_dict(self) -> dict:
        return {'last_epoch': self.last_epoch, 'warmup': self.warmup, 'max_steps': self.max_steps, 'max_lr': self.max_lr,
                'min_lr': self.min_lr, 'init_lr': self.init_lr, '_step_count': self.last_epoch + 1,
                '_last_lr': self.get_last_lr()}

    def load_state_dict(self, state: dict) -> None:
        self.last_epoch = state['last_epoch']


_ADAPTER_NAME = re.compile(r'^(encoder|decoder)\.layers\.\d+\.adapters\.([^.]+)\.')


def adapter_group(name: str):
    """('encoder' | 'decoder', adapter name) for an adapter parameter, None for every other parameter: the unit that is used
    or unused as a whole in a step under `*_adapters_by` (pasero/models/adapters.py:186-197, 219-229 select one adapter name
    per side for all layers)."""
    m = _ADAPTER_NAME.match(name)
    return (m.group(1), m.group(2)) if m else None


def groups_in_use(keys: list, local, distributed: bool, host_group=None) -> list:
    """The adapter groups that received a gradient on ANY rank in this step, in arena order (`local`: this rank's). The
    union is formed on the host (CPU tensor over gloo): the GPU is not synchronised and every rank gets the same answer,
    which DDP's find_unused_parameters obtains from an allreduced usage bitmap (pasero/training.py:236-250)."""
    flags = torch.tensor([int(g in local) for g in keys], dtype=torch.int32)
    if distributed and len(keys) > 0:
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=host_group)
    return [g for g, f in zip(keys, flags.tolist()) if f]


class PinnedRing:
    """Pinned host staging buffers for small per-step H2D copies. An asynchronous copy reads pinned memory when it EXECUTES,
    not when it is enqueued, and the host runs several steps ahead of the GPU (deferred logs): a single staging buffer would
    hand step t the scalars of step t + k. Each slot is reused only after the copy enqueued from it has completed."""

    def __init__(self, shape, dtype, slots: int = 8):
        self.slots = [torch.zeros(shape, dtype=dtype).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.i = 0

    def acquire(self) -> Tensor:
        if self.events[self.i] is not None:
            self.events[self.i].synchronize()
        return self.slots[self.i]

    def release(self) -> None:
        """Call after the copies out of the acquired slot were enqueued on the current stream."""
        event = torch.cuda.Event()
        event.record()
        self.events[self.i] = event
        self.i = (self.i + 1) % len(self.slots)


class ParamArena:
    """Flat storage for parameters and gradients (the model's 16-bit dtype), Adam moments (fp32) and — `master=True`: float16
    training with MixedPrecisionAdam / FlatFP16Adam semantics — an fp32 master copy of the parameters; parameters become views."""

    def __init__(self, model: torch.nn.Module, device, dtype=BF16, master: bool = False, group_key=None):
        """`group_key(name)` (optional): a hashable group for parameters that are only used in some steps (per-language /
        per-domain adapters), None for the always-used rest. Groups are laid out contiguously — the always-used parameters
        first, then one range per group (`group_ranges`) — so that the optimizer can step or skip a whole group with one launch."""
        self.params, self.names, self.offsets = [], [], []
        seen = set()
        entries = []
        for name, p in model.named_parameters():
            if not p.requires_grad or id(p) in seen:
                continue
            seen.add(id(p))
            entries.append((name, p, group_key(name) if group_key is not None else None))
        order = {None: 0}
        for _, _, g in entries:
            order.setdefault(g, len(order))
        entries.sort(key=lambda e: order[e[2]])       # stable: model order inside a group
        off = 0
        self.group_ranges = {}
        for name, p, g in entries:
            self.params.append(p)
            self.names.append(name)
            self.offsets.append(off)
            lo, _ = self.group_ranges.get(g, (off, off))
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            self.group_ranges[g] = (lo, off)
        self.param_group = [g for _, _, g in entries]
        self.numel = off
        self.param = torch.zeros(off, dtype=dtype, device=device)
        self.grad = torch.zeros(off, dtype=dtype, device=device)
        # fp32 main gradients of the GEMM weights (>= 2-D parameters): the weight-gradient GEMMs reduce-add into it, it is
        # rounded into `grad` (and zeroed) once per optimizer step, before the allreduce / the optimizer read `grad`
        self.grad32 = torch.zeros(off, dtype=torch.float32, device=device)
        # ... and of the biases flagged `_pb200_f32_grad` (in-projection, fc1, conv): their gradient is a by-product of the
        # weight-gradient GEMM (pb200_gemm_params.a_rowsum), added atomically in fp32
        self.is32 = [(p.ndim >= 2 and not getattr(p, '_pb200_bf16_grad', False)) or getattr(p, '_pb200_f32_grad', False)
                     for p in self.params]
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.param[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        # fp32 master parameters (optimization.py:183-191, 222-233): built from the 16-bit weights, updated by Adam, copied back
        self.master = self.param.float() if master else None

    def view(self, flat: Tensor, i: int) -> Tensor:
        p, o = self.params[i], self.offsets[i]
        return flat[o:o + p.numel()].view(p.shape)

    def ranges32(self, lo: int, hi: int) -> tuple:
        """(ranges, max_len): merged (begin, end) element ranges of the fp32-accumulated parameters inside arena slice
        [lo, hi) as an int64 [n, 2] device tensor, and the longest range (host int, sizes the flush grid)."""
        out = []
        for i, p in enumerate(self.params):
            o = self.offsets[i]
            if not self.is32[i] or o < lo or o >= hi:
                continue
            end = o + (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            if out and out[-1][1] == o:
                out[-1][1] = end
            else:
                out.append([o, end])
        max_len = max((e - b for b, e in out), default=0)
        return torch.tensor(out, dtype=torch.int64, device=self.grad.device).reshape(-1, 2), max_len

    def flush32(self, ranges: tuple) -> None:
        """grad (bf16) <- round(grad32) over `ranges`; grad32 <- 0."""
        ops.WgradStream.join()   # the weight-gradient GEMMs forked onto the side stream have to land first
        if ranges[1] > 0:
            K.grad32_flush(self.grad32, self.grad, ranges[0], ranges[1])

    def flush32_behind(self, ranges: tuple):
        """The same for one gradient bucket in the middle of backward, on the weight-gradient stream: ordered behind the
        bucket's GEMMs (already queued there) and behind the current stream's work so far, while the current stream carries on
        with the rest of backward. Returns the stream on which the bucket's 16-bit gradients are final."""
        if not self.grad.is_cuda:
            return None
        return ops.WgradStream.run_behind(lambda: K.grad32_flush(self.grad32, self.grad, ranges[0], ranges[1]) if ranges[1] > 0 else None)

    def rebind_grads(self) -> None:
        for i, p in enumerate(self.params):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + self.offsets[i] * 2:
                p.grad = self.view(self.grad, i)


# Step scalars (`stats`, fp64[8] on the device): [0] loss [1] nll [2] target tokens [3] lines [4] squared gradient norm (written by
# the norm kernel after the reduction) [5] found_inf (ditto) [6] failed micro-batches (data-loading error / CUDA OOM on some rank)
# [7] target ids outside the vocabulary. Slots [0, STATS_SYNCED) are summed over the data-parallel ranks right after the forward
# pass ([4] and [5] are still zero then): the logs, and the job's status (pasero/training.py:431: utils.gather_dict).
STATS_SYNCED = 7


class GradBuckets:
    """Bucketed allreduce of the gradient arena, overlapped with backward (replaces torch DDP's reducer,
    pasero/training.py:243-250). Buckets are contiguous arena slices of ~bucket_bytes, filled back to front."""

    def __init__(self, arena: ParamArena, bucket_bytes: int, group=None):
        self.arena, self.group = arena, group
        self.cuda = arena.grad.is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None   # CPU tensors (gloo tests): same logic, no streams
        self.buckets = []           # (start, end) element ranges, in the order they become ready
        self.param_bucket = {}
        n = len(arena.params)
        end_i = n
        while end_i > 0:
            start_i = end_i
            size = 0
            while start_i > 0 and (size == 0 or size + arena.params[start_i - 1].numel() * 2 <= bucket_bytes):
                start_i -= 1
                size += arena.params[start_i].numel() * 2
            lo = arena.offsets[start_i]
            hi = arena.offsets[end_i] if end_i < n else arena.numel
            b = len(self.buckets)
            self.buckets.append((lo, hi))
            for i in range(start_i, end_i):
                self.param_bucket[i] = b
            end_i = start_i
        self.pending = [0] * len(self.buckets)
        self.sizes = [0] * len(self.buckets)
        for i, b in self.param_bucket.items():
            self.sizes[b] += 1
        self.enabled = True
        self.launch_order = []
        self.pre_launch = None   # callable(bucket index) run on the compute stream before the bucket's allreduce
        self.pre_all = None      # callable() doing the same for the whole arena at once (allreduce_all)
        # list while the step's CUDA graph is being captured in "overlap" mode: instead of reducing a complete bucket, the
        # capture records an EXTERNAL event (an event-record node of the graph) at the point where the bucket's gradients
        # are final; after every replay the NCCL allreduces are enqueued on the side stream behind those events
        # (`reduce_behind_events`), so they run next to the rest of the replayed backward
        self.capture_events = None
        self.use_autograd_hooks = True   # set False when DirectGrads reports readiness itself (the hooks still fire)
        # True when some parameters are unused in some steps (`*_adapters_by`): which buckets complete during backward then
        # differs from rank to rank, and collectives must be issued in ONE order everywhere — every bucket waits for finish(),
        # which launches them by index (torch DDP's find_unused_parameters answers the same question with an extra
        # allreduce of a usage bitmap, pasero/training.py:236-250)
        self.ordered = False
        for i, p in enumerate(arena.params):
            p.register_post_accumulate_grad_hook(lambda param, i=i: self.param_ready(i) if self.use_autograd_hooks else None)

    def param_ready(self, i: int) -> None:
        """Parameter i's gradient for this micro-batch is complete; launch its bucket when it is the last one."""
        if not self.enabled or self.ordered:
            return
        b = self.param_bucket[i]
        if self.pending[b] < 0:
            return
        self.pending[b] += 1
        if self.pending[b] == self.sizes[b]:
            self._launch(b)

    def _launch(self, b: int) -> None:
        lo, hi = self.buckets[b]
        # pre_launch may finish the bucket on another stream (the fp32 -> 16-bit flush behind the weight-gradient GEMMs): that
        # stream, not the current one, is what the allreduce has to wait for
        final = self.pre_launch(b) if self.pre_launch is not None else None
        if final is None and self.cuda:
            final = torch.cuda.current_stream()
        if self.capture_events is not None:
            event = torch.cuda.Event(external=True)
            event.record(final)
            self.capture_events.append((b, event))
        elif self.cuda:
            self.stream.wait_stream(final)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        self.pending[b] = -1
        self.launch_order.append(b)

    def start_step(self, sync: bool) -> None:
        self.enabled = sync
        self.pending = [0] * len(self.buckets)
        self.launch_order = []

    def allreduce_all(self, stats: Optional[Tensor] = None) -> None:
        """Non-overlapped variant used between the step's CUDA graph and the optimizer kernels. Nothing runs next to these
        reductions, so buckets buy nothing here: ONE allreduce over the whole arena (a single large message, one launch)
        and the step scalars behind it."""
        if self.pre_all is not None:
            self.pre_all()
        elif self.pre_launch is not None:
            finals = {self.pre_launch(b) for b in range(len(self.buckets))} - {None}
            for s in finals:
                torch.cuda.current_stream().wait_stream(s)
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.arena.grad, op=dist.ReduceOp.SUM, group=self.group)
                if stats is not None:
                    dist.all_reduce(stats[:STATS_SYNCED], op=dist.ReduceOp.SUM, group=self.group)
            torch.cuda.current_stream().wait_stream(self.stream)
        else:
            dist.all_reduce(self.arena.grad, op=dist.ReduceOp.SUM, group=self.group)
            if stats is not None:
                dist.all_reduce(stats[:STATS_SYNCED], op=dist.ReduceOp.SUM, group=self.group)
        self.pending = [-1] * len(self.buckets)
        self.launch_order = list(range(len(self.buckets)))

    def reduce_behind_events(self, events: list, stats: Optional[Tensor], stats_event) -> None:
        """After a replay of a graph captured with `capture_events`: the step scalars (final after the forward pass) and then
        every bucket, in capture order (identical on all ranks), each allreduce waiting for its bucket's event node of THIS
        replay; the current stream then waits for the side stream (before the optimizer kernels read the gradients)."""
        with torch.cuda.stream(self.stream):
            if stats is not None:
                self.stream.wait_event(stats_event)
                dist.all_reduce(stats[:STATS_SYNCED], op=dist.ReduceOp.SUM, group=self.group)
            for b, event in events:
                lo, hi = self.buckets[b]
                self.stream.wait_event(event)
                dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.current_stream().wait_stream(self.stream)

    def finish(self, stats: Optional[Tensor] = None) -> None:
        """Flush buckets whose hooks never fired (unused parameters), reduce the step scalars, rejoin the stream."""
        for b in range(len(self.buckets)):
            if self.pending[b] != -1:
                self._launch(b)
        if self.capture_events is not None:
            return      # nothing was launched on the side stream: there is nothing to rejoin inside the capture
        if stats is not None:
            if self.cuda:
                self.stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    dist.all_reduce(stats[:STATS_SYNCED], op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(stats[:STATS_SYNCED], op=dist.ReduceOp.SUM, group=self.group)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)


class Trainer:
    def __init__(self, cfg: TrainingConfig, task, model, metrics=None, status_mp=None, device: Optional[str] = None,
                 data_parallel: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError('pasero_b200.Trainer needs a CUDA device: the sm_100a kernels have no CPU fallback')
        cfg = TrainingConfig.adopt(cfg)      # ours, or the reference's own TrainingConfig (cli/train.py:287)
        self.cfg, self.task, self.model = cfg, task, model
        self.device = torch.device(device or f'cuda:{torch.cuda.current_device()}')
        self.dtype = torch.float16 if cfg.dtype == 'float16' else BF16
        # float16 (pasero/training.py:289-306): dynamic loss scaling (GradScaler: init 2**7, x2 every growth_interval clean
        # steps, / 2 and a skipped update on an inf / nan gradient); Adam on the 16-bit parameters (memory_efficient_fp16,
        # the reference's default) or on an fp32 master copy (MixedPrecisionAdam / FlatFP16Adam)
        self.fp16 = self.dtype == torch.float16
        self.master_weights = self.fp16 and (cfg.flat_fp16 or not cfg.memory_efficient_fp16)
        self.loss_scale = 2.0 ** 7 if self.fp16 else 1.0
        self._growth_tracker = 0
        self.overflows = 0
        self.distributed = data_parallel and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.dp_size = dist.get_world_size() if self.distributed else 1
        cfg.dp_size = self.dp_size
        self.update_freq = max(1, cfg.virtual_dp_size // self.dp_size)
        self.growth_interval = 2 ** 14 // self.update_freq
        # `*_adapters_by` (pasero/models/adapters.py:74-105, 186-229): which adapters run depends on the batch's language /
        # domain, so their parameters are unused in some steps. The reference's Adam keeps a step counter PER PARAMETER and
        # skips parameters without a gradient (optimization.py:72-74, 112), DDP (find_unused_parameters) leaves globally
        # unused gradients alone and clip_grad_norm_ only sees existing gradients. Here: the arena lays every adapter set
        # ("group": side + adapter name) out contiguously, the 16-bit gradient arena is zeroed at the start of the step,
        # every rank issues the gradient buckets in index order, the union over the ranks of the groups that received a
        # gradient is formed on the HOST (a gloo allreduce of a bitmap: no GPU synchronisation) and Adam is launched per
        # used group with that group's own step count / bias corrections; unused groups are not touched at all.
        self.grouped = bool(getattr(model, 'find_unused_parameters', False))
        model.to(self.dtype)
        model.to(self.device)
        from . import reference_param_names
        ref_names = reference_param_names(model)     # before the arena: afterwards every parameter shares one storage
        self.arena = ParamArena(model, self.device, self.dtype, master=self.master_weights,
                                group_key=adapter_group if self.grouped else None)
        self.group_keys = [g for g in self.arena.group_ranges if g is not None]
        self.group_steps = {g: 0 for g in self.group_keys}      # Adam's state['step'] of each group's parameters
        self._group_of = {id(p): g for p, g in zip(self.arena.params, self.arena.param_group)}
        self._stepped_groups = ()
        self.host_group = None
        if self.grouped and self.distributed and dist.get_backend() != 'gloo':
            self.host_group = dist.new_group(backend='gloo')    # host-side bitmap of the groups in use, see above
        self.param_names = list(self.arena.names)
        # the reference's names of the TRAINABLE parameters: optimizer state is keyed by them (pasero/training.py:899-906)
        self.reference_names = [r for n in self.arena.names for r in ref_names[n]]
        self._reference_group = {r: g for n, g in zip(self.arena.names, self.arena.param_group) for r in ref_names[n]}
        if self.distributed:  # identical initial weights on every rank (DDP broadcasts rank 0's at construction)
            dist.broadcast(self.arena.param, src=0)
        # a communicator of its own for the gradient buckets, on HIGH-PRIORITY streams: the step's persistent GEMM kernels follow
        # each other without a gap (programmatic dependent launch), and a normal-priority NCCL kernel is only scheduled once
        # the whole backward has drained — which is exactly the overlap that is wanted (measured at N = 8: 0.01 ms gained
        # without, see profiles/README.md). NCCL's own CTA count is left alone: capping it (ncclConfig max_ctas 2 / 4 / 8) made the
        # N = 8 step slower (10.12 / 9.44 / 9.32 vs 9.19 ms, profiles/r02_n8_sweep.txt) — the exposed part is the LAST bucket's
        # allreduce, which only gets longer with fewer CTAs.
        group = None
        if self.distributed and os.environ.get('PB200_NCCL_PRIORITY', '1') != '0':
            try:
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                group = dist.new_group(backend='nccl', pg_options=opts)
            except Exception:   # noqa: BLE001  (another backend / an old torch): the default group does the job
                group = None
        self.buckets = GradBuckets(self.arena, cfg.bucket_bytes, group) if self.distributed else None
        if self.buckets is not None:
            self.buckets.ordered = self.grouped
        # parameter gradients go straight into the arena (no AccumulateGrad adds, no per-step memset)
        index = {id(p): i for i, p in enumerate(self.arena.params)}
        on_ready = (lambda p: self.buckets.param_ready(index[id(p)])) if self.buckets is not None else None
        self.direct = ops.DirectGrads({id(p): self.arena.view(self.arena.grad, i) for i, p in enumerate(self.arena.params)},
                                      on_ready,
                                      {id(p): self.arena.view(self.arena.grad32, i) for i, p in enumerate(self.arena.params)
                                       if self.arena.is32[i]}, params=self.arena.params)
        ops.DirectGrads.active = self.direct
        self._ranges32_all = self.arena.ranges32(0, self.arena.numel)
        if self.buckets is not None:
            self.buckets.use_autograd_hooks = False
            bucket_ranges = [self.arena.ranges32(lo, hi) for lo, hi in self.buckets.buckets]
            self.buckets.pre_launch = lambda b: self.arena.flush32_behind(bucket_ranges[b])
            self.buckets.pre_all = lambda: self.arena.flush32(self._ranges32_all)
        self.scheduler = LRScheduler(cfg)
        self.optimizer_steps = 0     # Adam's state['step']
        self.steps = 0
        self.metrics = metrics if metrics is not None else defaultdict(list)
        # row 0: the step scalars every kernel reads; rows 1..: the same with each adapter group's own Adam step count
        self._hyper_all = torch.zeros(1 + len(self.group_keys), 16, dtype=torch.float32, device=self.device)
        self.hyper = self._hyper_all[0]
        self._hyper_ring = PinnedRing((1 + len(self.group_keys), 16), torch.float32)
        self.stats = model.stats_buffer(self.device)
        # ring of pinned host buffers: the logs of a step can be read after later steps were enqueued (pipelined loops)
        self._stats_ring = [torch.zeros(8, dtype=torch.float64).pin_memory() for _ in range(4)]
        self.rng_base = torch.zeros(1, dtype=torch.int64, device=self.device)
        model.defer_logs = True
        ops.DropoutState.reseed(cfg.seed if cfg.seed is not None else int(time.time()))
        self._graphs = {}
        # Data parallel + CUDA graph: NCCL calls are not captured (capturing them hangs on this stack: NCCL 2.28.9, torch 2.11);
        # the step graph carries one EXTERNAL event-record node per gradient bucket instead and the allreduces are enqueued
        # on the side stream behind those events after every replay, so they run next to the rest of the replayed backward.
        # PB200_GRAPH_OVERLAP=0: one whole-arena allreduce between the step graph and the optimizer (the round-1 behaviour).
        self.overlap_graph_allreduce = os.environ.get('PB200_GRAPH_OVERLAP', '1') != '0'
        self.saved = True

    # ------------------------------------------------------------------ step pieces
    def _hyper_row(self, h: Tensor, t: int) -> None:
        """The step scalars of csrc/optim.cu for Adam step count `t`, computed in fp64 like the reference's Python floats."""
        cfg = self.cfg
        lr = self.scheduler.get_last_lr()[0]
        b1, b2 = cfg.adam_betas
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        h[0], h[1], h[2], h[3], h[4] = lr, b1, b2, cfg.adam_eps, cfg.weight_decay
        h[5], h[6], h[7], h[8] = bc1, math.sqrt(bc2), cfg.clip_norm, 1.0
        h[9], h[10], h[11], h[12] = 1 - b1, 1 - b2, lr / bc1, cfg.weight_decay * lr
        h[13], h[14] = 1.0 / self.loss_scale, float(self.master_weights)

    def _write_hyper(self) -> None:
        h = self._hyper_ring.acquire()
        self._hyper_row(h[0], self.optimizer_steps + 1)
        self.hyper.copy_(h[0], non_blocking=True)
        self._hyper_ring.release()

    def move_to_device(self, batch: dict) -> dict:
        """utils.move_to_device(batch, 'cuda', dtype) (pasero/utils.py:691-702): float tensors are cast to bf16."""
        out = {}
        for k, v in batch.items():
            if torch.is_tensor(v):
                v = v.to(self.device, non_blocking=True)
                if v.is_floating_point():
                    v = v.to(self.dtype)
            out[k] = v
        return out

    def _forward_backward(self, batch: dict, after_forward=None):
        # float16: `scaler.scale(loss).backward()` (training.py:399-408) — the scale enters where the gradient is born, in the
        # loss kernel that writes d loss / d logits, so every downstream gradient is the scaled one
        ops.ProjCEFn.loss_scale = self.loss_scale
        loss, logs = self.model(**batch)
        if after_forward is not None:
            after_forward()
        ops.ProjCEFn.unit_loss_grad = True      # the loss is the root of this backward (grad_output == 1): scoped to it
        try:
            loss.backward()
        finally:
            ops.ProjCEFn.unit_loss_grad = False
            ops.ProjCEFn.loss_scale = 1.0
        return logs

    def _groups_in_use(self, local) -> list:
        return groups_in_use(self.group_keys, local, self.distributed, self.host_group)

    def _optimizer_kernels(self, local_groups=None) -> None:
        K.grad_sqnorm(self.arena.grad, self.hyper, self.stats)
        a = self.arena
        if not self.grouped:
            K.adam_step(a.param, a.grad, a.exp_avg, a.exp_avg_sq, self.hyper, self.stats, master=a.master)
            return
        # per-group Adam (see __init__): the always-used parameters with the global step count, then every adapter group
        # some rank used, with its own; a group nobody used keeps parameters, moments and step count as they are
        if local_groups is None:
            local_groups = {self._group_of[i] for i in self.direct.written if i in self._group_of}
        used = self._groups_in_use(local_groups)
        launches = []
        base = a.group_ranges.get(None)
        if base is not None and base[1] > base[0]:
            launches.append((0, base))
        if used:
            h = self._hyper_ring.acquire()
            for g in used:
                row = 1 + self.group_keys.index(g)
                self._hyper_row(h[row], self.group_steps[g] + 1)
                launches.append((row, a.group_ranges[g]))
            self._hyper_all[1:].copy_(h[1:], non_blocking=True)
            self._hyper_ring.release()
        for row, (lo, hi) in launches:
            K.adam_step(a.param[lo:hi], a.grad[lo:hi], a.exp_avg[lo:hi], a.exp_avg_sq[lo:hi], self._hyper_all[row], self.stats,
                        master=a.master[lo:hi] if a.master is not None else None)
        self._stepped_groups = tuple(used)

    # ------------------------------------------------------------------ the step
    def train_step(self, batch_iterator: Iterator[dict]) -> dict:
        """One optimizer update over `update_freq` micro-batches (pasero/training.py:324-517). Returns the step's
        logs: loss / nll_loss in bits (summed over all ranks' tokens), num_tokens, num_lines, gnorm, lr.

        Failure handling (training.py:339-386, 411-421): a micro-batch whose loading raises, or whose forward / backward runs
        out of GPU memory, is replaced by nothing — this rank's gradients of the step are zeroed, so it still takes part in
        every collective with a neutral contribution — and counted in stats[6], which is summed over the ranks with the other
        step scalars. A step with a failure anywhere updates no parameter on any rank (the norm kernel raises the same flag
        the loss scaler's overflow uses) and its logs carry 'failed': True: the caller stops training, as the reference's
        main loop does on `status.failed()`."""
        self.model.train()
        self.stats.zero_()
        if self.grouped:
            self.arena.grad.zero_()      # gradients of the adapters this step does not use must read as zero (see __init__)
        ops.DirectGrads.active = self.direct
        self.direct.start_step()
        failed = 0
        for batch_id in range(self.update_freq):
            batch = None
            try:
                batch = next(batch_iterator)
            except StopIteration:
                raise
            except Exception:   # noqa: BLE001  (a data-loading error must not desynchronise the ranks)
                traceback.print_exc()
            last = batch_id == self.update_freq - 1
            if self.buckets is not None:
                self.buckets.start_step(sync=last and batch is not None)   # no_sync() on all but the last micro-batch (training.py:407)
            if batch is None:
                failed += 1
                continue
            try:
                batch = self.move_to_device(batch)
                batch.setdefault('meta', {})
                batch['meta']['step'] = self.steps
                self._forward_backward(batch)
            except RuntimeError as e:
                if 'out of memory' not in str(e):
                    raise
                traceback.print_exc()
                torch.cuda.empty_cache()
                failed += 1
                if self.buckets is not None:
                    self.buckets.start_step(sync=False)     # whatever was launched stays launched; the rest goes through finish()
            ops.WgradStream.join()   # per micro-batch: the forked GEMMs' operands are kept referenced until the join
        if failed:
            self.arena.grad32.zero_()
            self.arena.grad.zero_()
            self.stats[:4] = 0
            self.stats[6] = failed
            # (buckets of this step that were already launched carry partial gradients into their allreduce: every rank sees
            # the failure flag and skips the update, so their content no longer matters — only the order of the collectives)
        if self.buckets is not None:
            self.buckets.finish(self.stats)
            ops.WgradStream.join()      # (finish() may have flushed its buckets on the weight-gradient stream)
        else:
            self.arena.flush32(self._ranges32_all)
        self._write_hyper()
        self._optimizer_kernels()
        return self._finish_step()

    @torch.no_grad()
    def valid_step(self, batch: Optional[dict]) -> dict:
        """Teacher-forced validation loss of one batch (pasero/training.py:519-547): eval mode (no dropout), no
        gradients, logs summed over the data-parallel ranks. `batch=None` (a rank that ran out of validation batches,
        datasets.py:551-555) contributes nothing but still takes part in the reduction."""
        vals = torch.zeros(4, dtype=torch.float64, device=self.device)
        if batch is not None:
            batch = self.move_to_device(batch)
            batch.setdefault('meta', {})
            was_training = self.model.training
            self.model.eval()
            try:
                _, logs = self.model(**batch)
            finally:
                self.model.train(was_training)
            vals[0], vals[1], vals[2], vals[3] = logs['loss'], logs['nll_loss'], logs['num_tokens'], logs['num_lines']
        if self.distributed:
            dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        v = vals.tolist()
        if batch is None and not self.distributed:
            return {}
        return {'loss': v[0], 'nll_loss': v[1], 'num_tokens': int(v[2]), 'num_lines': int(v[3])}

    @torch.no_grad()
    def inference_step(self, batch: Optional[dict], **decoding_opts) -> list:
        """Decodes one validation batch between training steps (pasero/training.py:549-611): the teacher-forcing
        `decoder_input` is cut back to its prompt (BOS + target tags, `prompt_length`), the encoder runs once and
        pasero_b200.decoding.search (beam search over the incremental CUDA decoder) produces the hypotheses; the best one per
        sentence is returned as a dict with 'tokens' (LongTensor on the CPU), 'score', 'normalized_score' and 'idx' (the
        sample's index in its corpus), plus whatever `task.postprocess` adds when the task has one (detokenisation is the
        reference task's business: this package's task stops at token ids). Data parallel: every rank decodes its own
        batch (`batch=None`: nothing to decode) and the lists are gathered on all ranks, like utils.gather_list."""
        from . import decoding
        hyps = []
        if batch is not None:
            batch = self.move_to_device(dict(batch))
            meta = dict(batch.pop('meta', None) or {})
            dec = batch['decoder_input']
            if 'prompt_length' in batch:       # utils.mask_by_len(..., truncate=True): keep the prompt, pad the rest, cut
                plen = batch['prompt_length'].to(dec.device)
                keep = torch.arange(dec.size(1), device=dec.device).unsqueeze(0) < plen.unsqueeze(1)
                dec = dec.masked_fill(~keep, self.model.padding_idx)[:, :int(plen.max())]
            else:
                dec = dec[:, :1]
            was_training = self.model.training
            self.model.eval()
            try:
                enc_kwargs = {k: v for k, v in batch.items() if k in ('encoder_input', 'encoder_input_length', 'padded_layout')}
                encoder_out, encoder_mask, _ = self.model.encoder(**enc_kwargs, meta=meta)
                opts = dict(max_output_len=decoding_opts.pop('max_output_len', self.cfg.max_target_len),
                            beam_size=decoding_opts.pop('beam_size', 5))
                opts.update(decoding_opts)
                output = decoding.search(self.model.decoder, encoder_out, encoder_mask, decoder_input=dec, meta=meta, **opts)
            finally:
                self.model.train(was_training)
            indices = batch.get('indices')
            for i, nbest in enumerate(output):
                hyp = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in nbest[0].items()}
                hyp['score'], hyp['normalized_score'] = float(hyp['score']), float(hyp['normalized_score'])
                if hasattr(self.task, 'postprocess'):
                    sample_bin = {k: v[i] for k, v in batch.items() if torch.is_tensor(v)}
                    sample_bin['meta'] = meta
                    self.task.postprocess(sample_bin, hyp, detokenize=True)
                hyp['idx'] = int(indices[i]) if indices is not None else i
                hyps.append(hyp)
        if self.distributed:
            gathered = [None] * self.dp_size
            dist.all_gather_object(gathered, hyps)
            hyps = [h for part in gathered for h in part]
        return hyps

    def _update_loss_scale(self) -> bool:
        """GradScaler.update() (torch.amp.GradScaler: growth 2, backoff 0.5) on the found_inf flag the norm kernel left in
        stats[5]; reading it is float16 training's one host synchronisation per step (the scheduler and the step counter only
        advance on clean steps, pasero/training.py:489-503). Returns True when the update was skipped."""
        found_inf = bool(self.stats[5].item() != 0)
        if found_inf:
            self.loss_scale *= 0.5
            self._growth_tracker = 0
            self.overflows += 1
            if self.loss_scale < 1e-4:
                raise FloatingPointError('too many overflows: the loss scaling factor has reached its minimum value, '
                                         'try lowering the learning rate')
        else:
            self._growth_tracker += 1
            if self._growth_tracker == self.growth_interval:
                self.loss_scale *= 2.0
                self._growth_tracker = 0
        return found_inf

    def _finish_step(self) -> dict:
        if self.fp16 and self._update_loss_scale():      # overflow: nothing was updated, the step does not count
            host = self._stats_ring[self.steps % len(self._stats_ring)]
            host.copy_(self.stats, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            return LazyLogs(self, host, event, self.scheduler.get_last_lr()[0], overflow=True)
        self.scheduler.step()
        self.optimizer_steps += 1
        for g in self._stepped_groups:
            self.group_steps[g] += 1
        self._stepped_groups = ()
        self.steps += 1
        self.saved = False
        host = self._stats_ring[self.steps % len(self._stats_ring)]
        host.copy_(self.stats, non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        return LazyLogs(self, host, event, self.scheduler.get_last_lr()[0])

    @staticmethod
    def read_logs(host: Tensor, event, lr: float, loss_scale: float = 1.0, overflow: bool = False) -> dict:
        """Synchronises on the step's single deferred D2H copy and returns python numbers."""
        event.synchronize()
        s = host.tolist()
        if s[7] > 0:    # counted by the loss kernel (csrc/loss.cu); the reference's F.cross_entropy raises for these
            raise RuntimeError(f'{int(s[7])} target ids outside [0, vocabulary size) in this step')
        logs = {'loss': s[0] / math.log(2), 'nll_loss': s[1] / math.log(2), 'num_tokens': int(s[2]), 'num_lines': int(s[3]),
                'gnorm': math.sqrt(max(s[4], 0.0)) if not overflow else float('inf'), 'lr': lr, 'loss_scale': loss_scale}
        if overflow:
            logs['overflow'] = True
        if s[6] > 0:
            logs['failed'] = True      # a micro-batch failed on some rank: no parameter was updated (see train_step)
        return logs

    # ------------------------------------------------------------------ CUDA-graph path
    def train_step_graphed(self, batch: dict) -> 'LazyLogs':
        """Same update as train_step for update_freq == 1, with forward + backward + clip + Adam replayed from a CUDA
        graph keyed on the batch shape. Inputs are copied into static buffers; step scalars are refreshed in device
        memory before the replay; dropout offsets advance through `rng_base`."""
        assert self.update_freq == 1, 'graph capture covers the update_freq=1 step'
        tensors = {k: v for k, v in batch.items() if torch.is_tensor(v)}
        key = tuple((k, tuple(v.shape)) for k, v in sorted(tensors.items()))
        key = key + tuple(sorted((k, v) for k, v in batch.items() if isinstance(v, (int, tuple))))
        # anything in `meta` may select code paths (adapters by language / domain): a graph is only replayed for the same meta
        key = key + tuple(sorted((k, str(v)) for k, v in (batch.get('meta') or {}).items() if k != 'step'))
        if self.fp16:     # the loss scale is a launch argument of the captured loss kernel: one graph per scale value
            key = key + (('loss_scale', self.loss_scale),)
        self.model.train()
        # data parallel, NCCL: the allreduces stay outside the graph (capturing them hangs on this stack) but are ordered
        # behind external event nodes of the graph, one per bucket, so they overlap the replayed backward
        overlap = self.buckets is not None and self.overlap_graph_allreduce
        if key not in self._graphs:
            static = {}
            for k, v in tensors.items():
                dt = self.dtype if v.is_floating_point() else v.dtype
                static[k] = torch.empty(v.shape, dtype=dt, device=self.device)
            extras = {k: v for k, v in batch.items() if not torch.is_tensor(v)}
            K.RNG_OFFSET_BASE = self.rng_base
            ops.DropoutState.offset = 0
            self._write_hyper()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            if self.buckets is not None:
                self.buckets.start_step(sync=False)     # warm-up pass: no collectives
            with torch.cuda.stream(side):   # warm-up outside capture (lazy inits, workspace sizes, NCCL communicators)
                for k, v in tensors.items():
                    static[k].copy_(v)
                self._graph_body(static, extras)
                # the warm-up pass accumulated real gradients into the fp32 main-gradient arena; with buckets they are only
                # flushed (and zeroed) after the replay below, which would add this batch's gradient a second time
                self.arena.grad32.zero_()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            ops.DropoutState.offset = 0
            events, stats_event = None, None
            if self.buckets is not None:
                self.buckets.start_step(sync=overlap)
                if overlap:
                    events = self.buckets.capture_events = []
                    stats_event = torch.cuda.Event(external=True)
            try:
                with torch.cuda.graph(graph):
                    self._graph_body(static, extras, after_forward=stats_event.record if overlap else None)
                    if overlap:
                        self.buckets.finish(None)
                        ops.WgradStream.join()      # finish() flushed the remaining buckets on the weight-gradient stream
            finally:
                if self.buckets is not None:
                    self.buckets.capture_events = None
            # (a replay runs no Python: the adapter groups this batch shape + meta writes gradients for are noted now)
            groups = {self._group_of[i] for i in self.direct.written if i in self._group_of} if self.grouped else None
            self._graphs[key] = (graph, static, events, stats_event, groups)
        graph, static, events, stats_event, groups = self._graphs[key]
        for k, v in tensors.items():
            static[k].copy_(v, non_blocking=True)
        self._write_hyper()
        graph.replay()
        self.rng_base += 1
        if events is not None:
            self.buckets.reduce_behind_events(events, self.stats, stats_event)
        elif self.buckets is not None:
            self.buckets.allreduce_all(self.stats)
        self._optimizer_kernels(groups)
        return self._finish_step()

    def _graph_body(self, static: dict, extras: dict, after_forward=None) -> None:
        self.stats.zero_()
        if self.grouped:
            self.arena.grad.zero_()
        ops.DirectGrads.active = self.direct
        self.direct.start_step()
        batch = dict(static)
        batch.update(extras)
        self._forward_backward(batch, after_forward)
        ops.WgradStream.join()                       # (a captured graph must end with every forked stream joined)
        if self.buckets is None:
            self.arena.flush32(self._ranges32_all)   # with buckets: flushed per bucket right before its allreduce

    # ------------------------------------------------------------------ checkpoints (reference layout)
    def named_arena(self, flat: Tensor) -> dict:
        """A flat arena (gradients, Adam moments) as tensors keyed by the reference's parameter names: the trainable
        parameters are pointed at their slices of `flat` and read back through the model's state_dict hooks (fused
        k|v|q projection -> three `*_proj` entries, tap-major conv weights -> [Cout, Cin, KW]). Frozen parameters are not
        part of any arena and are left out, like in the reference's name-keyed optimizer state."""
        saved = [p.data for p in self.arena.params]
        wanted = set(self.reference_names)
        try:
            for i, p in enumerate(self.arena.params):
                p.data = self.arena.view(flat, i)
            return {k: v.detach().clone() for k, v in self.model.state_dict().items() if k in wanted}
        finally:
            for p, d in zip(self.arena.params, saved):
                p.data = d

    def _named_state(self, flat: Tensor) -> dict:
        return {k: v.cpu() for k, v in self.named_arena(flat).items()}

    def _load_named_state(self, flat: Tensor, state: dict) -> None:
        """Inverse of named_arena for the trainable parameters' entries of `state`; anything else in `state` (frozen
        parameters written by an older checkpoint) is ignored rather than copied into live weights."""
        saved = [p.data for p in self.arena.params]
        wanted = set(self.reference_names)
        try:
            for i, p in enumerate(self.arena.params):
                p.data = self.arena.view(flat, i)
            self.model.load_state_dict({k: v for k, v in state.items() if k in wanted}, strict=False)
        finally:
            for p, d in zip(self.arena.params, saved):
                p.data = d

    def model_state_dict(self) -> dict:
        return {k: v.detach().clone().cpu() for k, v in self.model.state_dict().items()}

    def optimizer_state_dict(self) -> dict:
        """{'state': {param_name: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} — the reference keys
        optimizer state by parameter *name* (pasero/training.py:899-906)."""
        m, v = self._named_state(self.arena.exp_avg), self._named_state(self.arena.exp_avg_sq)
        names = list(m)

        def steps_of(name):     # adapter groups (`*_adapters_by`) count their own steps, like the reference's per-parameter state
            g = self._reference_group.get(name)
            return self.optimizer_steps if g is None else self.group_steps[g]
        # (a parameter that never received a gradient has no optimizer state in the reference: optimization.py:72-74)
        state = {k: {'step': steps_of(k), 'exp_avg': m[k], 'exp_avg_sq': v[k]} for k in names if steps_of(k) > 0 or not self.grouped}
        group = {'lr': self.scheduler.get_last_lr()[0], 'betas': tuple(self.cfg.adam_betas), 'eps': self.cfg.adam_eps,
                 'weight_decay': self.cfg.weight_decay, 'amsgrad': False, 'params': names}
        return {'state': state, 'param_groups': [group]}

    def scaler_state_dict(self) -> dict:
        """torch.amp.GradScaler.state_dict() of the reference's scaler ({} when it is disabled: bfloat16)."""
        if not self.fp16:
            return {}
        return {'scale': self.loss_scale, 'growth_factor': 2.0, 'backoff_factor': 0.5, 'growth_interval': self.growth_interval,
                '_growth_tracker': self._growth_tracker}

    def save(self, model_dir: Optional[str] = None) -> None:
        """model_{steps}.bin / optimizer_{steps}.bin / metrics_{steps}.bin + model_last.bin symlink
        (pasero/training.py:736-780, 878-939)."""
        model_dir = model_dir or self.cfg.model_dir
        if self.distributed and dist.get_rank() != 0:
            return
        os.makedirs(model_dir, exist_ok=True)
        args = self.cfg.as_dict()
        torch.save({'args': args, 'model': self.model_state_dict(), 'steps': self.steps},
                   os.path.join(model_dir, f'model_{self.steps}.bin'))
        torch.save({'args': args, 'optimizer': self.optimizer_state_dict(), 'scheduler': self.scheduler.state_dict(),
                    'scaler': self.scaler_state_dict(), 'steps': self.steps}, os.path.join(model_dir, f'optimizer_{self.steps}.bin'))
        torch.save({'args': args, 'metrics': dict(self.metrics) if isinstance(self.metrics, dict) else {},
                    'last_valid': 0, 'best_score': None, 'patience': None, 'steps': self.steps},
                   os.path.join(model_dir, f'metrics_{self.steps}.bin'))
        link = os.path.join(model_dir, 'model_last.bin')
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(f'model_{self.steps}.bin', link)
        self.saved = True

    def load(self, model_dir: str, steps: Optional[int] = None, load_train_state: bool = True) -> None:
        if steps is None:
            target = os.readlink(os.path.join(model_dir, 'model_last.bin'))
            steps = int(target[len('model_'):-len('.bin')])
        ckpt = torch.load(os.path.join(model_dir, f'model_{steps}.bin'), map_location='cpu', weights_only=False)
        self.model.load_state_dict(ckpt['model'], strict=False)
        self.steps = ckpt.get('steps', 0)
        opt_path = os.path.join(model_dir, f'optimizer_{steps}.bin')
        if load_train_state and os.path.exists(opt_path):
            opt = torch.load(opt_path, map_location='cpu', weights_only=False)
            state = opt['optimizer']['state']
            self._load_named_state(self.arena.exp_avg, {k: v['exp_avg'] for k, v in state.items()})
            self._load_named_state(self.arena.exp_avg_sq, {k: v['exp_avg_sq'] for k, v in state.items()})
            base = [v['step'] for k, v in state.items() if self._reference_group.get(k) is None]
            self.optimizer_steps = int(base[0]) if base else max((int(v['step']) for v in state.values()), default=0)
            for g in self.group_keys:
                steps = [int(v['step']) for k, v in state.items() if self._reference_group.get(k) == g]
                self.group_steps[g] = steps[0] if steps else 0
            if self.grouped:     # groups without an entry were never stepped: their moments are zero
                absent = [g for g in self.group_keys if self.group_steps[g] == 0]
                for g in absent:
                    lo, hi = self.arena.group_ranges[g]
                    self.arena.exp_avg[lo:hi].zero_()
                    self.arena.exp_avg_sq[lo:hi].zero_()
            self.scheduler.load_state_dict(opt['scheduler'])
            if self.fp16 and opt.get('scaler'):
                self.loss_scale = float(opt['scaler']['scale'])
                self._growth_tracker = int(opt['scaler'].get('_growth_tracker', 0))
        if self.arena.master is not None:       # the master copy follows the loaded 16-bit weights (build_fp32_params)
            self.arena.master.copy_(self.arena.param)


class DevicePrefetcher:
    """Device-resident batches from an iterator of host batches (dicts / tuples / lists of pinned tensors): the
    host->device copies of batch i+1 are issued on a side stream while step i computes, which is the job the
    DataLoader's pinned-memory prefetch + `utils.move_to_device` do in the reference (pasero/training.py:367-381,
    pasero/utils.py:691-702). Non-tensor entries pass through unchanged."""

    def __init__(self, batches, device, transform=None, stream=None):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.transform = transform
        # pass the same stream to successive prefetchers: the caching allocator keeps one block pool per stream, so a
        # fresh stream starts with cudaMalloc calls (each one synchronises the device)
        self.stream = stream if stream is not None else torch.cuda.Stream(self.device)
        self._next = None
        self._fetch()

    def _to_device(self, x):
        if torch.is_tensor(x):
            return x.to(self.device, non_blocking=True)
        if isinstance(x, dict):
            return {k: self._to_device(v) for k, v in x.items()}
        if isinstance(x, (tuple, list)):
            return type(x)(self._to_device(v) for v in x)
        return x

    def _record(self, x, stream):
        if torch.is_tensor(x):
            x.record_stream(stream)
        elif isinstance(x, dict):
            for v in x.values():
                self._record(v, stream)
        elif isinstance(x, (tuple, list)):
            for v in x:
                self._record(v, stream)

    def _fetch(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            dev = self._to_device(host)
            if self.transform is not None:
                dev = self.transform(dev)
            event = torch.cuda.Event()
            event.record(self.stream)
        self._next = (dev, event)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, event = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(event)
        self._record(dev, cur)
        self._fetch()
        return dev


class LazyLogs(Mapping):
    """Step logs whose values are fetched from the deferred D2H copy on first access (no sync inside the step). A
    read-only Mapping: every access path (`[]`, `get`, `in`, `len`, iteration, `dict(logs)`, `{**logs}`) loads first.
    The pinned host slot is reused `len(ring)` steps later; reading after that raises instead of returning another
    step's numbers."""

    def __init__(self, trainer: 'Trainer', host: Tensor, event, lr: float, overflow: bool = False):
        self._src, self._data = (host, event, lr, trainer.loss_scale, overflow), None
        self._trainer, self._step = trainer, trainer.steps

    def _load(self) -> dict:
        if self._data is None:
            if self._trainer.steps - self._step >= len(self._trainer._stats_ring):
                raise RuntimeError(f'logs of step {self._step} were not read before their host buffer was reused '
                                   f'({len(self._trainer._stats_ring)} steps later)')
            self._data = Trainer.read_logs(*self._src)
        return self._data

    def __getitem__(self, k):
        return self._load()[k]

    def __iter__(self):
        return iter(self._load())

    def __len__(self):
        return len(self._load())

    def __repr__(self):
        return repr(self._load())
