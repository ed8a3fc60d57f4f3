"""Headline benchmark: target tokens/s of the speech-translation training step (BASELINE.json config 2:
IWSLT2023-style ST, 80-bin log-mel fbank + conv subsampler + 12/6-layer Transformer, synthetic 16 kHz audio).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torch.distributed.run)
    python bench.py --impl reference ...                      (the reference's algorithm on the host CPU cores)

A step = one pass of the hot path over one batch: PCM -> fused fbank+CMVN -> conv subsampler -> 12-layer encoder ->
6-layer decoder -> tied projection + label-smoothed CE -> backward -> (gradient allreduce) -> clip + Adam.
Per GPU: B = 64 utterances of 10 s (160 000 samples, 998 frames -> 250 encoder positions), 49-token decoder inputs
(48 targets) => 3 072 target tokens per GPU per step; weak scaling over GPUs.
Prints ONE JSON line (see the task contract for the keys).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B_PER_GPU, N_SAMPLES, T_DEC, VOCAB = 64, 160000, 49, 8032
METRIC = 'target tokens/sec/box (ST train step: fbank + 12/6 Transformer, fwd+bwd+Adam)'


def st_config(**kw):
    from pasero_b200 import TrainingConfig
    opts = dict(task='speech_translation', arch='transformer', encoder_layers=12, decoder_layers=6, input_dim=80,
                conv_input_dim=80, conv_channels=1024, conv_kernel_sizes=[5, 5], encoder_max_len=3000, dtype='bfloat16',
                batch_size=64000, label_smoothing=0.1, dropout=0.1, lr=0.0005, warmup=4000, max_steps=100000, seed=1234)
    opts.update(kw)
    return TrainingConfig.from_dict(opts)


def synthetic_batch(rank: int, batch: int):
    rng = np.random.RandomState(1234 + rank)
    pcm = (rng.rand(batch, N_SAMPLES).astype(np.float32) - 0.5)
    dec = rng.randint(4, VOCAB, size=(batch, T_DEC)).astype(np.int64)
    dec[:, 0] = 2
    dec[:, -1] = 2
    return torch.from_numpy(pcm), torch.from_numpy(dec)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------- reference arm
def oracle_step_factory(batch: int, threads: int):
    """The reference's algorithm (oracle port: numpy fbank+CMVN, torch-CPU fp32 model, Adam) on the host cores."""
    from oracle import fbank as of
    from oracle import model as om
    from oracle import optim as oo
    torch.set_num_threads(threads)
    cfg = om.STConfig(encoder_layers=12, decoder_layers=6, vocab_size=VOCAB)
    g = torch.Generator().manual_seed(0)
    sd = {}
    d, f, c = cfg.embed_dim, cfg.encoder_ffn_dim, cfg.conv_channels

    def w(*shape):
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        return (torch.randn(*shape, generator=g) * fan_in ** -0.5).requires_grad_(True)
    sd['encoder.subsample.conv_layers.0.weight'], sd['encoder.subsample.conv_layers.0.bias'] = w(c, 80, 5), w(c)
    sd['encoder.subsample.conv_layers.1.weight'], sd['encoder.subsample.conv_layers.1.bias'] = w(2 * d, c // 2, 5), w(2 * d)

    def block(p, cross):
        names = ['self_attn'] + (['encoder_attn'] if cross else [])
        for a in names:
            for proj in ('k_proj', 'v_proj', 'q_proj', 'out_proj'):
                sd[f'{p}.{a}.{proj}.weight'], sd[f'{p}.{a}.{proj}.bias'] = w(d, d), w(d)
            sd[f'{p}.{a}_layer_norm.weight'] = torch.ones(d, requires_grad=True)
            sd[f'{p}.{a}_layer_norm.bias'] = torch.zeros(d, requires_grad=True)
        sd[f'{p}.fc1.weight'], sd[f'{p}.fc1.bias'], sd[f'{p}.fc2.weight'], sd[f'{p}.fc2.bias'] = w(f, d), w(f), w(d, f), w(d)
        sd[f'{p}.final_layer_norm.weight'] = torch.ones(d, requires_grad=True)
        sd[f'{p}.final_layer_norm.bias'] = torch.zeros(d, requires_grad=True)
    for i in range(cfg.encoder_layers):
        block(f'encoder.layers.{i}', False)
    sd['decoder.embed_tokens.weight'] = w(VOCAB, d)
    for i in range(cfg.decoder_layers):
        block(f'decoder.layers.{i}', True)
    pcm, dec = synthetic_batch(0, batch)
    state, hp = {}, dict(lr=0.0005, warmup=4000, max_steps=100000, clip_norm=1.0)
    counter = [0]

    def step():
        feats = np.stack([of.fbank_cmvn(pcm[b].numpy(), cmvn=True) for b in range(batch)])
        b = {'encoder_input': torch.from_numpy(feats), 'encoder_input_length': torch.full((batch,), feats.shape[1]),
             'decoder_input': dec}
        logs = oo.train_step(cfg, sd, [b], state, counter[0], hp, om.forward)
        counter[0] += 1
        return logs['num_tokens']
    return step


def time_oracle(batch: int, steps: int, warmup: int, threads: int):
    step = oracle_step_factory(batch, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    tokens = 0
    for _ in range(steps):
        tokens += step()
    dt = time.perf_counter() - t0
    return tokens / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the SAME step as our arm: all 64 utterances of one GPU's batch per step (a few seconds each on the host cores)
    batch = B_PER_GPU
    value, ms = time_oracle(batch, max(1, args.steps), max(0, args.warmup), threads)
    sample = (f'full step: {batch} utterances x 10 s = {batch * (T_DEC - 1)} target tokens, {args.steps} steps after {args.warmup} '
              f'warm-up; oracle port of the reference (numpy fbank+CMVN, torch-CPU fp32 model, Adam) on {threads} threads')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'tokens/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'tokens/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_config(n_gpus: int) -> dict:
    return {'workload': 'BASELINE configs[1]: IWSLT2023-style ST, 80-bin fbank + conv subsampler (k=5,5) + 12/6-layer '
                        'Transformer d=512 ffn=2048 h=8, V=8032, dropout 0.1, label smoothing 0.1',
            'utterances_per_gpu': B_PER_GPU, 'samples_per_utterance': N_SAMPLES, 'frames': 998, 'encoder_len': 250,
            'target_len': T_DEC - 1, 'global_batch_utterances': B_PER_GPU * max(1, n_gpus),
            'parallelism': f'dp{max(1, n_gpus)}', 'l2_policy': 'per-step working set (>1 GB of activations) exceeds the 126 MB L2'}


# ---------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--dtype', default='bfloat16', choices=['bfloat16', 'float16'],
                    help='float16: the same step with loss scaling and FlatFP16Adam (fp32 master weights); reported next to the bf16 headline')
    ap.add_argument('--bucket-mb', type=float, default=0.0, help='gradient allreduce bucket size (MB); 0 = the configuration default (25)')
    ap.add_argument('--no-eager', action='store_true', help='skip the stock-PyTorch-eager comparator leg (N = 1 only)')
    ap.add_argument('--workload', default='st', choices=['st', 'whisper_small', 'big_adapters'],
                    help="st = the headline (BASELINE configs[1]/[2]); whisper_small / big_adapters = configs[3] / [4], reported separately")
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload != 'st':
        return run_other_workload(args)

    import torch.distributed as dist
    from pasero_b200 import build_model, frontend
    from pasero_b200 import kernels as K
    from pasero_b200.task import Dictionary, SpeechTranslationTask
    from pasero_b200.training import Trainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            del os.environ['NCCL_DEBUG']        # at these levels NCCL prints its version banner on stdout, which carries ONE JSON line
        dist.init_process_group('nccl', device_id=dev)
    assert world == max(1, args.gpus), f'--gpus {args.gpus} but WORLD_SIZE={world}'
    W = max(3, args.warmup)

    extra_cfg = {'bucket_bytes': int(args.bucket_mb * 2 ** 20)} if args.bucket_mb > 0 else {}
    if args.dtype == 'float16':
        extra_cfg.update(dtype='float16', memory_efficient_fp16=False, flat_fp16=True)
    cfg = st_config(**extra_cfg)
    elem_dtype = torch.float16 if args.dtype == 'float16' else torch.bfloat16
    task = SpeechTranslationTask(cfg, Dictionary(tokens=[f't{i}' for i in range(VOCAB - 4)]))
    torch.manual_seed(1234)
    model = build_model(cfg, task)
    trainer = Trainer(cfg, task, model)
    sub = model.encoder.subsample

    pcm_host, dec_host = synthetic_batch(rank, B_PER_GPU)
    pcm_host, dec_host = pcm_host.pin_memory(), dec_host.pin_memory()
    plan = frontend.FbankPlan([N_SAMPLES] * B_PER_GPU, dev)
    F = plan.max_frames
    rows, front = sub.padded_rows(F, 0)
    tokens_per_step = B_PER_GPU * (T_DEC - 1) * world
    pcm_dev = pcm_host.to(dev).reshape(-1).contiguous()
    dec_dev = dec_host.to(dev)
    use_graph = not args.no_graph

    def step(pcm_flat, dec):
        _, padded = frontend.fbank(pcm_flat, plan, cmvn=True, padded_frames=rows, front_pad=front, keep_fp32=False, dtype=elem_dtype)
        batch = {'encoder_input': padded, 'encoder_input_length': plan.frames_tensor, 'decoder_input': dec,
                 'padded_layout': (F,), 'meta': {}}
        if use_graph:
            return trainer.train_step_graphed(batch)
        return trainer.train_step(iter([batch]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`) ----
    for _ in range(W):
        logs = step(pcm_dev, dec_dev)
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        logs = step(pcm_dev, dec_dev)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clock_info = clocks.stop() if clocks else None
    last = dict(logs.items())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    value = tokens_per_step / (ms_per_step / 1e3)

    # ---- end-to-end through the public API with host buffers (`e2e`) ----
    # Every step's PCM + token ids start in pinned HOST memory and are copied inside the timed region
    # (DevicePrefetcher: the copy of batch i+1 runs on a side stream while step i computes, as the reference's
    # pinned-memory DataLoader prefetch does); every step's logs are read back to the host (one step late, so that the
    # read does not drain the queue).
    from pasero_b200.training import DevicePrefetcher
    copy_stream = torch.cuda.Stream(dev)
    for pcm_d, dec_d in DevicePrefetcher(((pcm_host, dec_host) for _ in range(W)), dev, stream=copy_stream):   # warm-up
        step(pcm_d.reshape(-1), dec_d)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss_sink, prev = 0.0, None
    for pcm_d, dec_d in DevicePrefetcher(((pcm_host, dec_host) for _ in range(args.steps)), dev, stream=copy_stream):
        lg = step(pcm_d.reshape(-1), dec_d)
        if prev is not None:
            loss_sink += prev['loss']    # device -> host read of the previous step's result
        prev = lg
    loss_sink += prev['loss']
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = tokens_per_step / (t.item() / args.steps / 1e3)
    h2d = pcm_host.numel() * 4 + dec_host.numel() * 8
    d2h = 8 * 8

    # ---- dominant kernel roofline (the tcgen05 GEMM) + per-family table, all launches of one step timed in situ ----
    # (every rank runs the instrumented step: it contains the gradient allreduce)
    roofline, gemm_share, launches, families = step_rooflines(trainer, step, pcm_dev, dec_dev, ms_per_step)
    # ---- BASELINE's second metric: the fbank(+CMVN) frontend alone ----
    fbank = time_fbank(pcm_dev, plan, rows, front) if rank == 0 else None
    if fbank is not None:
        families['fbank'] = {k: fbank[k] for k in ('ms_per_step', 'achieved', 'unit', 'peak', 'frac')}

    cpu, eager = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_utt, n_steps = 8, 3
        v, cms = time_oracle(n_utt, n_steps, 1, threads)
        cpu = {'value': v, 'unit': 'tokens/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n_steps} steps over {n_utt} of the step\'s {B_PER_GPU} utterances ({n_utt * (T_DEC - 1)} target tokens each) after 1 '
                         f'warm-up step; oracle port of the reference (numpy fbank+CMVN, torch-CPU fp32 model, Adam); the full-size '
                         f'step is what `--impl reference` times', 'ms_per_step': cms}
    if rank == 0 and world == 1 and not args.no_eager:
        del trainer, model          # free the arenas before the eager model is built
        torch.cuda.empty_cache()
        eager = time_gpu_eager(dev)

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.dtype == 'bfloat16' else 'f16', 'data': 'synthetic', 'config': workload_config(world),
            'e2e': {'value': e2e_value, 'unit': 'tokens/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches * args.steps if launches else None, 'clocks': clock_info, 'roofline': roofline,
            'cpu_baseline': cpu, 'fbank': fbank, 'families': families, 'gpu_eager': eager,
            'gemm_share_of_step': gemm_share, 'cuda_graph': use_graph,
            'last_step_logs': {k: (float(v) if isinstance(v, (int, float)) else v) for k, v in last.items()},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_other_workload(args):
    """BASELINE configs[3] (Whisper-small fine-tuning shape: 30 s mel inputs, 12+12 layers, d = 768, V = 51 865) and
    configs[4] (Transformer-big + language adapters shape: d = 1024, V = 70 376, text batches, adapters trained, base frozen):
    the same Trainer step on synthetic features / token ids. Reported separately from the headline; same JSON keys."""
    import torch.distributed as dist
    from pasero_b200 import TrainingConfig, build_model
    from pasero_b200.task import Dictionary, SpeechTranslationTask
    from pasero_b200.training import DevicePrefetcher, Trainer
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    W = max(3, args.warmup)
    rng = np.random.RandomState(4321 + rank)
    if args.workload == 'whisper_small':
        V, B, T = 51865, 16, 129
        cfg = TrainingConfig.from_dict(dict(
            arch='whisper_base', encoder_layers=12, decoder_layers=12, embed_dim=768, encoder_ffn_dim=3072, decoder_ffn_dim=3072,
            encoder_attention_heads=12, decoder_attention_heads=12, conv_channels=768, encoder_max_len=3000, decoder_max_len=448,
            batch_size=64000, dropout=0.1, label_smoothing=0.1, dtype='bfloat16', lr=1e-4, warmup=4000, max_steps=100000, seed=1234))
        task = SpeechTranslationTask(cfg, Dictionary(tokens=[f't{i}' for i in range(V - 4)]))
        enc_host = torch.from_numpy(rng.randn(B, 3000, 80).astype(np.float32)).to(torch.bfloat16)
        enc_len = torch.full((B,), 3000, dtype=torch.int64)
        what = ('BASELINE configs[3]: Whisper-small shape (12+12 layers, d=768, 12 heads, ffn 3072, pre-norm, GELU, learned positions, '
                'conv k=3,3 strides 1,2), synthetic 30 s mel inputs (16, 3000, 80), 128 targets per line, V=51865, bf16')
    else:
        V, B, T, S = 70376, 128, 33, 32
        cfg = TrainingConfig.from_dict(dict(
            task='translation', arch='adapter_transformer', encoder_layers=6, decoder_layers=6, embed_dim=1024, encoder_ffn_dim=4096,
            decoder_ffn_dim=4096, encoder_attention_heads=16, decoder_attention_heads=16, encoder_adapter_dim=64, decoder_adapter_dim=64,
            encoder_max_len=256, batch_size=4096, dropout=0.1, label_smoothing=0.1, dtype='bfloat16', lr=5e-4, warmup=4000,
            max_steps=100000, seed=1234))
        task = SpeechTranslationTask.text(cfg, Dictionary(tokens=[f't{i}' for i in range(V - 4)]))
        enc_host = torch.from_numpy(rng.randint(4, V, size=(B, S)).astype(np.int64))
        enc_len = torch.full((B,), S, dtype=torch.int64)
        what = ('BASELINE configs[4]: Transformer-big shape (6+6 layers, d=1024, 16 heads, ffn 4096, V=70376, shared embeddings) + '
                'bottleneck adapters (dim 64) after every layer, base frozen (adapter_transformer), 128 x 32 text batches, bf16; '
                'projection + CE over vocabulary chunks (no [M, V] logits in HBM)')
    dec = rng.randint(4, V, size=(B, T)).astype(np.int64)
    dec[:, 0] = 2
    dec[:, -1] = 2
    dec_host = torch.from_numpy(dec)
    torch.manual_seed(1234)
    model = build_model(cfg, task)
    trainer = Trainer(cfg, task, model)
    enc_host, dec_host = enc_host.pin_memory(), dec_host.pin_memory()
    enc_dev, dec_dev, len_dev = enc_host.to(dev), dec_host.to(dev), enc_len.to(dev)
    tokens_per_step = B * (T - 1) * world

    def step(enc, d):
        return trainer.train_step_graphed({'encoder_input': enc, 'encoder_input_length': len_dev, 'decoder_input': d, 'meta': {}})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(W):
        logs = step(enc_dev, dec_dev)
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        logs = step(enc_dev, dec_dev)
    e1.record()
    barrier()
    clock_info = clocks.stop() if clocks else None
    last = dict(logs.items())
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    copy_stream = torch.cuda.Stream(dev)
    for a, b in DevicePrefetcher(((enc_host, dec_host) for _ in range(W)), dev, stream=copy_stream):
        step(a, b)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sink, prev = 0.0, None
    for a, b in DevicePrefetcher(((enc_host, dec_host) for _ in range(args.steps)), dev, stream=copy_stream):
        lg = step(a, b)
        if prev is not None:
            sink += prev['loss']
        prev = lg
    sink += prev['loss']
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    roofline, gemm_share, launches, families = step_rooflines(trainer, step, enc_dev, dec_dev, ms_per_step, eager_launches=2)
    if rank == 0:
        print(json.dumps({
            'metric': 'target tokens/sec/box (train step, fwd+bwd+Adam)', 'value': tokens_per_step / (ms_per_step / 1e3), 'unit': 'tokens/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': W, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': what, 'lines_per_gpu': B, 'target_len': T - 1, 'parallelism': f'dp{world}',
                       'l2_policy': 'per-step working set exceeds the 126 MB L2'},
            'e2e': {'value': tokens_per_step / (e2e_ms / 1e3), 'unit': 'tokens/s',
                    'h2d_bytes_per_step': enc_host.numel() * enc_host.element_size() + dec_host.numel() * 8, 'd2h_bytes_per_step': 64},
            'gpu_launches': launches * args.steps if launches else None, 'clocks': clock_info, 'roofline': roofline, 'cpu_baseline': None,
            'families': families, 'gemm_share_of_step': gemm_share, 'cuda_graph': True,
            'trainable_parameters': trainer.arena.numel, 'last_step_logs': {k: (float(v) if isinstance(v, (int, float)) else v) for k, v in last.items()},
        }))
    if world > 1:
        dist.destroy_process_group()


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def time_fbank(pcm_dev, plan, rows, front, iters: int = 20) -> dict:
    """frames/s and algorithmic GB/s of the fused fbank + utterance-CMVN frontend alone, on the step's own PCM batch
    (resident in HBM), L2 flushed between launches (256 MB write), CUDA events, median. Algorithmic bytes = fp32 PCM in +
    fp32 features out, each once (961.3 B per frame)."""
    from pasero_b200 import frontend
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=pcm_dev.device)
    for _ in range(3):
        frontend.fbank(pcm_dev, plan, cmvn=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frontend.fbank(pcm_dev, plan, cmvn=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    nbytes = 4 * plan.total_samples + 4 * 80 * plan.total_frames
    peaks = measured_peaks()
    peak = peaks.get('hbm_gbs', 6650.0)
    gbs = nbytes / ms / 1e6
    return {'frames_per_s': plan.total_frames / (ms / 1e3), 'ms_per_step': ms, 'achieved': gbs, 'unit': 'GB/s', 'peak': peak,
            'frac': gbs / peak, 'bound': 'hbm', 'algorithmic_bytes': nbytes, 'frames': plan.total_frames,
            'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s (B200_PROFILING.md)',
            'method': f'kernel alone, {iters} launches, L2 flushed between them, CUDA events, median; fp32 features out (the '
                      'training step itself only keeps the bf16 padded layout)'}


def time_gpu_eager(dev, steps: int = 10, warmup: int = 3) -> dict:
    """Comparator demanded by SURVEY 8d: the SAME step written with stock PyTorch ops in bf16 on this GPU (cuDNN conv1d,
    cuBLAS linears, F.scaled_dot_product_attention with a float key-padding mask as pasero/models/modules.py:659-718 builds it,
    F.layer_norm, F.cross_entropy on fp32 logits as transformer.py:355-372, fused torch Adam) — what the reference's model
    code dispatches to. Feature extraction = torchaudio.compliance.kaldi.fbank per utterance when torchaudio is present."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    d, f, c, H, V, Le, Ld = 512, 2048, 1024, 8, VOCAB, 12, 6
    P = {}

    def w(*shape):
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        return torch.nn.Parameter((torch.randn(*shape, device=dev) * fan_in ** -0.5).to(torch.bfloat16))

    def attn_params(p):
        for n in ('q', 'k', 'v', 'o'):
            P[f'{p}.{n}.w'], P[f'{p}.{n}.b'] = w(d, d), w(d)

    def ln_params(p):
        P[f'{p}.g'] = torch.nn.Parameter(torch.ones(d, device=dev, dtype=torch.bfloat16))
        P[f'{p}.b'] = torch.nn.Parameter(torch.zeros(d, device=dev, dtype=torch.bfloat16))
    P['c0.w'], P['c0.b'], P['c1.w'], P['c1.b'] = w(c, 80, 5), w(c), w(2 * d, c // 2, 5), w(2 * d)
    for i in range(Le):
        attn_params(f'e{i}.sa'); ln_params(f'e{i}.ln1'); ln_params(f'e{i}.ln2')
        P[f'e{i}.fc1.w'], P[f'e{i}.fc1.b'], P[f'e{i}.fc2.w'], P[f'e{i}.fc2.b'] = w(f, d), w(f), w(d, f), w(d)
    P['emb'] = w(V, d)
    for i in range(Ld):
        attn_params(f'd{i}.sa'); attn_params(f'd{i}.ca'); ln_params(f'd{i}.ln1'); ln_params(f'd{i}.ln2'); ln_params(f'd{i}.ln3')
        P[f'd{i}.fc1.w'], P[f'd{i}.fc1.b'], P[f'd{i}.fc2.w'], P[f'd{i}.fc2.b'] = w(f, d), w(f), w(d, f), w(d)
    opt = torch.optim.Adam(list(P.values()), lr=1e-5, fused=True)

    def mha(p, x, mem, mask, causal):
        B, T, _ = x.shape
        S = mem.shape[1]
        q = F.linear(x, P[f'{p}.q.w'], P[f'{p}.q.b']).view(B, T, H, 64).transpose(1, 2)
        k = F.linear(mem, P[f'{p}.k.w'], P[f'{p}.k.b']).view(B, S, H, 64).transpose(1, 2)
        v = F.linear(mem, P[f'{p}.v.w'], P[f'{p}.v.b']).view(B, S, H, 64).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None if causal else mask, is_causal=causal)
        return F.linear(o.transpose(1, 2).reshape(B, T, d), P[f'{p}.o.w'], P[f'{p}.o.b'])

    def block(x, p, fn):
        return F.layer_norm(x + F.dropout(fn(x), 0.1), (d,), P[f'{p}.g'], P[f'{p}.b'])

    def ffn(p):
        return lambda x: F.linear(F.relu(F.linear(x, P[f'{p}.fc1.w'], P[f'{p}.fc1.b'])), P[f'{p}.fc2.w'], P[f'{p}.fc2.b'])
    try:
        import torchaudio
        kaldi = torchaudio.compliance.kaldi
    except Exception:
        kaldi = None
    pcm, dec = synthetic_batch(0, B_PER_GPU)
    pcm, dec = pcm.to(dev), dec.to(dev)
    pos_e = torch.randn(250, d, device=dev, dtype=torch.bfloat16)
    pos_d = torch.randn(T_DEC - 1, d, device=dev, dtype=torch.bfloat16)

    def step():
        if kaldi is not None:
            feats = torch.stack([kaldi.fbank(pcm[b:b + 1] * 32768.0, num_mel_bins=80, sample_frequency=16000.0) for b in range(pcm.shape[0])])
            feats = (feats - feats.mean(1, keepdim=True)) / feats.std(1, keepdim=True, unbiased=False)
        else:
            feats = torch.randn(pcm.shape[0], 998, 80, device=dev)
        x = feats.to(torch.bfloat16).transpose(1, 2)
        x = F.glu(F.conv1d(x, P['c0.w'], P['c0.b'], stride=2, padding=2), dim=1)
        x = F.glu(F.conv1d(x, P['c1.w'], P['c1.b'], stride=2, padding=2), dim=1).transpose(1, 2)
        B, S, _ = x.shape
        x = F.dropout(x * math.sqrt(d) + pos_e[:S], 0.1)
        mask = torch.zeros(B, 1, 1, S, device=dev, dtype=torch.bfloat16)      # float key-padding mask (all keys valid here)
        for i in range(Le):
            x = block(x, f'e{i}.ln1', lambda h: mha(f'e{i}.sa', h, h, mask, False))
            x = block(x, f'e{i}.ln2', ffn(f'e{i}'))
        y = F.dropout(F.embedding(dec[:, :-1], P['emb'], padding_idx=1) * math.sqrt(d) + pos_d, 0.1)
        for i in range(Ld):
            y = block(y, f'd{i}.ln1', lambda h: mha(f'd{i}.sa', h, h, None, True))
            y = block(y, f'd{i}.ln2', lambda h: mha(f'd{i}.ca', h, x, mask, False))
            y = block(y, f'd{i}.ln3', ffn(f'd{i}'))
        logits = y @ P['emb'].T
        loss = F.cross_entropy(logits.float().view(-1, V), dec[:, 1:].reshape(-1), ignore_index=1, reduction='sum', label_smoothing=0.1)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'value': B_PER_GPU * (T_DEC - 1) / (ms / 1e3), 'unit': 'tokens/s', 'ms_per_step': ms, 'steps': steps,
            'what': 'stock PyTorch eager bf16 on this GPU: ' + ('torchaudio kaldi.fbank per utterance + CMVN, ' if kaldi is not None else
                    '(torchaudio missing: random features, no fbank) ') + 'cuDNN conv1d + GLU, cuBLAS linears, F.scaled_dot_product_attention '
                    '(float key-padding mask / is_causal), F.layer_norm, fp32 F.cross_entropy(label_smoothing), clip_grad_norm_, fused '
                    'torch.optim.Adam; same shapes, dropout 0.1, random init; a comparator, not part of the product path'}


def event_bracket_overhead(dev) -> float:
    """ms that an (event, kernel, event) bracket inside a CUDA graph reports on top of the kernel's own duration: the
    record nodes keep the graph from staging the next launch behind the running kernel. Measured with a ~1 us
    elementwise kernel: bracketed time minus its back-to-back time in an un-instrumented graph chain."""
    z = torch.zeros(1024, device=dev)
    n = 64

    def chain(bracket):
        evs = []
        g = torch.cuda.CUDAGraph()
        z.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(n):
                if bracket:
                    e0, e1 = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
                    e0.record()
                z.add_(1.0)
                if bracket:
                    e1.record()
                    evs.append((e0, e1))
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record()
        torch.cuda.synchronize()
        if bracket:
            return sum(a.elapsed_time(b) for a, b in evs) / n
        return s.elapsed_time(e) / n
    return max(0.0, chain(True) - chain(False))


def step_rooflines(trainer, step, pcm_dev, dec_dev, ms_per_step, eager_launches: int = 4):
    """Live timing of every kernel family inside the real step: the CUDA graph of the step is re-captured with an
    external-event record node before and after every kernel wrapper (pb200_gemm_bf16 launches and all the others),
    replayed, and the event pairs of the last replay are read back. (Events around eager launches over-count: each
    bracket then also holds the host's launch latency.) Split-K launches are timed with their memset and reduction
    kernels; FLOPs / bytes are the algorithmic ones (2*M*N*K per GEMM; pasero_b200.kernels work estimators).
    Returns (roofline of the dominant kernel = the tcgen05 GEMM, its share of the step, launches per step, per-family table)."""
    from pasero_b200 import kernels as K
    from pasero_b200 import ops
    peaks = measured_peaks()
    peak = peaks.get('bf16_tflops_sustained')
    peak_kind = 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)'
    if peak is None:
        peak, peak_kind = 1400.0, 'fallback (B200_PROFILING.md, sustained)'
    hbm = peaks.get('hbm_gbs', 6650.0)
    saved = dict(trainer._graphs)
    trainer._graphs.clear()
    K.PROFILE, K.PROFILE_EXTERNAL, K.COUNT = [], True, {'launches': 0}
    K.PROFILE_ALL, K.PROFILE_WORK = [], []
    # kernel durations are taken with every kernel on ONE stream: in the production step the weight-gradient GEMMs run on
    # a second stream next to the dependent chain, where a bracket would also hold the time its kernel waited for SMs
    two_streams = ops.WgradStream.enabled
    ops.WgradStream.enabled = False
    serial_ms = ms_per_step
    marks = []
    orig_body = trainer._graph_body

    def body(*a, **kw):
        marks.append(len(K.PROFILE_ALL))
        orig_body(*a, **kw)
        marks.append(len(K.PROFILE_ALL))
    trainer._graph_body = body
    try:
        step(pcm_dev, dec_dev)            # warm-up pass + capture (both go through the instrumented wrappers)
        trainer._graph_body = orig_body
        if trainer._graphs:
            prof = K.PROFILE[len(K.PROFILE) // 2:]   # the event pairs recorded during capture = the graph's nodes
            # one call = the eager launches around the graph (fbank + CMVN / layout, grad-norm + Adam), warm-up body and captured body (N each)
            launches = (K.COUNT['launches'] - eager_launches) // 2 + eager_launches
            other = list(zip(K.PROFILE_ALL, K.PROFILE_WORK))[marks[2]:marks[3]]   # non-GEMM wrappers of the captured body
        else:                                        # --no-graph: plain events around eager launches
            prof, launches = K.PROFILE, K.COUNT['launches']
            other = list(zip(K.PROFILE_ALL, K.PROFILE_WORK))
        K.PROFILE, K.COUNT, K.PROFILE_ALL, K.PROFILE_WORK = None, None, None, None
        for _ in range(3):
            step(pcm_dev, dec_dev)        # replays: every event node is re-recorded
        torch.cuda.synchronize()
        overhead = event_bracket_overhead(pcm_dev.device) if trainer._graphs else 0.0
        # pb200::gemm_ln_kernel (Linear + dropout + residual + LayerNorm in one kernel, csrc/gemm_ln.cu) is timed with the GEMMs
        # but is a kernel of its own: its time also holds the LayerNorm tail, so it gets its own family row and the roofline
        # of the dominant kernel (pb200::gemm_kernel) is taken over gemm_kernel's launches only; `frac_with_fused_ln` counts
        # both kernels' time against the Linear FLOPs alone.
        tags = K_TAGS = K.PROFILE_TAGS[len(K.PROFILE_TAGS) - len(prof):]
        fused = [(f, s_, e_) for (f, s_, e_), t in zip(prof, tags) if 'fused' in t]
        prof_all = prof
        prof = [(f, s_, e_) for (f, s_, e_), t in zip(prof, tags) if 'fused' not in t]
        tags = [t for t in tags if 'fused' not in t]
        flops = sum(f for f, _, _ in prof)
        ms_raw = sum(s.elapsed_time(e) for _, s, e in prof)
        ms = ms_raw - overhead * len(prof)
        fused_flops = sum(f for f, _, _ in fused)
        fused_ms = sum(s.elapsed_time(e) for _, s, e in fused) - overhead * len(fused)
        # ---- per-family table ----
        fam = {}
        for ((tag, e0, e1), work) in other:
            name = tag.split(' ')[0]
            key = ('attention' if name.startswith('attention') else 'layernorm' if name.startswith('dropout_add_ln') else
                   'loss' if name.startswith('ls_ce') else 'conv_glue' if name in ('glu_fwd', 'glu_bwd', 'col2im', 'pad_rows', 'gelu_fwd', 'gelu_bwd')
                   else 'embedding' if name.startswith('embed') else 'gradient_flush' if name.startswith('grad32') else 'other')
            g = fam.setdefault(key, {'launches': 0, 'ms': 0.0, 'work': 0.0, 'unit': None})
            g['launches'] += 1
            g['ms'] += max(0.0, e0.elapsed_time(e1) - overhead)
            if work is not None:
                g['unit'] = work[0]
                g['work'] += work[1]
        families = {'gemm': {'launches': len(prof), 'ms_per_step': ms, 'achieved': flops / (ms / 1e3) / 1e12, 'unit': 'TFLOP/s',
                             'peak': peak, 'frac': flops / (ms / 1e3) / 1e12 / peak}}
        if fused:
            families['linear_ln_fused'] = {'launches': len(fused), 'ms_per_step': fused_ms, 'achieved': fused_flops / (fused_ms / 1e3) / 1e12,
                                           'unit': 'TFLOP/s', 'peak': peak, 'frac': fused_flops / (fused_ms / 1e3) / 1e12 / peak,
                                           'note': 'pb200::gemm_ln_kernel: out-projection / fc2 GEMM + dropout + residual + LayerNorm; '
                                                   'FLOPs of the Linear only, the time also holds the LayerNorm tail'}
        for key, g in fam.items():
            row = {'launches': g['launches'], 'ms_per_step': g['ms']}
            if g['unit'] == 'flop' and g['ms'] > 0:
                a = g['work'] / (g['ms'] / 1e3) / 1e12
                row.update({'achieved': a, 'unit': 'TFLOP/s', 'peak': peak, 'frac': a / peak})
            elif g['unit'] == 'byte' and g['ms'] > 0:
                a = g['work'] / (g['ms'] / 1e3) / 1e9
                row.update({'achieved': a, 'unit': 'GB/s', 'peak': hbm, 'frac': a / hbm})
            families[key] = row
        if os.environ.get('BENCH_GEMM_DUMP'):   # per-shape table on stderr
            groups = {}
            for (f, s_, e_), tag in zip(prof_all, K_TAGS):
                g = groups.setdefault(tag, [0, 0.0, f])
                g[0] += 1
                g[1] += s_.elapsed_time(e_)
            for tag, (n, t, f) in sorted(groups.items(), key=lambda x: -x[1][1]):
                sys.stderr.write(f'{tag:60s} n={n:3d} total {t * 1e3:8.1f} us  avg {t / n * 1e3:7.1f} us  {f * n / t / 1e9:7.1f} TFLOP/s\n')
        if two_streams:   # the single-stream step the kernel times belong to (un-instrumented), for the share below
            K.PROFILE_EXTERNAL = False
            trainer._graphs.clear()
            for _ in range(2):
                step(pcm_dev, dec_dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                step(pcm_dev, dec_dev)
            e1.record()
            torch.cuda.synchronize()
            serial_ms = e0.elapsed_time(e1) / 5
    finally:
        trainer._graph_body = orig_body
        ops.WgradStream.enabled = two_streams
        K.PROFILE, K.PROFILE_EXTERNAL, K.COUNT, K.PROFILE_ALL, K.PROFILE_WORK = None, False, None, None, None
        trainer._graphs.clear()
        trainer._graphs.update(saved)
    achieved = flops / (ms / 1e3) / 1e12
    # DRAM traffic of the same kernels from the committed ncu pass (profiles/r02_traffic.json, tools/ncu_traffic.py): per step
    traffic, ncu_frac = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
        traffic = tr.get('gemm_dram_bytes_per_step')
        if tr.get('gemm_ms_per_step_ncu'):
            ncu_frac = flops / (tr['gemm_ms_per_step_ncu'] / 1e3) / 1e12 / peak
    except Exception:
        pass
    roofline = {'bound': 'tensor', 'kernel': 'pb200::gemm_kernel (tcgen05 cta_group::2, all instantiations of one step)',
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                'traffic_source': 'profiles/r02_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the GEMM '
                                  'launches of one step' if traffic is not None else None,
                'frac_from_ncu_launch_list': ncu_frac,
                'frac_with_fused_ln': (flops + fused_flops) / ((ms + fused_ms) / 1e3) / 1e12 / peak,
                'fused_linear_ln_launches': len(fused), 'fused_linear_ln_ms_per_step': fused_ms,
                'launches_timed': len(prof), 'gemm_ms_per_step': ms, 'gemm_ms_per_step_raw': ms_raw,
                'event_bracket_overhead_us': overhead * 1e3, 'achieved_raw': flops / (ms_raw / 1e3) / 1e12,
                'algorithmic_gflop_per_step': flops / 1e9, 'peak_source': peak_kind,
                'single_stream_step_ms': serial_ms,
                'method': 'external CUDA-event record nodes around every GEMM launch inside the replayed step graph; the '
                          'bracket overhead (same bracket around a ~1 us kernel minus its un-instrumented chain time) is '
                          'subtracted per launch; *_raw keep the uncorrected sums; frac_from_ncu_launch_list = the same FLOPs over '
                          'the GEMM time of the committed ncu launch list (cold-cache, serialised). Timed with all kernels on one '
                          'stream (single_stream_step_ms); the production step overlaps the weight-gradient GEMMs on a second stream'}
    return roofline, ms / serial_ms, launches, families


if __name__ == '__main__':
    main()
