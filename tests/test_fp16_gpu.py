"""`dtype: float16` on the B200: the float16 build of every kernel (csrc/common.cuh: compiled with -DPB200_HALF, entry points
pb200_*_f16) and float16 training — dynamic loss scaling and the two optimizer variants of the reference
(pasero/training.py:289-306, 481-496; pasero/optimization.py:167-298, 366-387).

  * kernels: the tests of tests/test_kernels_gpu.py, re-run with float16 operands against the same fp32 torch references and the
    same tolerances (stated there for bfloat16; float16 carries three more mantissa bits and the inputs stay far from its range);
  * model: forward / backward of the tiny ST model against the fp32 oracle — loss within 1 %, activations within 3 % of the
    tensor maximum, gradients cosine >= 0.985 (the bfloat16 tolerances of test_parity_gpu.py);
  * training: Trainer.train_step against oracle.optim.fp16_train_step (pinned on CPU to the reference's own MixedPrecisionAdam /
    FlatFP16Adam and to torch.amp.GradScaler, tests/test_oracle_cpu.py): loss within 1 % per step, identical loss-scale
    trajectory, and the skipped update / halved scale on an overflow."""
import math

import pytest
import torch

import tests.test_kernels_gpu as tk
from tests.golden import synth
from tests.test_host_cpu import tiny_task, tiny_training_cfg
from tests.test_oracle_cpu import tiny_cfg

pytestmark = pytest.mark.gpu
F16 = torch.float16


@pytest.fixture(autouse=True)
def float16_operands(monkeypatch):
    monkeypatch.setattr(tk, 'DTYPE', F16)


# ---- the kernel tests, float16 build -------------------------------------------------------------------------------------
from tests.test_kernels_gpu import (test_gemm_forward, test_gemm_tile512, test_gemm_epilogues, test_gemm_dgrad_wgrad,  # noqa: E402,F401
                                    test_wgrad_fused_bias_grad, test_gemm_conv_view, test_attention, test_attention_dropout,
                                    test_dropout_add_ln, test_ln_bwd_training_shape_kernel, test_linear_dropout_add_ln_fused,
                                    test_sublayers_fused_linear_ln_matches_two_kernel_path, test_dropout_mask_consistency,
                                    test_colsum, test_embed, test_glu, test_gelu_padded_layout, test_ls_ce, test_adam_and_clip,
                                    test_projection_ce_chunked)


def test_float16_build_is_what_runs():
    """The float16 entry points are distinct symbols, and a float16 GEMM really rounds to half (a value bfloat16 cannot hold)."""
    from pasero_b200 import _lib
    lib, half = _lib.load(), _lib.load(F16)
    assert half.pb200_gemm_bf16 is not lib.pb200_gemm_bf16 and half.pb200_num_sms is lib.pb200_num_sms
    k = tk._k()
    x = torch.full((128, 64), 1.0, dtype=F16, device='cuda')
    w = torch.zeros(128, 64, dtype=F16, device='cuda')
    w[:, 0] = 1.0 + 2.0 ** -10          # exactly representable in float16, rounds to 1.0 in bfloat16
    y = k.linear_fwd(x, w)
    assert y.dtype == F16 and torch.all(y == 1.0 + 2.0 ** -10)


def test_loss_scale_enters_the_loss_gradient():
    k = tk._k()
    B, T, V = 3, 7, 520
    logits = tk._rand(B * T, V, seed=1)
    target = torch.randint(4, V, (B, T), device='cuda')
    target[0, -2:] = 1
    stats = torch.zeros(8, dtype=torch.float64, device='cuda')
    g1, gs = logits.clone(), logits.clone()
    k.ls_ce_fwd_bwd(g1, target, 1, 0.1, stats)
    k.ls_ce_fwd_bwd(gs, target, 1, 0.1, stats.clone(), grad_scale=128.0)
    x = logits.float().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(x, target.view(-1), ignore_index=1, label_smoothing=0.1, reduction='sum')
    loss.backward()
    tk._close(g1, x.grad, 1e-3, 1e-2, 'd loss / d logits')
    tk._close(gs, 128.0 * x.grad, 128e-3, 1e-2, 'scaled d loss / d logits')


@pytest.mark.parametrize('V,D,M,cols', [(40003, 256, 96, 8192), (1003, 128, 77, 256)])
def test_loss_scale_enters_the_chunked_projection_ce(V, D, M, cols):
    """float16 + a vocabulary of chunked size (the NLLB dictionary of the IWSLT2023 recipe: 256 k entries): the loss scale is
    applied where the chunk's d loss / d logits is written (pb200_ce_chunk_grad), so dx and the table gradient are the scaled
    gradients of fp32 PyTorch, and the loss itself is unscaled."""
    import torch.nn.functional as F
    from pasero_b200 import kernels as K, modules, ops
    saved = (K.CE_CHUNK_MIN_VOCAB, K.CE_CHUNK_COLS, ops.ProjCEFn.loss_scale)
    emb = modules.Embedding(V, D, 1).to(F16).cuda()
    g = torch.Generator(device='cuda').manual_seed(6)
    x0 = (torch.randn(M, D, generator=g, device='cuda') * 2).to(F16)
    target = torch.randint(4, V, (1, M), generator=g, device='cuda')
    target[0, -3:] = 1
    scale = 128.0
    try:
        K.CE_CHUNK_MIN_VOCAB, K.CE_CHUNK_COLS = 1, cols
        ops.ProjCEFn.loss_scale = scale
        x = x0.clone().requires_grad_(True)
        stats = torch.zeros(8, dtype=torch.float64, device='cuda')
        loss = ops.ProjCEFn.apply(x, emb.weight, target, 1, 0.1, stats, V)
        loss.backward()
    finally:
        K.CE_CHUNK_MIN_VOCAB, K.CE_CHUNK_COLS, ops.ProjCEFn.loss_scale = saved
    xr = x0.float().requires_grad_(True)
    wr = emb.weight.detach().float()[:V].requires_grad_(True)
    ref = F.cross_entropy(xr @ wr.T, target.reshape(-1), ignore_index=1, reduction='sum', label_smoothing=0.1)
    ref.backward()
    assert abs(loss.item() - ref.item()) / ref.item() < 5e-3
    assert x.grad.dtype == F16 and torch.isfinite(x.grad).all() and torch.isfinite(emb.weight.grad).all()
    tk._close(x.grad.float(), scale * xr.grad, 2e-2 * scale * xr.grad.abs().max().item(), 2e-2, 'scaled dx')
    tk._close(emb.weight.grad.float()[:V], scale * wr.grad, 2e-2 * scale * wr.grad.abs().max().item(), 3e-2, 'scaled dtable')


def test_trainer_float16_large_vocabulary():
    """dtype float16 + FlatFP16Adam on a model whose vocabulary takes the chunked projection + CE path (V >= 32 768): the
    combination the IWSLT2023 recipe trains with its real NLLB dictionary. Loss per step within 1 % of the oracle's float16
    step with master weights, same loss-scale trajectory, no overflow."""
    from oracle import model as om
    from oracle import optim as oo
    from pasero_b200 import TrainingConfig, build_model, kernels as K
    from pasero_b200.task import Dictionary, SpeechTranslationTask
    from pasero_b200.training import Trainer
    import numpy as np
    V = 33000
    assert V >= K.CE_CHUNK_MIN_VOCAB
    t = synth.TINY
    cfg = tiny_training_cfg(dtype='float16', memory_efficient_fp16=False, flat_fp16=True, seed=1)
    task = SpeechTranslationTask(cfg, Dictionary(tokens=[f'tok{i}' for i in range(V - 4)]))
    torch.manual_seed(7)
    model = build_model(cfg, task)
    sd0 = {k: v.detach().clone().half().float() for k, v in model.state_dict().items() if v.is_floating_point() and '_float_tensor' not in k}
    model.load_state_dict(sd0, strict=False)
    trainer = Trainer(cfg, task, model)
    rng = np.random.RandomState(8)
    batches = []
    for b in synth.train_batches()[:4]:
        b = dict(b)
        dec = b['decoder_input'].clone()
        body = dec >= 4
        dec[body] = torch.from_numpy(rng.randint(4, V, int(body.sum())))
        b['decoder_input'] = dec
        batches.append(b)
    import dataclasses
    ocfg = dataclasses.replace(tiny_cfg(), vocab_size=V)
    sd = {k: v.clone().to(F16).requires_grad_(True) for k, v in sd0.items()}
    hp = dict(lr=synth.TRAIN['lr'], warmup=synth.TRAIN['warmup'], max_steps=1000, clip_norm=1.0)
    scaler, ostate = oo.LossScaler(), {}
    master = {k: v.detach().float().clone() for k, v in sd.items()}
    for step in range(6):
        ob = {k: (v.to(F16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batches[step % 4].items()}
        want = oo.fp16_train_step(ocfg, sd, [ob], ostate, step, hp, om.forward, scaler, master=master)
        logs = trainer.train_step(iter([batches[step % 4]]))
        assert 'overflow' not in logs and logs['num_tokens'] == want['num_tokens']
        assert abs(logs['loss'] - want['loss']) / want['loss'] < 1e-2, (step, logs['loss'], want['loss'])
        assert abs(logs['gnorm'] - want['gnorm']) / want['gnorm'] < 5e-2, (step, logs['gnorm'], want['gnorm'])
        assert trainer.loss_scale == scaler.scale


@pytest.mark.parametrize('master', [False, True])
def test_adam_float16_variants_and_skipped_step(master):
    """optimization.Adam on float16 parameters (16-bit rounding points: unscale, normalise, clip) and the fp32 master copy of
    MixedPrecisionAdam / FlatFP16Adam, against the oracle's restatement; then an inf gradient: found_inf set, nothing changes."""
    from oracle import optim as oo
    k = tk._k()
    n = 8 * 1000
    p0 = tk._rand(n, scale=0.1, seed=1)
    scale, ntok, lr, clip, step = 256.0, 96.0, 1e-3, 1.0, 1
    g = (tk._rand(n, scale=0.5, seed=2).float() * scale).to(F16)
    b1, b2, eps = 0.9, 0.999, 1e-8
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    hyper = torch.tensor([lr, b1, b2, eps, 0.0, bc1, math.sqrt(bc2), clip, 1.0, 1 - b1, 1 - b2, lr / bc1, 0.0, 1.0 / scale,
                          float(master), 0], dtype=torch.float32, device='cuda')
    stats = torch.zeros(8, dtype=torch.float64, device='cuda')
    stats[2] = ntok
    p, m, v = p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    mw = p0.float() if master else None
    k.grad_sqnorm(g, hyper, stats)
    k.adam_step(p, g, m, v, hyper, stats, master=mw)
    # oracle (CPU), same arithmetic
    gr = (g.cpu().float() if master else g.cpu().clone()).mul(1.0 / scale)
    gr.mul_(1.0 / ntok)
    norm = oo.clip_grad_norm_([gr], clip)
    assert stats[5].item() == 0 and abs(math.sqrt(stats[4].item()) - norm.item()) / norm.item() < 1e-4
    pr = {'p': p0.cpu().float() if master else p0.cpu().clone()}
    oo.adam_step(pr, {'p': gr}, {}, lr)
    want = pr['p'].to(F16)
    diff = (p.cpu().float() - want.float()).abs()
    assert diff.max() <= 2 ** -10 * p0.float().abs().max().item() and (p.cpu() != want).float().mean() < 2e-3
    if master:
        assert torch.allclose(mw.cpu(), pr['p'], atol=1e-7, rtol=1e-5)
    # overflow: the norm is not finite -> found_inf, and the Adam kernel leaves everything as it is
    g2 = g.clone()
    g2[123] = float('inf')
    before = (p.clone(), m.clone(), v.clone(), None if mw is None else mw.clone())
    k.grad_sqnorm(g2, hyper, stats)
    assert stats[5].item() == 1
    k.adam_step(p, g2, m, v, hyper, stats, master=mw)
    assert torch.equal(p, before[0]) and torch.equal(m, before[1]) and torch.equal(v, before[2])
    assert mw is None or torch.equal(mw, before[3])


def _build(dtype='float16', **kw):
    from pasero_b200 import build_model
    cfg = tiny_training_cfg(dtype=dtype, **kw)
    task = tiny_task(cfg)
    model = build_model(cfg, task)
    model.load_state_dict(synth.tiny_state_dict(), strict=False)
    return cfg, task, model


def _rel(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-6)).item()


def test_model_forward_backward_float16_vs_oracle():
    from oracle import model as om
    cfg, task, model = _build()
    model.to(F16).cuda().train()
    batch = synth.model_batch()
    cb = {k: (v.cuda().to(F16) if torch.is_tensor(v) and v.is_floating_point() else (v.cuda() if torch.is_tensor(v) else v))
          for k, v in batch.items()}
    loss, logs = model(**cb)
    sd = {k: v.clone().requires_grad_(True) for k, v in synth.tiny_state_dict().items()}
    oloss, ologs = om.forward(tiny_cfg(), sd, batch['encoder_input'], batch['encoder_input_length'], batch['decoder_input'])
    assert logs['num_tokens'] == ologs['num_tokens']
    assert abs(loss.item() - oloss.item()) / oloss.item() < 1e-2
    enc_out, _, _ = model.encoder(cb['encoder_input'], cb['encoder_input_length'])
    assert enc_out.dtype == F16
    model.zero_grad()
    loss.backward()
    oloss.backward()
    saved = [(p, p.data) for p in model.parameters()]
    for p in model.parameters():
        p.data = p.grad
    try:
        grads = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if k in sd}
    finally:
        for p, d in saved:
            p.data = d
    bad = {}
    for k, g in grads.items():
        if k.endswith('k_proj.bias'):
            continue
        cos = torch.nn.functional.cosine_similarity(g.flatten(), sd[k].grad.flatten(), dim=0).item()
        if _rel(g, sd[k].grad) > 0.3 or cos < 0.985:
            bad[k] = (_rel(g, sd[k].grad), cos)
    assert not bad, bad


@pytest.mark.parametrize('opts', [dict(), dict(memory_efficient_fp16=False), dict(memory_efficient_fp16=False, flat_fp16=True)])
def test_trainer_float16_against_oracle(opts):
    """30 optimizer steps: loss within 1 % of the oracle's float16 step at every step, same loss scale, gnorm within 5 %."""
    from oracle import model as om
    from oracle import optim as oo
    from pasero_b200.training import Trainer
    cfg, task, model = _build(**opts)
    trainer = Trainer(cfg, task, model)
    assert trainer.fp16 and trainer.master_weights == (not cfg.memory_efficient_fp16)
    assert (trainer.arena.master is not None) == trainer.master_weights and trainer.arena.param.dtype == F16
    osd = {k: v.clone().to(F16).requires_grad_(True) for k, v in synth.tiny_state_dict().items()}
    master = {k: v.detach().float().clone() for k, v in osd.items()} if trainer.master_weights else None
    scaler = oo.LossScaler(init_scale=2.0 ** 7, growth_interval=2 ** 14)
    state, hp = {}, dict(lr=synth.TRAIN['lr'], warmup=synth.TRAIN['warmup'], max_steps=1000, clip_norm=1.0)
    batches = synth.train_batches()
    obatches = [{k: (v.to(F16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in b.items()} for b in batches]
    for step in range(30):
        logs = trainer.train_step(iter([batches[step % len(batches)]]))
        ologs = oo.fp16_train_step(tiny_cfg(), osd, [obatches[step % len(batches)]], state, step, hp, om.forward, scaler, master=master)
        assert not ologs['overflow'] and 'overflow' not in logs
        assert logs['num_tokens'] == ologs['num_tokens'] and logs['loss_scale'] == ologs['loss_scale'] == 128.0
        assert abs(logs['loss'] - ologs['loss']) / ologs['loss'] < 1e-2, (step, logs['loss'], ologs['loss'])
        assert abs(logs['gnorm'] - ologs['gnorm']) / ologs['gnorm'] < 5e-2, (step, logs['gnorm'], ologs['gnorm'])
        assert abs(logs['lr'] - ologs['lr']) < 1e-12
    assert trainer.steps == 30


def test_trainer_float16_overflow_skips_the_update_and_halves_the_scale(tmp_path):
    from pasero_b200.training import Trainer
    cfg, task, model = _build(memory_efficient_fp16=False)
    trainer = Trainer(cfg, task, model)
    batches = synth.train_batches()
    trainer.train_step(iter([batches[0]]))
    trainer.loss_scale = 2.0 ** 40           # the scaled gradients overflow float16
    before = (trainer.arena.param.clone(), trainer.arena.master.clone(), trainer.arena.exp_avg.clone())
    lr_before, steps_before = trainer.scheduler.get_last_lr()[0], trainer.steps
    logs = trainer.train_step(iter([batches[1]]))
    assert logs['overflow'] and logs['loss_scale'] == 2.0 ** 39 and trainer.overflows == 1
    assert trainer.steps == steps_before and trainer.scheduler.get_last_lr()[0] == lr_before
    assert torch.equal(trainer.arena.param, before[0]) and torch.equal(trainer.arena.master, before[1])
    assert torch.equal(trainer.arena.exp_avg, before[2])
    trainer.loss_scale = 128.0
    logs = trainer.train_step(iter([batches[1]]))
    assert 'overflow' not in logs and trainer.steps == steps_before + 1 and math.isfinite(logs['gnorm'])
    # the scaler's state travels with the checkpoint (optimizer_N.bin: 'scaler'), the master copy follows the loaded weights
    trainer.save(str(tmp_path))
    opt = torch.load(str(tmp_path / f'optimizer_{trainer.steps}.bin'), weights_only=False)
    assert opt['scaler']['scale'] == 128.0 and opt['scaler']['growth_interval'] == 2 ** 14
    cfg2, task2, model2 = _build(memory_efficient_fp16=False)
    t2 = Trainer(cfg2, task2, model2)
    t2.load(str(tmp_path))
    assert t2.loss_scale == 128.0 and torch.equal(t2.arena.param, trainer.arena.param)
    assert torch.equal(t2.arena.master, t2.arena.param.float())
