"""ORACLE (test infrastructure, not product code): CPU restatement of Pasero's optimizer step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

  lr_at                 LRScheduler.get_lr            pasero/optimization.py:40-52
  clip_grad_norm_       clip_grad_norm_               pasero/optimization.py:390-427
  adam_step             Adam.step (fairseq-style)     pasero/optimization.py:65-149
  train_step            Trainer.train_step            pasero/training.py:324-517 (single process, update_freq
                                                      micro-batches, grads * dp_size / num_tokens, clip, Adam, scheduler)

PARITY PIN: tests/golden/st_tiny_train.pt — loss curve, gnorm, lr and final parameters produced by the unmodified
reference `Trainer.train_step` (CPU, fp32) on seeded batches.
"""
import math

import torch


def lr_at(step: int, lr: float, warmup: int, init_lr: float = 0.0, min_lr: float = 0.0, max_steps: int = 1) -> float:
    if warmup == 0:
        value = lr - step * (lr - min_lr) / max_steps
    elif step < warmup:
        value = init_lr + step * (lr - init_lr) / warmup
    else:
        value = lr * math.sqrt(warmup / step)
    return max(value, min_lr)


@torch.no_grad()
def clip_grad_norm_(grads: list, max_norm: float) -> torch.Tensor:
    if len(grads) > 0:
        total_norm = torch.norm(torch.stack([torch.norm(g, p=2, dtype=torch.float) for g in grads]))
    else:
        total_norm = torch.tensor(0.0)
    if max_norm > 0:
        clip_coef = (float(max_norm) / (total_norm + 1e-6)).clamp_(max=1)
        for g in grads:
            g.mul_(clip_coef)
    return total_norm


@torch.no_grad()
def adam_step(params: dict, grads: dict, state: dict, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
              weight_decay: float = 0.0) -> None:
    """params/grads: name -> tensor (updated in place); state: name -> {'step', 'exp_avg', 'exp_avg_sq'} (fp32)."""
    beta1, beta2 = betas
    for name, p in params.items():
        if name not in grads or grads[name] is None:
            continue
        grad = grads[name].float()
        p32 = p.float()
        st = state.setdefault(name, {})
        if not st:
            st['step'] = 0
            st['exp_avg'] = torch.zeros_like(p32)
            st['exp_avg_sq'] = torch.zeros_like(p32)
        st['step'] += 1
        st['exp_avg'].mul_(beta1).add_(grad, alpha=1 - beta1)
        st['exp_avg_sq'].mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        bc1 = 1 - beta1 ** st['step']
        bc2 = 1 - beta2 ** st['step']
        denom = (st['exp_avg_sq'].sqrt() / math.sqrt(bc2)).add_(eps)
        if weight_decay != 0:
            p32.add_(p32, alpha=-weight_decay * lr)
        p32.addcdiv_(st['exp_avg'], denom, value=-(lr / bc1))
        p.copy_(p32)


def train_step(cfg, sd: dict, batches: list, opt_state: dict, step: int, hp: dict, forward_fn) -> dict:
    """One optimizer update over `batches` (the micro-batches of this step). `sd` holds leaf tensors with
    requires_grad; hp: lr, warmup, init_lr, min_lr, max_steps, betas, eps, weight_decay, clip_norm, dp_size."""
    names = [k for k, v in sd.items() if v.requires_grad]
    for k in names:
        sd[k].grad = None
    logs = {'loss': 0.0, 'nll_loss': 0.0, 'num_tokens': 0, 'num_lines': 0}
    for batch in batches:
        # (`*_adapters_by`: the batch's meta selects the adapters, pasero/models/adapters.py:186-197; parameters of the
        # others keep grad None and are skipped by clip_grad_norm_ and Adam below, like in the reference)
        kw = {'meta': batch['meta']} if getattr(cfg, 'extra', {}).get('encoder_adapters_by') or getattr(cfg, 'extra', {}).get('decoder_adapters_by') else {}
        loss, blogs = forward_fn(cfg, sd, batch['encoder_input'], batch['encoder_input_length'], batch['decoder_input'], **kw)
        loss.backward()
        for k, v in blogs.items():
            logs[k] += v
    coeff = hp.get('dp_size', 1) / logs['num_tokens']
    grads = {}
    for k in names:
        if sd[k].grad is not None:
            sd[k].grad.mul_(coeff)
            grads[k] = sd[k].grad
    gnorm = clip_grad_norm_(list(grads.values()), hp.get('clip_norm', 1.0))
    lr = lr_at(step, hp['lr'], hp.get('warmup', 4000), hp.get('init_lr', 0.0), hp.get('min_lr', 0.0), hp.get('max_steps', 1))
    with torch.no_grad():
        adam_step({k: sd[k] for k in names}, grads, opt_state, lr, hp.get('betas', (0.9, 0.999)), hp.get('eps', 1e-8),
                  hp.get('weight_decay', 0.0))
    logs['gnorm'] = gnorm.item()
    logs['lr'] = lr_at(step + 1, hp['lr'], hp.get('warmup', 4000), hp.get('init_lr', 0.0), hp.get('min_lr', 0.0),
                       hp.get('max_steps', 1))  # the reference logs the scheduler's lr *after* scheduler.step()
    return logs


# --------------------------------------------------------------------------------------------------
# float16 training (dtype: float16): dynamic loss scaling + the two optimizer variants
#   LossScaler            torch.amp.GradScaler as configured at pasero/training.py:289-293 (init 2**7, growth 2,
#                         backoff 0.5, growth_interval 2**14 // update_freq) and used at :481-496
#   fp16_train_step       Trainer.train_step with the scaler enabled: scaled backward (:399-408), unscale_ (+ fp16_to_fp32 for
#                         the master-copy optimizers, optimization.py:366-387, 195-201), grads * dp_size / num_tokens, clip,
#                         scaler.step (skipped on inf / nan) and scaler.update; `master=False` is optimization.Adam on the
#                         16-bit parameters (memory_efficient_fp16, :302-306), `master=True` MixedPrecisionAdam / FlatFP16Adam
#                         (:167-298: identical arithmetic, flat or per-tensor storage)
# PARITY PIN: tests/test_oracle_cpu.py runs the reference's own MixedPrecisionAdam / FlatFP16Adam (imported from
# /root/reference when present) and torch.amp.GradScaler('cpu') against these restatements.
# --------------------------------------------------------------------------------------------------
class LossScaler:
    def __init__(self, init_scale: float = 2.0 ** 7, growth_interval: int = 2 ** 14, growth_factor: float = 2.0,
                 backoff_factor: float = 0.5):
        self.scale, self.growth_interval = float(init_scale), int(growth_interval)
        self.growth_factor, self.backoff_factor, self.tracker = growth_factor, backoff_factor, 0

    def update(self, found_inf: bool) -> None:
        if found_inf:
            self.scale *= self.backoff_factor
            self.tracker = 0
        else:
            self.tracker += 1
            if self.tracker == self.growth_interval:
                self.scale *= self.growth_factor
                self.tracker = 0


def fp16_train_step(cfg, sd: dict, batches: list, opt_state: dict, step: int, hp: dict, forward_fn, scaler: LossScaler,
                    master: dict = None) -> dict:
    """`sd`: float16 leaf tensors. master: None, or name -> fp32 copy of the parameter (updated here, sd receives its rounded
    image). Returns the logs plus 'overflow' and 'loss_scale' (the scale AFTER scaler.update, as the reference logs it)."""
    names = [k for k, v in sd.items() if v.requires_grad]
    for k in names:
        sd[k].grad = None
    logs = {'loss': 0.0, 'nll_loss': 0.0, 'num_tokens': 0, 'num_lines': 0}
    for batch in batches:
        loss, blogs = forward_fn(cfg, sd, batch['encoder_input'], batch['encoder_input_length'], batch['decoder_input'])
        (loss * scaler.scale).backward()
        for k, v in blogs.items():
            logs[k] += v
    inv = 1.0 / scaler.scale
    coeff = hp.get('dp_size', 1) / logs['num_tokens']
    grads = {}
    with torch.no_grad():
        for k in names:
            g = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
            g = g.float() if master is not None else g      # fp16_to_fp32: the copy happens BEFORE unscaling
            g = g.mul(inv)                                  # GradScaler.unscale_ (rounds to fp16 in the 16-bit variant)
            grads[k] = g
        found_inf = any(not torch.isfinite(g).all() for g in grads.values())
        for g in grads.values():
            g.mul_(coeff)
        gnorm = clip_grad_norm_(list(grads.values()), hp.get('clip_norm', 1.0))
        lr = lr_at(step, hp['lr'], hp.get('warmup', 4000), hp.get('init_lr', 0.0), hp.get('min_lr', 0.0), hp.get('max_steps', 1))
        if not found_inf:
            target = master if master is not None else {k: sd[k] for k in names}
            adam_step({k: target[k] for k in names}, grads, opt_state, lr, hp.get('betas', (0.9, 0.999)), hp.get('eps', 1e-8),
                      hp.get('weight_decay', 0.0))
            if master is not None:
                for k in names:
                    sd[k].copy_(master[k])                  # fp32_to_fp16
    scaler.update(found_inf)
    logs.update({'gnorm': gnorm.item(), 'overflow': found_inf, 'loss_scale': scaler.scale})
    logs['lr'] = lr_at(step + (0 if found_inf else 1), hp['lr'], hp.get('warmup', 4000), hp.get('init_lr', 0.0),
                       hp.get('min_lr', 0.0), hp.get('max_steps', 1))
    return logs
