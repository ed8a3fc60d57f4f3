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
        loss, blogs = forward_fn(cfg, sd, batch['encoder_input'], batch['encoder_input_length'], batch['decoder_input'])
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
