"""pasero_b200 — B200-native (sm_100a) implementation of Pasero's speech-translation training step.

Host code is Python/PyTorch (device memory, autograd graph, streams, torch.distributed); every compute op is a
hand-written CUDA kernel reached through the C ABI in include/pasero_b200.h (libpasero_b200.so, loaded with
ctypes). There is no CPU or eager-PyTorch fallback: importing the package works anywhere, running it needs
the built library and a B200.
"""
from .config import DistributedConfig, TrainingConfig, TransformerConfig  # noqa: F401

__all__ = ['TransformerConfig', 'TrainingConfig', 'DistributedConfig', 'build_model', 'frontend']


def build_model(cfg: TrainingConfig, task):
    """`arch_cls(model_cfg, dist_cfg, task)` of the reference (cli/train.py:220) for arch == 'transformer'."""
    from .transformer import Transformer
    if cfg.arch not in TrainingConfig.ARCH_PRESETS:
        raise NotImplementedError(f'arch {cfg.arch!r}: only {sorted(TrainingConfig.ARCH_PRESETS)} are on the ST hot path')
    model = Transformer(cfg.model_cfg, cfg, task)
    apply_params_regex(model, cfg.extra.get('freeze_params_regex'), cfg.extra.get('train_params_regex'))
    return model


def reference_param_names(model) -> dict:
    """our parameter name -> the reference's names for the same storage. The k/v/q projections are one [3D, D] parameter
    here and three `*_proj` parameters there; the conv weights are stored tap-major. Call before the parameters move into
    a ParamArena (afterwards they all share one storage)."""
    import re
    from collections import defaultdict
    by_storage = defaultdict(list)
    for key, value in model.state_dict(keep_vars=True).items():
        if hasattr(value, 'untyped_storage'):
            by_storage[value.untyped_storage().data_ptr()].append(key)
    out = {}
    for name, p in model.named_parameters():
        refs = by_storage.get(p.untyped_storage().data_ptr())
        if not refs:   # converted copies in the state_dict (conv weights): positional name map
            refs = [re.sub(r'subsample\.(weight|bias)(\d+)$', r'subsample.conv_layers.\2.\1', name)]
        out[name] = refs
    return out


def apply_params_regex(model, freeze_params_regex=None, train_params_regex=None) -> None:
    """`--freeze-params-regex` / `--train-params-regex` (cli/train.py:230-242): the regex is matched against the
    REFERENCE's parameter names and overrides the architecture's own freezing (adapter training)."""
    if not freeze_params_regex and not train_params_regex:
        return
    try:
        import regex as re_mod
    except ImportError:   # the reference uses the `regex` package; `re` agrees on everything but exotic syntax
        import re as re_mod
    names = reference_param_names(model)
    for name, param in model.named_parameters():
        if freeze_params_regex:
            votes = {not re_mod.match(freeze_params_regex, ref) for ref in names[name]}
        else:
            votes = {bool(re_mod.match(train_params_regex, ref)) for ref in names[name]}
        if len(votes) != 1:
            raise NotImplementedError(f'{name} fuses {names[name]}, which the regex splits into trained and frozen parts')
        param.requires_grad = votes.pop()
    for module in model.modules():   # the fused backward nodes skip a projection's weight AND bias gradients together
        w, b = getattr(module, 'in_proj_weight', None), getattr(module, 'in_proj_bias', None)
        if w is not None and b is not None and w.requires_grad != b.requires_grad:
            raise NotImplementedError('a projection with a trained bias and a frozen weight (or the reverse) is not covered')
