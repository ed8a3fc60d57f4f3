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
    return Transformer(cfg.model_cfg, cfg, task)
