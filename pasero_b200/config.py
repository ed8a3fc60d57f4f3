"""Configuration surface of the ST training path, mirroring the reference's flat YAML / CLI option names
(pasero/config.py: TransformerConfig :1054-1319, TrainingConfig :1469-1868, TranslationTaskConfig :1871-2090,
DistributedConfig :500-546). Only the options this path reads are typed here; anything else found in a
reference YAML is kept in `extra` (and rejected if it switches on a subsystem that is out of scope); a key that is not an option of
the reference either is an error, as in the reference (pasero/config.py:1748-1750).
"""
from dataclasses import dataclass, field, fields
from typing import Optional

import yaml


@dataclass
class TransformerConfig:
    # architecture (defaults = the reference's `transformer` arch: Transformer Base, post-norm, ReLU, sinusoidal)
    embed_dim: int = 512
    encoder_layers: int = 6
    decoder_layers: int = 6
    encoder_attention_heads: int = 8
    decoder_attention_heads: int = 8
    encoder_ffn_dim: int = 2048
    decoder_ffn_dim: int = 2048
    dropout: float = 0.1
    attention_dropout: float = 0.0
    activation_dropout: float = 0.0
    embed_dropout: Optional[float] = None
    decoder_dropout: Optional[float] = None
    label_smoothing: Optional[float] = None        # task-dependent default: 0.1 for (speech_)translation
    activation_fn: str = 'relu'
    encoder_prenorm: bool = False
    decoder_prenorm: bool = False
    scale_embed: bool = True
    scale_attn: bool = True
    attention_key_bias: bool = True                # False for the Whisper family: k_proj has no bias (modules.py:516-522)
    # adapter_transformer (pasero/config.py:1323-1369): bottleneck adapters after each layer, base parameters frozen
    adapters: bool = False                         # set by the `adapter_*` architectures
    encoder_adapter_dim: int = 64
    decoder_adapter_dim: int = 64
    encoder_adapter_layer_ids: Optional[tuple] = None
    decoder_adapter_layer_ids: Optional[tuple] = None
    encoder_adapters: Optional[tuple] = None       # names of stacked adapters (default: ('default',)); () disables
    decoder_adapters: Optional[tuple] = None
    encoder_adapters_by: tuple = ()                # 'source_lang' / 'target_lang' / 'lang' / 'domain': one adapter per value
    decoder_adapters_by: tuple = ()
    adapter_zero_init: bool = False
    train_all_params: bool = False
    shared_embeddings: bool = True
    tied_output_projection: bool = True
    input_dim: Optional[int] = None
    conv_input_dim: Optional[int] = None
    conv_channels: Optional[int] = None
    conv_kernel_sizes: Optional[tuple] = None
    conv_strides: Optional[tuple] = None
    conv_activation: str = 'glu'
    encoder_positional_encoding: str = 'sinusoidal'
    decoder_positional_encoding: str = 'sinusoidal'
    positional_encoding_shift: int = 2
    encoder_max_len: int = 256
    decoder_max_len: Optional[int] = None          # task-dependent default: 256 for (speech_)translation
    norm_eps: float = 1e-5
    prompt_loss: float = 1.0
    model_type: str = 'encoder_decoder'

    # options of the reference's TransformerConfig (pasero/config.py:1054-1319) that select code outside this path: a
    # reference config object is accepted only while they keep the listed value
    REFERENCE_ONLY = dict(rms_norm=False, has_bias=True, norm_bias=True, encoder_embed_norm=False, decoder_embed_norm=False,
                          shared_norm=False, parallel_attention=False, attention_heads_kv=None, sliding_window=None,
                          lora_rank=0, shift_encoder_layers=None, shift_decoder_layers=None, max_qkv=None,
                          checkpoint_activations=False)

    @classmethod
    def adopt(cls, cfg, task: str = 'speech_translation') -> 'TransformerConfig':
        """This path's view of a model configuration. `cfg` is either one of ours (returned as is) or the reference's own
        `TransformerConfig` / `AdapterTransformerConfig` object as `cli/train.py:220` hands it to a registered model:
        every option this path reads is copied by name; adapters are switched on by the config CLASS in the reference
        (pasero/config.py:1322-1369, models/adapters.py:62-119), i.e. by the presence of `encoder_adapter_dim`."""
        if isinstance(cfg, cls):
            return cfg
        out = cls()
        for f in fields(cls):
            if hasattr(cfg, f.name):
                v = getattr(cfg, f.name)
                setattr(out, f.name, tuple(v) if isinstance(v, list) else v)
        out.adapters = bool(getattr(cfg, 'adapters', hasattr(cfg, 'encoder_adapter_dim')))
        if out.decoder_prenorm is None:
            out.decoder_prenorm = out.encoder_prenorm
        if out.conv_activation is None:
            out.conv_activation = 'glu'
        if out.model_type is None:
            out.model_type = 'encoder_decoder'
        bad = []
        for name, keep in cls.REFERENCE_ONLY.items():
            v = getattr(cfg, name, keep)
            if name == 'attention_heads_kv' and v in (out.encoder_attention_heads, out.decoder_attention_heads):
                continue
            if v != keep and not (keep is None and not v):
                bad.append(f'{name}={v!r}')
        if bad:
            raise NotImplementedError('pasero_b200 does not cover: ' + ', '.join(bad))
        return out.finalize(task)

    def finalize(self, task: str = 'speech_translation') -> 'TransformerConfig':
        if self.label_smoothing is None:
            self.label_smoothing = 0.1 if task in ('translation', 'speech_translation') else 0.0
        if self.decoder_max_len is None:
            self.decoder_max_len = 256
        if self.conv_kernel_sizes is not None:
            self.conv_kernel_sizes = tuple(self.conv_kernel_sizes)
            self.conv_strides = tuple(self.conv_strides) if self.conv_strides else tuple(2 for _ in self.conv_kernel_sizes)
        unsupported = []
        if self.activation_dropout:
            unsupported.append('activation_dropout > 0')
        if self.activation_fn not in ('relu', 'gelu'):
            unsupported.append(f'activation_fn={self.activation_fn}')
        if self.conv_activation not in ('glu', 'gelu'):
            unsupported.append(f'conv_activation={self.conv_activation}')
        for pe in (self.encoder_positional_encoding, self.decoder_positional_encoding):
            if pe not in ('sinusoidal', 'learned'):
                unsupported.append(f'positional encoding {pe}')
        if not self.tied_output_projection:
            unsupported.append('untied output projection')
        if self.prompt_loss != 1.0:
            unsupported.append('prompt_loss != 1')
        if self.model_type != 'encoder_decoder':
            unsupported.append(f'model_type={self.model_type}')
        if self.embed_dim % self.encoder_attention_heads or self.embed_dim // self.encoder_attention_heads != 64 or \
                self.embed_dim // self.decoder_attention_heads != 64:
            unsupported.append('head_dim != 64')
        if unsupported:
            raise NotImplementedError('pasero_b200 (round 1) does not cover: ' + ', '.join(unsupported))
        return self


@dataclass
class DistributedConfig:
    dp_size: int = 1
    tp_size: int = 1
    dp_rank: int = 0
    tp_rank: int = 0
    dtype: str = 'bfloat16'
    seed: Optional[int] = None
    sequence_parallel: bool = True


@dataclass
class TrainingConfig(DistributedConfig):
    task: str = 'speech_translation'
    arch: str = 'transformer'
    model_dir: str = 'models/pasero_b200'
    lr: float = 0.0005
    adam_betas: tuple = (0.9, 0.999)
    adam_eps: float = 1e-8
    warmup: int = 4000
    init_lr: float = 0.0
    min_lr: float = 0.0
    weight_decay: float = 0.0
    clip_norm: float = 1.0
    virtual_dp_size: int = 1
    max_steps: int = 1000
    log_interval: int = 100
    valid_interval: int = 1000
    save_interval: Optional[int] = None
    # task / data options of the ST path (TranslationTaskConfig / dataset configs)
    batch_size: int = 4000
    batch_size_multiple: Optional[int] = None
    lines_per_batch: Optional[int] = None
    max_source_len: Optional[int] = None
    max_target_len: Optional[int] = None
    buffer_size: int = 5000
    # float16 training (pasero/config.py:1537-1550, training.py:289-306): with `dtype: float16` the loss is scaled dynamically
    # (GradScaler); memory_efficient_fp16 (default) = Adam on the 16-bit parameters, otherwise an fp32 master copy
    # (MixedPrecisionAdam, or FlatFP16Adam with flat_fp16 — the same arithmetic on this path's flat arenas)
    flat_fp16: bool = False
    memory_efficient_fp16: bool = True
    bucket_bytes: int = 25 * 2 ** 20      # gradient allreduce bucket (torch DDP default: 25 MB)
    cuda_graphs: bool = False              # capture forward+backward+optimizer per batch shape
    model_cfg: TransformerConfig = field(default_factory=TransformerConfig)
    extra: dict = field(default_factory=dict)

    OUT_OF_SCOPE = ('fsdp', 'amp', 'lora_rank', 'moe_', 'checkpoint_activations')

    # model-config presets selected by `arch` (pasero/config.py: register_model_config). `whisper_base` :2539-2566.
    _BIG = dict(embed_dim=1024, encoder_ffn_dim=4096, decoder_ffn_dim=4096, encoder_attention_heads=16, decoder_attention_heads=16)
    _NLLB_1B3 = dict(_BIG, encoder_layers=24, decoder_layers=24, encoder_ffn_dim=8192, decoder_ffn_dim=8192, encoder_prenorm=True,
                     decoder_prenorm=True)
    ARCH_PRESETS = {
        'transformer': {},
        'adapter_transformer': dict(adapters=True),
        # pasero/config.py:2183-2188 (BASELINE config 5) and :2229-2240, 2499-2501 (the IWSLT2023 recipe's NLLB 1.3B body)
        'transformer_big': dict(_BIG),
        'adapter_transformer_big': dict(_BIG, adapters=True),
        'nllb_1b3': dict(_NLLB_1B3),
        'adapter_nllb_1b3': dict(_NLLB_1B3, adapters=True),
        'whisper_base': dict(encoder_layers=6, decoder_layers=6, embed_dim=512, encoder_ffn_dim=2048, decoder_ffn_dim=2048,
                             encoder_attention_heads=8, decoder_attention_heads=8, encoder_prenorm=True, decoder_prenorm=True,
                             activation_fn='gelu', encoder_positional_encoding='learned', decoder_positional_encoding='learned',
                             positional_encoding_shift=0, scale_embed=False, input_dim=80, conv_input_dim=80, conv_channels=512,
                             conv_kernel_sizes=(3, 3), conv_strides=(1, 2), conv_activation='gelu', encoder_max_len=3000,
                             decoder_max_len=448, attention_key_bias=False),
    }

    @classmethod
    def adopt(cls, cfg) -> 'TrainingConfig':
        """This path's view of a training configuration: `cfg` is one of ours (returned as is) or the reference's own
        `TrainingConfig` (pasero/config.py:1469-1868) as `cli/train.py:287` hands it to the Trainer. Options are copied by
        name; options that switch on a subsystem outside the data-parallel ST step raise."""
        if isinstance(cfg, cls):
            return cfg
        out = cls()
        skip = {'model_cfg', 'extra'}
        for f in fields(cls):
            if f.name not in skip and hasattr(cfg, f.name) and getattr(cfg, f.name) is not None:
                v = getattr(cfg, f.name)
                setattr(out, f.name, tuple(v) if isinstance(v, list) else v)
        for name in ('freeze_params_regex', 'train_params_regex'):
            if getattr(cfg, name, None):
                out.extra[name] = getattr(cfg, name)
        bad = [f'{k}={getattr(cfg, k)!r}' for k in ('fsdp', 'amp') if getattr(cfg, k, False)]
        if getattr(cfg, 'tp_size', 1) != 1:
            bad.append(f'tp_size={cfg.tp_size}')
        if bad:
            raise NotImplementedError('options outside the ST data-parallel path: ' + ', '.join(bad))
        if out.dtype not in ('bfloat16', 'float16'):
            raise NotImplementedError('the sm_100a kernels compute in bfloat16 or float16 (dtype: bfloat16 | float16)')
        out.model_cfg = TransformerConfig.adopt(getattr(cfg, 'model_cfg', None) or TransformerConfig(), out.task)
        out.save_interval = out.save_interval or out.valid_interval
        out.max_source_len = out.max_source_len or out.model_cfg.encoder_max_len
        out.max_target_len = out.max_target_len or out.model_cfg.decoder_max_len
        out._reference = cfg
        return out

    @staticmethod
    def reference_option_names() -> set:
        """Every option name the reference's config classes declare (generated by tests/golden/make_golden.py options)."""
        import json
        import os
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_options.json')) as f:
            return set(json.load(f))

    @classmethod
    def from_argv(cls, argv) -> 'TrainingConfig':
        """The reference's command-line surface (pasero/config.py:1711-1760): `-c/--config FILE.yaml` first, then
        `--option-name value...` overrides (CLI > YAML > defaults); a flag without a value is True, several values make a list."""
        opts, i, argv = {}, 0, list(argv)
        cli = {}
        while i < len(argv):
            tok = argv[i]
            if tok in ('-c', '--config'):
                with open(argv[i + 1]) as f:
                    opts.update(yaml.safe_load(f) or {})
                i += 2
                continue
            if not tok.startswith('--'):
                raise ValueError(f'unexpected argument {tok!r}')
            j = i + 1
            vals = []
            while j < len(argv) and not (argv[j].startswith('--') and not argv[j][2:3].isdigit()) and argv[j] != '-c':
                vals.append(yaml.safe_load(argv[j]))
                j += 1
            cli[tok[2:].replace('-', '_')] = True if not vals else (vals[0] if len(vals) == 1 else vals)
            i = j
        opts.update(cli)
        return cls.from_dict(opts)

    @classmethod
    def from_dict(cls, opts: dict) -> 'TrainingConfig':
        opts = {k.replace('-', '_'): v for k, v in opts.items()}
        model_fields = {f.name for f in fields(TransformerConfig)}
        train_fields = {f.name for f in fields(cls)} - {'model_cfg', 'extra'}
        cfg = cls()
        arch = opts.get('arch', cfg.arch)
        if arch not in cls.ARCH_PRESETS:
            raise NotImplementedError(f'arch {arch!r}: only {sorted(cls.ARCH_PRESETS)} are on the ST hot path')
        for k, v in cls.ARCH_PRESETS[arch].items():
            setattr(cfg.model_cfg, k, v)
        def coerce(v):   # YAML 1.1 reads "1e-4" as a string: numbers written that way still mean numbers
            if isinstance(v, str):
                try:
                    return float(v) if any(c in v for c in '.eE') else int(v)
                except ValueError:
                    return v
            return v
        for k, v in opts.items():
            v = [coerce(x) for x in v] if isinstance(v, list) else coerce(v)
            if k in model_fields:
                setattr(cfg.model_cfg, k, v)
            elif k in train_fields:
                setattr(cfg, k, tuple(v) if isinstance(v, list) else v)
            else:
                cfg.extra[k] = v
        # an option neither this path nor the reference knows is a typo: error, as the reference does (config.py:1748-1750)
        unknown = sorted(set(cfg.extra) - cls.reference_option_names())
        if unknown:
            raise ValueError(f'unknown option(s): {", ".join(unknown)}')
        for k, v in cfg.extra.items():
            if v and any(k.startswith(p) for p in cls.OUT_OF_SCOPE):
                raise NotImplementedError(f'option {k}={v!r} selects a subsystem outside the ST data-parallel path')
        if cfg.tp_size != 1:
            raise NotImplementedError('tensor parallelism is out of scope (pasero/training.py:180-183)')
        if cfg.dtype not in ('bfloat16', 'float16'):
            raise NotImplementedError('the sm_100a kernels compute in bfloat16 or float16 (dtype: bfloat16 | float16)')
        cfg.model_cfg.finalize(cfg.task)
        cfg.save_interval = cfg.save_interval or cfg.valid_interval
        cfg.max_source_len = cfg.max_source_len or cfg.model_cfg.encoder_max_len
        cfg.max_target_len = cfg.max_target_len or cfg.model_cfg.decoder_max_len
        return cfg

    @classmethod
    def from_yaml(cls, path: str, **overrides) -> 'TrainingConfig':
        with open(path) as f:
            opts = yaml.safe_load(f) or {}
        opts.update(overrides)
        return cls.from_dict(opts)

    def as_dict(self) -> dict:
        """Flat dict in the reference's `cfg.as_dict()` style (stored under 'args' in checkpoints)."""
        ref = getattr(self, '_reference', None)
        if ref is not None:       # adopted from the reference's own TrainingConfig: its as_dict() is the authority
            return ref.as_dict()
        out = {f.name: getattr(self, f.name) for f in fields(self) if f.name not in ('model_cfg', 'extra')}
        out.update({f.name: getattr(self.model_cfg, f.name) for f in fields(self.model_cfg)})
        out.update(self.extra)
        return out
