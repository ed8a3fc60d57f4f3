"""The reference's encoder-decoder Transformer for the speech-translation task (pasero/models/transformer.py),
re-built on the sm_100a kernels. Same constructor contract `(cfg, dist_cfg, task)`, same
`forward(**batch) -> (loss, logs)`, `encoder(...)` / `decoder(...)` / `compute_loss(...)` entry points and
state_dict keys, so it drops in behind the reference's model registry for this path.
"""
import math
import os
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import modules, ops
from .config import DistributedConfig, TransformerConfig
from .modules import len_to_mask

BF16 = torch.bfloat16


def defined(*args):
    return next((x for x in args if x is not None), None)


_LANG_PREFIX, _DOMAIN_PREFIX = 'lang:', 'domain:'


def _adapter_names(cfg: TransformerConfig, task, side: str) -> tuple:
    """adapters.py:67-105: explicit names, or one adapter per language / domain of the task, or ('default',)."""
    explicit = getattr(cfg, f'{side}_adapters')
    by = getattr(cfg, f'{side}_adapters_by')
    if explicit is not None:
        names = list(explicit)
    elif by:
        names = []
        for key in by:
            prefix = _DOMAIN_PREFIX if key == 'domain' else _LANG_PREFIX
            names += [f'{prefix}{v}' for v in sorted(task.get_langs_or_domains(key))]
    else:
        names = ['default']
    return tuple(dict.fromkeys(names))


class _AdapterMixin:
    """AdapterTransformer{Encoder,Decoder}Layer of the reference (adapters.py:232-301): adapters applied after the layer."""

    def init_adapters(self, cfg, dim: int, names, enabled: bool) -> None:
        self.adapters = nn.ModuleDict({uid: modules.AdapterLayer(cfg.embed_dim, dim, cfg.adapter_zero_init) for uid in names}) \
            if enabled else None
        self.adapters_in_use = tuple(names) if enabled else ()

    def apply_adapters(self, x: Tensor) -> Tensor:
        for uid in self.adapters_in_use:
            x = self.adapters[uid](x)
        return x


class TransformerEncoderLayer(nn.Module, _AdapterMixin):
    """pasero/models/transformer.py:913-1099"""

    def __init__(self, cfg: TransformerConfig, dist_cfg: DistributedConfig, layer_id: int, adapter_names=()):
        super().__init__()
        self.cfg, self.layer_id = cfg, layer_id
        self.init_adapters(cfg, cfg.encoder_adapter_dim, adapter_names, bool(
            cfg.adapters and cfg.encoder_adapter_dim and adapter_names and
            (cfg.encoder_adapter_layer_ids is None or layer_id in cfg.encoder_adapter_layer_ids)))
        self.self_attn = modules.MultiheadAttention(cfg.embed_dim, cfg.encoder_attention_heads, dropout=cfg.attention_dropout,
                                                    scaled=cfg.scale_attn, key_bias=cfg.attention_key_bias)
        self.self_attn_layer_norm = nn.LayerNorm(cfg.embed_dim, eps=cfg.norm_eps)
        self.fc1 = modules.Linear(cfg.embed_dim, cfg.encoder_ffn_dim)
        self.fc1.bias._pb200_f32_grad = True   # see ops._wgrad_bias
        self.fc2 = modules.Linear(cfg.encoder_ffn_dim, cfg.embed_dim)
        self.final_layer_norm = nn.LayerNorm(cfg.embed_dim, eps=cfg.norm_eps)
        self.prenorm = cfg.encoder_prenorm
        self.dropout_p = cfg.dropout
        self.name = f'enc_{layer_id}'

    def forward(self, x: Tensor, padding_mask: Optional[Tensor], /, return_layers=(), key_lengths: Optional[Tensor] = None):
        p = self.dropout_p if self.training else 0.0
        eps = self.cfg.norm_eps
        ln1, ln2 = self.self_attn_layer_norm, self.final_layer_norm
        if not self.prenorm and self.cfg.activation_fn == 'relu':   # post-norm ReLU (the ST default): fused autograd nodes
            x = ops.attn_sublayer(x, None, self.self_attn, ln1, key_lengths if key_lengths is not None else
                                  self.self_attn.key_lengths_of(padding_mask), False, p, eps)
            x = ops.ffn_sublayer(x, self.fc1, self.fc2, ln2, p, eps)
            return self.apply_adapters(x), ({self.name: x} if self.name in return_layers else {})
        residual = x
        if self.prenorm:
            x = ops.dropout_add_ln(x, None, ln1.weight, ln1.bias, 0.0, eps)
        x, _ = self.self_attn(x, x, x, attn_mask=padding_mask, key_lengths=key_lengths)
        if self.prenorm:
            x = ops.dropout_add_ln(x, residual, None, None, p, eps)
        else:
            x = ops.dropout_add_ln(x, residual, ln1.weight, ln1.bias, p, eps)
        residual = x
        if self.prenorm:
            x = ops.dropout_add_ln(x, None, ln2.weight, ln2.bias, 0.0, eps)
        x = ops.ffn(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.cfg.activation_fn)
        if self.prenorm:
            x = ops.dropout_add_ln(x, residual, None, None, p, eps)
        else:
            x = ops.dropout_add_ln(x, residual, ln2.weight, ln2.bias, p, eps)
        return self.apply_adapters(x), ({self.name: x} if self.name in return_layers else {})


class TransformerDecoderLayer(nn.Module, _AdapterMixin):
    """pasero/models/transformer.py:1102-1417"""

    def __init__(self, cfg: TransformerConfig, dist_cfg: DistributedConfig, layer_id: int, adapter_names=()):
        super().__init__()
        self.cfg, self.layer_id = cfg, layer_id
        self.init_adapters(cfg, cfg.decoder_adapter_dim, adapter_names, bool(
            cfg.adapters and cfg.decoder_adapter_dim and adapter_names and
            (cfg.decoder_adapter_layer_ids is None or layer_id in cfg.decoder_adapter_layer_ids)))
        self.self_attn = modules.MultiheadAttention(cfg.embed_dim, cfg.decoder_attention_heads, dropout=cfg.attention_dropout,
                                                    causal=True, scaled=cfg.scale_attn, key_bias=cfg.attention_key_bias)
        self.self_attn_layer_norm = nn.LayerNorm(cfg.embed_dim, eps=cfg.norm_eps)
        self.encoder_attn = modules.MultiheadAttention(cfg.embed_dim, cfg.decoder_attention_heads,
                                                       dropout=cfg.attention_dropout, scaled=cfg.scale_attn,
                                                       key_bias=cfg.attention_key_bias)
        self.encoder_attn_layer_norm = nn.LayerNorm(cfg.embed_dim, eps=cfg.norm_eps)
        self.fc1 = modules.Linear(cfg.embed_dim, cfg.decoder_ffn_dim)
        self.fc1.bias._pb200_f32_grad = True   # see ops._wgrad_bias
        self.fc2 = modules.Linear(cfg.decoder_ffn_dim, cfg.embed_dim)
        self.final_layer_norm = nn.LayerNorm(cfg.embed_dim, eps=cfg.norm_eps)
        self.prenorm = cfg.decoder_prenorm
        self.dropout_p = defined(cfg.decoder_dropout, cfg.dropout)
        self.name = f'dec_{layer_id}'

    def _block(self, x, ln, fn, p, eps):
        residual = x
        if self.prenorm:
            x = ops.dropout_add_ln(x, None, ln.weight, ln.bias, 0.0, eps)
        x = fn(x)
        if self.prenorm:
            return ops.dropout_add_ln(x, residual, None, None, p, eps)
        return ops.dropout_add_ln(x, residual, ln.weight, ln.bias, p, eps)

    @property
    def fused_sublayers(self) -> bool:
        """Post-norm + ReLU: every sub-block is one autograd node (ops.AttnSublayerFn / FFNSublayerFn)."""
        return not self.prenorm and self.cfg.activation_fn == 'relu'

    def forward(self, x: Tensor, encoder_out: Tensor, encoder_mask: Optional[Tensor], padding_mask: Optional[Tensor], /,
                return_layers=(), state=None, encoder_lengths: Optional[Tensor] = None, encoder_grad_cell=None, **kwargs):
        if state is not None:
            raise NotImplementedError('incremental decoding is outside the training path')
        p = self.dropout_p if self.training else 0.0
        eps = self.cfg.norm_eps
        if self.fused_sublayers:
            x = ops.attn_sublayer(x, None, self.self_attn, self.self_attn_layer_norm, None, True, p, eps)
            enc_len = encoder_lengths if encoder_lengths is not None else self.encoder_attn.key_lengths_of(encoder_mask)
            x = ops.attn_sublayer(x, encoder_out, self.encoder_attn, self.encoder_attn_layer_norm, enc_len, False, p, eps,
                                  key_grad_cell=encoder_grad_cell)
            x = ops.ffn_sublayer(x, self.fc1, self.fc2, self.final_layer_norm, p, eps)
            return self.apply_adapters(x), ({self.name: x} if self.name in return_layers else {})
        x = self._block(x, self.self_attn_layer_norm, lambda h: self.self_attn(h, h, h)[0], p, eps)
        x = self._block(x, self.encoder_attn_layer_norm,
                        lambda h: self.encoder_attn(h, encoder_out, encoder_out, attn_mask=encoder_mask,
                                                    key_lengths=encoder_lengths)[0], p, eps)
        x = self._block(x, self.final_layer_norm,
                        lambda h: ops.ffn(h, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                          self.cfg.activation_fn), p, eps)
        return self.apply_adapters(x), ({self.name: x} if self.name in return_layers else {})


def _select_adapters(layers, by: tuple, meta: Optional[dict]) -> None:
    """AdapterTransformer{Encoder,Decoder}.forward (adapters.py:186-197, 219-229): with `*_adapters_by`, the adapters of
    the batch's language / domain are the ones in use."""
    if not by:
        return
    names = tuple(f"{_DOMAIN_PREFIX if key == 'domain' else _LANG_PREFIX}{(meta or {})[key]}" for key in by)
    for layer in layers:
        if layer.adapters is not None:
            layer.adapters_in_use = names


class TransformerEncoder(nn.Module):
    """pasero/models/transformer.py:610-765 (speech features: conv subsampler + sinusoidal positions)."""

    def __init__(self, cfg: TransformerConfig, dist_cfg: DistributedConfig, task=None, embed=None):
        super().__init__()
        self.cfg = cfg
        self.padding_idx = getattr(task, 'padding_idx', 1)
        num_embeddings = getattr(task, 'encoder_num_embeddings', 0)
        # SpeechTranslationTask.encoder_num_embeddings == 0 (speech_translation.py:29-31): features, no embedding table.
        # Text tasks (`--task translation`): token ids through an embedding table, shared with the decoder when
        # `shared_embeddings` (transformer.py:150-176, 624-640)
        if num_embeddings > 0:
            self.embed_tokens = embed if embed is not None else modules.Embedding(num_embeddings, cfg.embed_dim, self.padding_idx)
            self.subsample = self.in_linear = None
            encoder_max_len = cfg.encoder_max_len
        else:
            self.embed_tokens = None
            input_dim = cfg.input_dim or cfg.embed_dim
            if not cfg.conv_kernel_sizes:
                raise NotImplementedError('feature inputs expect a conv subsampler (conv_kernel_sizes)')
            conv_input_dim = cfg.conv_input_dim or input_dim
            # in_linear (transformer.py:647-651): Linear(input_dim -> conv_input_dim) + ReLU in front of the subsampler,
            # e.g. 768-d wav2vec features -> 80 in the IWSLT2023 recipe; state_dict keys in_linear.0.{weight,bias}
            self.in_linear = nn.ModuleList([modules.Linear(input_dim, conv_input_dim)]) if conv_input_dim != input_dim else None
            self.subsample = modules.ConvolutionSubsampler(conv_input_dim, cfg.conv_channels or conv_input_dim, cfg.embed_dim,
                                                           cfg.conv_kernel_sizes, cfg.conv_strides, cfg.conv_activation)
            encoder_max_len = int(self.subsample.get_new_length(torch.tensor(cfg.encoder_max_len)))
        self.embed_positions = modules.positional_embedding(cfg.encoder_positional_encoding, encoder_max_len, cfg.embed_dim,
                                                            cfg.positional_encoding_shift)
        self.embed_scale = math.sqrt(cfg.embed_dim) if cfg.scale_embed else 1.0
        self.dropout_p = defined(cfg.embed_dropout, cfg.dropout)
        self.adapter_names = _adapter_names(cfg, task, 'encoder') if cfg.adapters else ()
        self.layers = nn.ModuleList([TransformerEncoderLayer(cfg, dist_cfg, i, self.adapter_names)
                                     for i in range(cfg.encoder_layers)])
        self.layer_norm = nn.LayerNorm(cfg.embed_dim, eps=cfg.norm_eps) if cfg.encoder_prenorm else None

    def forward(self, encoder_input: Tensor, encoder_input_length: Tensor, return_layers=(), meta: dict = {},
                padded_layout: Optional[tuple] = None, **kwargs):
        """encoder_input (B, S, 80) bf16 features (or, with padded_layout=(frames,), the zero-padded layout written
        by the fbank kernel). Returns (encoder_out (B, S', D), padding_mask (B, S') bool True = pad, layer_outputs)."""
        return_layers = return_layers or ()
        _select_adapters(self.layers, self.cfg.encoder_adapters_by if self.cfg.adapters else (), meta)
        B = encoder_input.shape[0]
        D = self.cfg.embed_dim
        p = self.dropout_p if self.training else 0.0
        learned = self.embed_positions.weight if isinstance(self.embed_positions, modules.LearnedPositionalEmbedding) else None
        if encoder_input.ndim == 2:      # token ids (text tasks)
            if self.embed_tokens is None:
                raise RuntimeError('token-id encoder input needs a task with encoder_num_embeddings > 0')
            S = encoder_input.shape[1]
            length = encoder_input_length
            # this is the last use of the (possibly shared) table in backward order: it reports the parameter ready
            x = ops.EmbedFn.apply(self.embed_tokens.weight, encoder_input, self.embed_positions.table(S), float(self.embed_scale),
                                  float(p), S, self.padding_idx, learned, self.embed_positions.shift, True).view(B, S, D)
        else:
            x = encoder_input
            dt = next(self.parameters()).dtype      # the model's 16-bit type (bfloat16, or float16 with `dtype: float16`)
            if x.dtype != dt:
                x = x.to(dt)
            if self.in_linear is not None:
                if padded_layout is not None:
                    raise NotImplementedError('the fused fbank layout feeds the subsampler directly (no in_linear)')
                lin = self.in_linear[0]
                x = ops.LinearReluFn.apply(x.reshape(-1, x.shape[-1]), lin.weight, lin.bias).view(B, x.shape[1], -1)
            if padded_layout is not None:
                x, length = self.subsample.forward_padded(x, padded_layout[0], encoder_input_length)
            else:
                x, length = self.subsample(x, encoder_input_length)
            S = x.numel() // (B * D)
            pos = self.embed_positions.table(S)
            x = ops.EmbedFn.apply(x.reshape(B * S, D), None, pos, float(self.embed_scale), float(p), S, -1, learned,
                                  self.embed_positions.shift).view(B, S, D)
        padding_mask = len_to_mask(length, size=S)
        key_lengths = length.clamp(max=S).to(torch.int32)
        layer_outputs = {}
        for layer in self.layers:
            x, out = layer(x, padding_mask, return_layers=return_layers, key_lengths=key_lengths)
            layer_outputs.update(out)
        if self.layer_norm is not None:
            x = ops.dropout_add_ln(x, None, self.layer_norm.weight, self.layer_norm.bias, 0.0, self.cfg.norm_eps)
        self._last_lengths = key_lengths
        return x, padding_mask, layer_outputs


class TransformerDecoder(nn.Module):
    """pasero/models/transformer.py:767-910"""

    def __init__(self, cfg: TransformerConfig, dist_cfg: DistributedConfig, task=None, embed=None):
        super().__init__()
        self.cfg = cfg
        self.task = task
        self.padding_idx = task.padding_idx
        self.embed_tokens = embed if embed is not None else modules.Embedding(task.decoder_num_embeddings, cfg.embed_dim,
                                                                              self.padding_idx)
        self.embed_positions = modules.positional_embedding(cfg.decoder_positional_encoding, cfg.decoder_max_len, cfg.embed_dim,
                                                            cfg.positional_encoding_shift)
        self.embed_scale = math.sqrt(cfg.embed_dim) if cfg.scale_embed else 1.0
        self.dropout_p = defined(cfg.embed_dropout, cfg.decoder_dropout, cfg.dropout)
        self.adapter_names = _adapter_names(cfg, task, 'decoder') if cfg.adapters else ()
        self.layers = nn.ModuleList([TransformerDecoderLayer(cfg, dist_cfg, i, self.adapter_names)
                                     for i in range(cfg.decoder_layers)])
        self.layer_norm = nn.LayerNorm(cfg.embed_dim, eps=cfg.norm_eps) if cfg.decoder_prenorm else None
        self.output_projection = None  # tied (cfg.tied_output_projection)
        self.table_shared_with_encoder = False   # set by Transformer when the encoder embeds tokens with the same table

    # what a search loop asks a decoder (pasero/models/transformer.py: Decoder.bos_idx / eos_idx / max_len)
    @property
    def eos_idx(self) -> int:
        return self.task.eos_idx

    @property
    def bos_idx(self) -> int:
        return getattr(self.task, 'bos_idx', self.task.eos_idx)

    @property
    def max_len(self) -> int:
        return self.cfg.decoder_max_len

    def features(self, encoder_out, encoder_mask, decoder_input, return_layers=(), encoder_lengths=None, meta=None):
        _select_adapters(self.layers, self.cfg.decoder_adapters_by if self.cfg.adapters else (), meta)
        B, T = decoder_input.shape
        D = self.cfg.embed_dim
        p = self.dropout_p if self.training else 0.0
        pos = self.embed_positions.table(T)
        learned = self.embed_positions.weight if isinstance(self.embed_positions, modules.LearnedPositionalEmbedding) else None
        x = ops.EmbedFn.apply(self.embed_tokens.weight, decoder_input, pos, float(self.embed_scale), float(p), T,
                              self.padding_idx, learned, self.embed_positions.shift, not self.table_shared_with_encoder).view(B, T, D)
        layer_outputs = {}
        # every layer's cross-attention reads the encoder output: PB200_GRAD_CELL=1 sums their gradients for it in the dgrad
        # GEMMs' epilogues (ops.GradCell) instead of by one autograd add kernel per extra layer. Off by default: measured on
        # config 2 the five saved adds (~55 us) are paid back by the slower residual epilogue of five 16 000 x 512 x 1 024 GEMMs
        # (step 8.77 -> 8.71 - 8.73 ms, inside the run-to-run noise; GEMM family 5.28 -> 5.44 ms)
        fused = sum(1 for layer in self.layers if layer.fused_sublayers)
        cell = ops.GradCell(fused) if (fused > 1 and torch.is_grad_enabled() and encoder_out.requires_grad and
                                       os.environ.get('PB200_GRAD_CELL', '0') == '1') else None
        for layer in self.layers:
            x, out = layer(x, encoder_out, encoder_mask, None, return_layers=return_layers, encoder_lengths=encoder_lengths,
                           encoder_grad_cell=cell)
            layer_outputs.update(out)
        if self.layer_norm is not None:
            x = ops.dropout_add_ln(x, None, self.layer_norm.weight, self.layer_norm.bias, 0.0, self.cfg.norm_eps)
        return x, layer_outputs

    def forward(self, encoder_out: Tensor, encoder_mask: Tensor, decoder_input: Tensor, prompt_mask=None, state=None,
                return_layers=(), meta: dict = {}, encoder_lengths=None, **kwargs):
        """Returns (logits (B, T, V), layer_outputs). `state` (a dict, empty at the first call): incremental decoding —
        `decoder_input` then holds only the newest tokens and the cached keys / values live in `state` (pasero_b200.decoding)."""
        if state is not None:
            from . import decoding
            return decoding.incremental_forward(self, encoder_out, encoder_mask, decoder_input, state, encoder_lengths, meta)
        x, layer_outputs = self.features(encoder_out, encoder_mask, decoder_input, return_layers or (), encoder_lengths, meta)
        return self.embed_tokens.projection(x), layer_outputs


class Transformer(nn.Module):
    """pasero/models/transformer.py:106-608 (`register_model('transformer')`), ST path."""

    def __init__(self, cfg: TransformerConfig, dist_cfg: DistributedConfig, task):
        super().__init__()
        # `cfg`: ours, or the reference's TransformerConfig / AdapterTransformerConfig when built through the reference's
        # registry (`get_architecture(cfg.model_cfg)(cfg.model_cfg, cfg, task)`, cli/train.py:220)
        self.reference_cfg = cfg
        cfg = TransformerConfig.adopt(cfg, getattr(dist_cfg, 'task', 'speech_translation'))
        self.cfg, self.dist_cfg, self.task = cfg, dist_cfg, task
        self.find_unused_parameters = False
        self.batch_by = None
        if getattr(task, 'encoder_num_embeddings', 0) > 0 and cfg.shared_embeddings:   # transformer.py:150-176
            self.decoder = self.build_decoder(embed=None)
            self.encoder = self.build_encoder(embed=self.decoder.embed_tokens)
            self.decoder.table_shared_with_encoder = True
        else:
            self.encoder = self.build_encoder()
            self.decoder = self.build_decoder(embed=None)
        if cfg.adapters:   # AdapterTransformer (adapters.py:62-119)
            by = sorted(set(cfg.encoder_adapters_by) | set(cfg.decoder_adapters_by))
            self.batch_by = by or None
            self.find_unused_parameters = bool(by)
            if not cfg.train_all_params:
                for name, param in self.named_parameters():
                    if 'adapters' not in name.split('.'):
                        param.requires_grad = False
        self.defer_logs = False   # True: logs hold device scalars (no host sync inside forward)
        self._stats = None

    shard_count, shard_id, is_sharded = 1, 0, False

    @property
    def padding_idx(self) -> int:
        return self.task.padding_idx

    @property
    def total_param_count(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def build_encoder(self, embed=None):
        return TransformerEncoder(self.cfg, self.dist_cfg, task=self.task, embed=embed)

    def build_decoder(self, embed=None):
        return TransformerDecoder(self.cfg, self.dist_cfg, task=self.task, embed=embed)

    def set_ddp_params_and_buffers_to_ignore(self) -> None:
        pass

    def stats_buffer(self, device) -> Tensor:
        """fp64[8] device accumulator shared with the optimizer kernels: loss, nll, num_tokens, num_lines, gnorm^2."""
        if self._stats is None or self._stats.device != device:
            self._stats = torch.zeros(8, dtype=torch.float64, device=device)
        return self._stats

    def forward(self, encoder_input: Tensor = None, encoder_input_length: Tensor = None, decoder_input: Tensor = None,
                prompt_mask: Optional[Tensor] = None, **kwargs):
        """Returns (loss, logs) like the reference (transformer.py:227-283): summed label-smoothed CE and
        {'loss', 'nll_loss'} in bits, 'num_tokens', 'num_lines'."""
        target = decoder_input[:, 1:]
        dec_in = decoder_input[:, :-1]
        enc_kwargs = {k: v for k, v in kwargs.items() if k in ('return_layers', 'meta', 'padded_layout')}
        encoder_out, encoder_mask, _ = self.encoder(encoder_input, encoder_input_length, **enc_kwargs)
        x, _ = self.decoder.features(encoder_out, encoder_mask, dec_in, (), self.encoder._last_lengths, kwargs.get('meta'))
        return self.compute_loss(x, target, {}, fused=True)

    def compute_loss(self, logits_or_features: Tensor, target: Tensor, layer_outputs: dict, *args, fused: bool = False, **kwargs):
        """With fused=True the first argument holds decoder features (B, T, D) and the tied projection runs inside
        the loss op; otherwise it holds logits (B, T, V) as in the reference (transformer.py:323-380)."""
        B, T = target.shape
        dev = target.device
        stats = self.stats_buffer(dev)
        before = stats[:3].clone()
        eps = float(self.cfg.label_smoothing or 0.0)
        if fused:
            x = logits_or_features.reshape(B * T, -1)
            emb = self.decoder.embed_tokens
            loss = ops.ProjCEFn.apply(x, emb.weight, target, self.padding_idx, eps, stats, emb.vocab_size)
        else:
            loss = LogitsCEFn.apply(logits_or_features.reshape(B * T, -1), target, self.padding_idx, eps, stats)
        delta = stats[:3] - before
        stats[3] += B
        if self.defer_logs:
            logs = {'loss': delta[0] / math.log(2), 'nll_loss': delta[1] / math.log(2), 'num_tokens': delta[2], 'num_lines': B}
        else:
            d = delta.tolist()
            logs = {'loss': d[0] / math.log(2), 'nll_loss': d[1] / math.log(2), 'num_tokens': int(d[2]), 'num_lines': B}
        return loss, logs


class LogitsCEFn(torch.autograd.Function):
    """Label-smoothed CE on materialised logits (reference-style compute_loss entry point)."""

    @staticmethod
    def forward(ctx, logits, target, padding_idx, label_smoothing, stats):
        from . import kernels as K
        grad = logits.detach().clone().contiguous()
        before = stats[0].clone()
        K.ls_ce_fwd_bwd(grad, target, padding_idx, label_smoothing, stats, write_grad=True)
        ctx.save_for_backward(grad)
        return (stats[0] - before).float()

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return grad * dloss.to(grad.dtype), None, None, None, None
