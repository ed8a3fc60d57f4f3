"""ORACLE (test infrastructure, not product code): CPU restatement of Pasero's speech-translation
Transformer forward pass and loss, written functionally over a reference-format `state_dict`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
Plain PyTorch ops on CPU (any dtype; fp32 is the ground truth the bf16 CUDA path is measured against);
gradients come from torch autograd. Each function cites the reference code it restates.

PARITY PIN: tests/golden/st_tiny_*.pt were produced by tests/golden/make_golden.py running the unmodified
reference (imported from /root/reference) on seeded inputs; tests/test_oracle_cpu.py checks this file
against those vectors (loss, logs, encoder output, logits and every parameter gradient).
"""
import dataclasses
import math
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor


@dataclass
class STConfig:
    """The subset of pasero.config.TransformerConfig (config.py:1054-1319) the ST path reads."""
    embed_dim: int = 512
    encoder_layers: int = 12
    decoder_layers: int = 6
    encoder_attention_heads: int = 8
    decoder_attention_heads: int = 8
    encoder_ffn_dim: int = 2048
    decoder_ffn_dim: int = 2048
    input_dim: int = 80
    conv_channels: int = 1024
    conv_kernel_sizes: tuple = (5, 5)
    conv_strides: tuple = (2, 2)
    conv_activation: str = 'glu'              # 'gelu' for the Whisper family (config.py:2559)
    positional_encoding: str = 'sinusoidal'   # 'learned': LearnedPositionalEmbedding (modules.py:460-484)
    encoder_prenorm: bool = False
    decoder_prenorm: bool = False
    activation_fn: str = 'relu'
    scale_embed: bool = True
    positional_encoding_shift: int = 2
    norm_eps: float = 1e-5
    label_smoothing: float = 0.1
    padding_idx: int = 1
    vocab_size: int = 8032
    encoder_max_len: int = 3000
    decoder_max_len: int = 256
    extra: dict = field(default_factory=dict)


def len_to_mask(lengths: Tensor, size: int) -> Tensor:
    """pasero/utils.py:258-268 — True at padding positions."""
    return torch.arange(size, device=lengths.device).unsqueeze(0) >= lengths.unsqueeze(1)


def conv_out_length(length, kernel_sizes, strides):
    """ConvolutionSubsampler.get_new_length, pasero/models/modules.py:804-817."""
    for k, s in zip(kernel_sizes, strides):
        length = 1 + torch.div(length - k + 2 * (k // 2), s, rounding_mode='floor')
    return length


def sinusoidal_table(num_embeddings: int, dim: int, shift: int) -> Tensor:
    """SinusoidalPositionalEmbedding.__init__, pasero/models/modules.py:417-430 (fairseq [sin | cos] layout)."""
    num_embeddings += shift
    half = dim // 2
    w = math.log(10000) / (half - 1)
    w = torch.exp(torch.arange(half, dtype=torch.float) * -w)
    w = torch.arange(num_embeddings, dtype=torch.float).unsqueeze(1) * w.unsqueeze(0)
    w = torch.cat([torch.sin(w), torch.cos(w)], dim=1).view(num_embeddings, -1)
    if dim % 2 == 1:
        w = torch.cat([w, torch.zeros(num_embeddings, 1)], dim=1)
    return w


def positions(table: Tensor, length: int, shift: int, dtype) -> Tensor:
    """SinusoidalPositionalEmbedding.forward, modules.py:435-457: rows shift .. shift+length-1, cast to model dtype."""
    return table.to(dtype)[shift:shift + length].unsqueeze(0)


def attention(sd, prefix: str, query: Tensor, key: Tensor, heads: int, key_padding_mask: Optional[Tensor],
              causal: bool) -> Tensor:
    """MultiheadAttention.forward, pasero/models/modules.py:579-739 (training path, no state, no return_attn):
    separate q/k/v/out projections with bias, scale 1/sqrt(d_h), -inf key-padding mask (:659-677) or causal with
    the padding mask dropped (:602-605), softmax in fp32 inside SDPA."""
    B, T, D = query.shape
    S = key.shape[1]
    dh = D // heads
    q = F.linear(query, sd[prefix + 'q_proj.weight'], sd[prefix + 'q_proj.bias']).view(B, T, heads, dh).transpose(1, 2)
    k = F.linear(key, sd[prefix + 'k_proj.weight'], sd.get(prefix + 'k_proj.bias')).view(B, S, heads, dh).transpose(1, 2)
    v = F.linear(key, sd[prefix + 'v_proj.weight'], sd[prefix + 'v_proj.bias']).view(B, S, heads, dh).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(2, 3)).float() * (1.0 / math.sqrt(dh))
    if causal:
        cm = torch.ones(T, S, dtype=torch.bool, device=query.device).triu(1 + S - T)
        scores = scores.masked_fill(cm, float('-inf'))
    elif key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask[:, None, None, :], float('-inf'))
    probs = torch.softmax(scores, dim=-1).to(v.dtype)
    out = torch.matmul(probs, v).transpose(1, 2).reshape(B, T, D)
    return F.linear(out, sd[prefix + 'out_proj.weight'], sd[prefix + 'out_proj.bias'])


def _act(cfg: STConfig, x: Tensor) -> Tensor:
    return F.gelu(x) if cfg.activation_fn == 'gelu' else F.relu(x)


def _ln(sd, prefix: str, x: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + 'weight'], sd[prefix + 'bias'], eps)


def encoder_layer(cfg: STConfig, sd, prefix: str, x: Tensor, padding_mask: Tensor) -> Tensor:
    """TransformerEncoderLayer.forward, pasero/models/transformer.py:1056-1099 (dropout = 0 for parity)."""
    residual = x
    if cfg.encoder_prenorm:
        x = _ln(sd, prefix + 'self_attn_layer_norm.', x, cfg.norm_eps)
    x = attention(sd, prefix + 'self_attn.', x, x, cfg.encoder_attention_heads, padding_mask, causal=False)
    x = residual + x
    if not cfg.encoder_prenorm:
        x = _ln(sd, prefix + 'self_attn_layer_norm.', x, cfg.norm_eps)
    residual = x
    if cfg.encoder_prenorm:
        x = _ln(sd, prefix + 'final_layer_norm.', x, cfg.norm_eps)
    x = F.linear(_act(cfg, F.linear(x, sd[prefix + 'fc1.weight'], sd[prefix + 'fc1.bias'])),
                 sd[prefix + 'fc2.weight'], sd[prefix + 'fc2.bias'])
    x = residual + x
    if not cfg.encoder_prenorm:
        x = _ln(sd, prefix + 'final_layer_norm.', x, cfg.norm_eps)
    return x


def decoder_layer(cfg: STConfig, sd, prefix: str, x: Tensor, enc: Tensor, enc_mask: Tensor) -> Tensor:
    """TransformerDecoderLayer.forward, pasero/models/transformer.py:1341-1417."""
    pre = cfg.decoder_prenorm
    residual = x
    if pre:
        x = _ln(sd, prefix + 'self_attn_layer_norm.', x, cfg.norm_eps)
    x = attention(sd, prefix + 'self_attn.', x, x, cfg.decoder_attention_heads, None, causal=True)
    x = residual + x
    if not pre:
        x = _ln(sd, prefix + 'self_attn_layer_norm.', x, cfg.norm_eps)
    residual = x
    if pre:
        x = _ln(sd, prefix + 'encoder_attn_layer_norm.', x, cfg.norm_eps)
    x = attention(sd, prefix + 'encoder_attn.', x, enc, cfg.decoder_attention_heads, enc_mask, causal=False)
    x = residual + x
    if not pre:
        x = _ln(sd, prefix + 'encoder_attn_layer_norm.', x, cfg.norm_eps)
    residual = x
    if pre:
        x = _ln(sd, prefix + 'final_layer_norm.', x, cfg.norm_eps)
    x = F.linear(_act(cfg, F.linear(x, sd[prefix + 'fc1.weight'], sd[prefix + 'fc1.bias'])),
                 sd[prefix + 'fc2.weight'], sd[prefix + 'fc2.bias'])
    x = residual + x
    if not pre:
        x = _ln(sd, prefix + 'final_layer_norm.', x, cfg.norm_eps)
    return x


def adapters(cfg: STConfig, sd, prefix: str, x: Tensor, names) -> Tensor:
    """AdapterLayer.forward stacked after a Transformer layer (pasero/models/modules.py:344-357, adapters.py:248-252,
    296-300): x + up(relu(down(LayerNorm(x)))) for every adapter name present in the state dict (eps = nn.LayerNorm default)."""
    for name in names:
        p = f'{prefix}adapters.{name}.'
        if p + 'down.weight' not in sd:
            continue
        h = F.layer_norm(x, (x.shape[-1],), sd[p + 'layer_norm.weight'], sd[p + 'layer_norm.bias'], 1e-5)
        h = F.relu(F.linear(h, sd[p + 'down.weight'], sd[p + 'down.bias']))
        x = x + F.linear(h, sd[p + 'up.weight'], sd[p + 'up.bias'])
    return x


def encode(cfg: STConfig, sd, encoder_input: Tensor, encoder_input_length: Tensor, layer_outputs: Optional[dict] = None):
    """ConvolutionSubsampler.forward (modules.py:819-834) + TransformerEncoder.forward (transformer.py:697-752)."""
    if encoder_input.ndim == 2:     # token ids (`--task translation`): the shared embedding table, transformer.py:728-732
        x = F.embedding(encoder_input.clip(min=0), sd['encoder.embed_tokens.weight'], padding_idx=cfg.padding_idx)
        length = encoder_input_length
        kernel_sizes = strides = ()
    else:
        kernel_sizes, strides = cfg.conv_kernel_sizes, cfg.conv_strides
        x = encoder_input
        if 'encoder.in_linear.0.weight' in sd:      # Linear + ReLU in front of the subsampler, transformer.py:647-651, 735-736
            x = F.relu(F.linear(x, sd['encoder.in_linear.0.weight'], sd['encoder.in_linear.0.bias']))
        x = x.transpose(1, 2)
        for i, (k, s) in enumerate(zip(kernel_sizes, strides)):
            x = F.conv1d(x, sd[f'encoder.subsample.conv_layers.{i}.weight'], sd[f'encoder.subsample.conv_layers.{i}.bias'],
                         stride=s, padding=k // 2)
            x = F.glu(x, dim=1) if cfg.conv_activation == 'glu' else F.gelu(x)   # modules.py:803, 831-832
        x = x.transpose(1, 2)
        length = conv_out_length(encoder_input_length, kernel_sizes, strides)
    if cfg.scale_embed:
        x = x * math.sqrt(cfg.embed_dim)
    mask = len_to_mask(length, x.shape[1])
    max_len = int(conv_out_length(torch.tensor(cfg.encoder_max_len), kernel_sizes, strides))
    if cfg.positional_encoding == 'learned':   # rows shift .. shift + S - 1 of the trainable table
        x = x + sd['encoder.embed_positions.weight'][cfg.positional_encoding_shift:cfg.positional_encoding_shift + x.shape[1]].unsqueeze(0)
    else:
        table = sinusoidal_table(max_len, cfg.embed_dim, cfg.positional_encoding_shift)
        x = x + positions(table, x.shape[1], cfg.positional_encoding_shift, x.dtype).to(x.device)
    if layer_outputs is not None:
        layer_outputs['enc_embed'] = x
    for i in range(cfg.encoder_layers):
        x = encoder_layer(cfg, sd, f'encoder.layers.{i}.', x, mask)
        x = adapters(cfg, sd, f'encoder.layers.{i}.', x, cfg.extra.get('encoder_adapters', ('default',)))
        if layer_outputs is not None:
            layer_outputs[f'enc_{i}'] = x
    if cfg.encoder_prenorm:
        x = _ln(sd, 'encoder.layer_norm.', x, cfg.norm_eps)
    return x, mask


def decode(cfg: STConfig, sd, enc: Tensor, enc_mask: Tensor, decoder_input: Tensor, layer_outputs: Optional[dict] = None):
    """Embedding.forward / projection (modules.py:916-947) + TransformerDecoder.forward (transformer.py:830-898)."""
    emb = sd['decoder.embed_tokens.weight']
    x = F.embedding(decoder_input.clip(min=0), emb, padding_idx=cfg.padding_idx)
    if cfg.scale_embed:
        x = x * math.sqrt(cfg.embed_dim)
    if cfg.positional_encoding == 'learned':
        x = x + sd['decoder.embed_positions.weight'][cfg.positional_encoding_shift:cfg.positional_encoding_shift + x.shape[1]].unsqueeze(0)
    else:
        table = sinusoidal_table(cfg.decoder_max_len, cfg.embed_dim, cfg.positional_encoding_shift)
        x = x + positions(table, x.shape[1], cfg.positional_encoding_shift, x.dtype).to(x.device)
    for i in range(cfg.decoder_layers):
        x = decoder_layer(cfg, sd, f'decoder.layers.{i}.', x, enc, enc_mask)
        x = adapters(cfg, sd, f'decoder.layers.{i}.', x, cfg.extra.get('decoder_adapters', ('default',)))
        if layer_outputs is not None:
            layer_outputs[f'dec_{i}'] = x
    if cfg.decoder_prenorm:
        x = _ln(sd, 'decoder.layer_norm.', x, cfg.norm_eps)
    return torch.matmul(x, emb.T)


def compute_loss(cfg: STConfig, logits: Tensor, target: Tensor):
    """Transformer.compute_loss, pasero/models/transformer.py:323-380: summed label-smoothed CE on fp32 logits,
    a second plain-NLL evaluation for logging, logs in bits."""
    lg = logits.float().view(-1, logits.size(-1))
    tg = target.reshape(-1)
    loss = F.cross_entropy(lg, tg, ignore_index=cfg.padding_idx, reduction='sum', label_smoothing=cfg.label_smoothing)
    with torch.no_grad():
        nll = F.cross_entropy(lg, tg, ignore_index=cfg.padding_idx, reduction='sum')
    logs = {
        'loss': loss.item() / math.log(2),
        'nll_loss': nll.item() / math.log(2),
        'num_tokens': int((tg != cfg.padding_idx).sum().item()),
        'num_lines': target.size(0),
    }
    return loss, logs


def adapters_in_use(cfg: STConfig, side: str, meta: Optional[dict]):
    """AdapterTransformer{Encoder,Decoder}.forward (pasero/models/adapters.py:186-197, 219-229): with `{side}_adapters_by`
    (cfg.extra) the adapters of a batch are named after its language(s) / domain, `lang:<value>` or `domain:<value>`."""
    by = cfg.extra.get(f'{side}_adapters_by') or ()
    if not by:
        return cfg.extra.get(f'{side}_adapters', ('default',))
    return tuple(f"{'domain:' if key == 'domain' else 'lang:'}{meta[key]}" for key in by)


def forward(cfg: STConfig, sd, encoder_input: Tensor, encoder_input_length: Tensor, decoder_input: Tensor,
            layer_outputs: Optional[dict] = None, meta: Optional[dict] = None):
    """Transformer.forward, pasero/models/transformer.py:227-283: shift (target = dec[:, 1:], input = dec[:, :-1])."""
    if cfg.extra.get('encoder_adapters_by') or cfg.extra.get('decoder_adapters_by'):
        cfg = dataclasses.replace(cfg, extra={**cfg.extra, 'encoder_adapters_by': (), 'decoder_adapters_by': (),
                                              'encoder_adapters': adapters_in_use(cfg, 'encoder', meta),
                                              'decoder_adapters': adapters_in_use(cfg, 'decoder', meta)})
    target = decoder_input[:, 1:]
    dec_in = decoder_input[:, :-1]
    enc, mask = encode(cfg, sd, encoder_input, encoder_input_length, layer_outputs)
    logits = decode(cfg, sd, enc, mask, dec_in, layer_outputs)
    if layer_outputs is not None:
        layer_outputs['encoder_out'] = enc
        layer_outputs['logits'] = logits
    return compute_loss(cfg, logits, target)
