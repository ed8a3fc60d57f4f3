"""Host-side mirrors of the reference's building blocks (pasero/models/modules.py) for the ST path, backed by
the sm_100a kernels. Constructor arguments, forward signatures and state_dict keys follow the reference; the
parameter *storage* is laid out for the kernels (fused k|v|q projection, tap-major conv weights) and converted
in the state_dict hooks, so checkpoints are interchangeable with the reference's.
"""
import math
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import kernels as K
from . import ops

BF16 = torch.bfloat16


def len_to_mask(lengths: Tensor, size: Optional[int] = None) -> Tensor:
    """pasero/utils.py:258-268 — boolean mask, True at padding positions."""
    size = size or int(lengths.max())
    return torch.arange(size, device=lengths.device).unsqueeze(0) >= lengths.unsqueeze(1)


def mask_to_len(mask: Tensor, dim: int = -1) -> Tensor:
    """pasero/utils.py:246-255"""
    return (mask.flip(dims=(dim,)).cumsum(dim) > 0).sum(dim)


class Linear(nn.Module):
    """Parameter holder with nn.Linear's names and default initialisation (pasero/models/modules.py:67-100)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1 / math.sqrt(in_features)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        return ops.linear(x, self.weight, self.bias)


class SinusoidalPositionalEmbedding(nn.Module):
    """pasero/models/modules.py:415-457 (fairseq layout [sin | cos], rows offset by `shift`)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, shift: int = 0):
        super().__init__()
        self.shift = shift
        num_embeddings += shift
        self.embedding_dim = embedding_dim
        half_dim = embedding_dim // 2
        weight = math.log(10000) / (half_dim - 1)
        weight = torch.exp(torch.arange(half_dim, dtype=torch.float) * -weight)
        weight = torch.arange(num_embeddings, dtype=torch.float).unsqueeze(1) * weight.unsqueeze(0)
        weight = torch.cat([torch.sin(weight), torch.cos(weight)], dim=1).view(num_embeddings, -1)
        if embedding_dim % 2 == 1:
            weight = torch.cat([weight, torch.zeros(num_embeddings, 1)], dim=1)
        self.weight = weight
        self.register_buffer('_float_tensor', torch.FloatTensor(1))
        self._cache = None

    def table(self, length: int, offset: int = 0) -> Tensor:
        """rows shift+offset .. shift+offset+length-1 as a contiguous [length, D] tensor in the model dtype/device (`offset`:
        positions already consumed, incremental decoding — modules.py:449-457)"""
        assert offset + length + self.shift - 1 < self.weight.size(0), (
            f'input sequence is too long: {offset + length}, positional embedding size: {self.weight.size(0)}')
        ref = self._float_tensor
        if self._cache is None or self._cache.device != ref.device or self._cache.dtype != ref.dtype:
            self._cache = self.weight.to(ref)     # built in fp32, cast to the model dtype (modules.py:447)
        return self._cache[self.shift + offset:self.shift + offset + length]

    def forward(self, length: int, offset: int = 0) -> Tensor:
        return self.table(length, offset).unsqueeze(0).detach()


class LearnedPositionalEmbedding(nn.Module):
    """pasero/models/modules.py:460-484: a trainable [num_embeddings + shift, D] table, row shift + t for position t
    (same rows for every batch element on the training path)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, shift: int = 0):
        super().__init__()
        self.shift = shift
        self.weight = nn.Parameter(torch.empty(num_embeddings + shift, embedding_dim))
        nn.init.normal_(self.weight, mean=0, std=embedding_dim ** -0.5)   # utils.embed_init
        self.weight._pb200_bf16_grad = True   # gradient = batch sum written in bf16 (not a GEMM weight gradient)

    def table(self, length: int, offset: int = 0) -> Tensor:
        assert offset + length + self.shift - 1 < self.weight.size(0), (
            f'input sequence is too long: {offset + length}, positional embedding size: {self.weight.size(0)}')
        return self.weight[self.shift + offset:self.shift + offset + length]

    def forward(self, length: int, offset: int = 0) -> Tensor:
        return self.table(length, offset).unsqueeze(0)


def positional_embedding(kind: str, num_embeddings: int, embedding_dim: int, shift: int):
    if kind == 'learned':
        return LearnedPositionalEmbedding(num_embeddings, embedding_dim, shift)
    return SinusoidalPositionalEmbedding(num_embeddings, embedding_dim, shift)


class AdapterLayer(nn.Module):
    """pasero/models/modules.py:248-357 in its Bapna et al. configuration (LayerNorm, biases, ReLU, residual, scaling 1):
    x + up(relu(down(LayerNorm(x)))). Built from the same kernels as a feed-forward block: LayerNorm, the fused
    bias+ReLU / ReLU-mask GEMM pair, and the residual add."""

    def __init__(self, input_dim: int, projection_dim: int, zero_init: bool = False):
        super().__init__()
        if projection_dim % 8:
            raise NotImplementedError('adapter bottleneck dimension must be a multiple of 8')
        self.down = Linear(input_dim, projection_dim)
        self.up = Linear(projection_dim, input_dim)
        self.layer_norm = nn.LayerNorm(input_dim)
        if zero_init:
            nn.init.kaiming_uniform_(self.down.weight, a=math.sqrt(5))
            nn.init.zeros_(self.up.weight)
        else:
            nn.init.uniform_(self.down.weight, -1e-6, 1e-6)
            nn.init.uniform_(self.up.weight, -1e-6, 1e-6)
        nn.init.zeros_(self.down.bias)
        nn.init.zeros_(self.up.bias)

    def forward(self, x: Tensor) -> Tensor:
        ln = self.layer_norm
        h = ops.dropout_add_ln(x, None, ln.weight, ln.bias, 0.0, ln.eps)
        h = ops.ffn(h, self.down.weight, self.down.bias, self.up.weight, self.up.bias)
        return ops.dropout_add_ln(h, x, None, None, 0.0, ln.eps)     # residual + h


class MultiheadAttention(nn.Module):
    """pasero/models/modules.py:487-771, training path: separate q/k/v/out projections with bias, 1/sqrt(d_h)
    scaling, key-padding mask given as (B, S) bool (True = masked) or causal masking.

    Storage: `in_proj_weight` [3D, D] with rows (k | v | q) and `in_proj_bias` [3D], so self-attention needs one
    projection GEMM and cross-attention two; state_dict exposes the reference's k_proj/v_proj/q_proj/out_proj keys.
    """

    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0, causal: bool = False, scaled: bool = True,
                 key_bias: bool = True):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout, self.causal, self.scaled = embed_dim, num_heads, dropout, causal, scaled
        # key_bias=False (Whisper family): the k rows of the fused bias stay zero, are not in the state_dict and get no
        # gradient (a key bias shifts every score of a query by the same amount: softmax ignores it)
        self.key_bias = key_bias
        self.head_dim = embed_dim // num_heads
        if self.head_dim != 64:
            raise NotImplementedError('pasero_b200 attention kernels are specialised for head_dim 64')
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.in_proj_bias._pb200_f32_grad = True   # gradient accumulated in fp32 by the weight-gradient GEMM (ops._wgrad_bias)
        self.out_proj = Linear(embed_dim, embed_dim)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        D = self.embed_dim
        for i in range(3):  # k, v, q: xavier_uniform with gain 1/sqrt(2) (modules.py:565-573)
            nn.init.xavier_uniform_(self.in_proj_weight.data[i * D:(i + 1) * D], gain=1 / math.sqrt(2))
        nn.init.uniform_(self.in_proj_bias, -1 / math.sqrt(D), 1 / math.sqrt(D))
        if not self.key_bias:
            nn.init.constant_(self.in_proj_bias.data[:D], 0.0)
        nn.init.xavier_uniform_(self.out_proj.weight, gain=1.0)
        nn.init.constant_(self.out_proj.bias, 0.0)

    # -- reference-named views ---------------------------------------------------------------------
    def _block(self, i: int):
        D = self.embed_dim
        return self.in_proj_weight[i * D:(i + 1) * D], self.in_proj_bias[i * D:(i + 1) * D]

    _ORDER = {'k_proj': 0, 'v_proj': 1, 'q_proj': 2}

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name, i in self._ORDER.items():
            w, b = self._block(i)
            destination[prefix + name + '.weight'] = w if keep_vars else w.detach()
            if name != 'k_proj' or self.key_bias:
                destination[prefix + name + '.bias'] = b if keep_vars else b.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for name, i in self._ORDER.items():
            for kind, dst in zip(('weight', 'bias'), self._block(i)):
                if kind == 'bias' and name == 'k_proj' and not self.key_bias:
                    continue
                key = prefix + name + '.' + kind
                if key in state_dict:
                    with torch.no_grad():
                        dst.copy_(state_dict[key])
                elif strict:
                    missing_keys.append(key)

    @staticmethod
    def key_lengths_of(attn_mask: Optional[Tensor]) -> Optional[Tensor]:
        """(B, S) bool key-padding mask (True = masked, padding at the end of each row) -> int32 (B,) key counts"""
        if attn_mask is None:
            return None
        if attn_mask.dim() != 2:
            raise NotImplementedError('only (B, S) key-padding masks are supported')
        return (~attn_mask).sum(dim=1).to(torch.int32)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, attn_mask: Optional[Tensor] = None,
                state: Optional[dict] = None, return_attn: bool = False, key_lengths: Optional[Tensor] = None):
        """query (B, T, D); key/value (B, S, D), value must be the same tensor as key (as everywhere on the ST path).
        attn_mask: (B, S) bool key-padding mask with the padding at the end of each row (what Task.collate
        produces), or pass `key_lengths` (int32 (B,)) directly to skip the mask -> length conversion.
        Returns (attn (B, T, D), None)."""
        if state is not None or return_attn:
            raise NotImplementedError('incremental decoding / attention maps are outside the training path')
        if value is not key:
            raise NotImplementedError('key and value must be the same tensor')
        B, T, D = query.shape
        S = key.shape[1]
        if self.causal:
            attn_mask = key_lengths = None  # padding mask is dropped for causal attention (modules.py:602-605)
        elif key_lengths is None and attn_mask is not None:
            if attn_mask.dim() != 2:
                raise NotImplementedError('only (B, S) key-padding masks are supported')
            key_lengths = (~attn_mask).sum(dim=1).to(torch.int32)
        self_attn = key is query
        out = ops.AttentionFn.apply(
            query.reshape(B * T, D), None if self_attn else key.reshape(B * S, D), self.in_proj_weight, self.in_proj_bias,
            self.out_proj.weight, self.out_proj.bias, key_lengths, B, T, S, self.num_heads, self.causal,
            ops.attention_scale(self.head_dim, self.scaled), self.key_bias, float(self.dropout) if self.training else 0.0)
        return out.view(B, T, D), None


class ConvolutionSubsampler(nn.Module):
    """pasero/models/modules.py:774-834: Conv1d(k, stride 2, padding k//2) + GLU per layer, no masking of padded
    frames. Weights are stored tap-major [Cout, KW * Cin] (the implicit-GEMM B operand); state_dict exposes the
    reference's `conv_layers.{i}.weight` [Cout, Cin, KW] / `.bias`."""

    def __init__(self, in_channels: int, mid_channels: int, out_channels: int, kernel_sizes=(3, 3), strides=None,
                 activation: str = 'glu'):
        super().__init__()
        if activation not in ('glu', 'gelu'):
            raise NotImplementedError(f'conv activation {activation!r} is not covered')
        self.activation = activation
        r = 2 if activation == 'glu' else 1      # GLU halves the channels (modules.py:788)
        self.kernel_sizes = tuple(kernel_sizes)
        self.strides = tuple(strides) if strides else tuple(2 for _ in kernel_sizes)
        assert len(self.strides) == len(self.kernel_sizes)
        self.channels = []
        for i, k in enumerate(self.kernel_sizes):
            cin = in_channels if i == 0 else mid_channels // r
            cout = out_channels * r if i == len(self.kernel_sizes) - 1 else mid_channels
            w = torch.empty(cout, cin, k)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))            # nn.Conv1d default
            bound = 1 / math.sqrt(cin * k)
            self.register_parameter(f'weight{i}', nn.Parameter(w.permute(0, 2, 1).reshape(cout, k * cin).contiguous()))
            self.register_parameter(f'bias{i}', nn.Parameter(torch.empty(cout).uniform_(-bound, bound)))
            getattr(self, f'bias{i}')._pb200_f32_grad = True
            self.channels.append((cin, cout))

    @property
    def weights(self):
        return [getattr(self, f'weight{i}') for i in range(len(self.kernel_sizes))]

    @property
    def biases(self):
        return [getattr(self, f'bias{i}') for i in range(len(self.kernel_sizes))]

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for i, (k, (cin, cout)) in enumerate(zip(self.kernel_sizes, self.channels)):
            w = self.weights[i].detach().view(cout, k, cin).permute(0, 2, 1).contiguous()
            destination[f'{prefix}conv_layers.{i}.weight'] = w
            destination[f'{prefix}conv_layers.{i}.bias'] = self.biases[i].detach()
        # (the tap-major storage parameters weight{i}/bias{i} are deliberately not exposed)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for i, (k, (cin, cout)) in enumerate(zip(self.kernel_sizes, self.channels)):
            wk, bk = f'{prefix}conv_layers.{i}.weight', f'{prefix}conv_layers.{i}.bias'
            with torch.no_grad():
                if wk in state_dict:
                    self.weights[i].copy_(state_dict[wk].permute(0, 2, 1).reshape(cout, k * cin))
                elif strict:
                    missing_keys.append(wk)
                if bk in state_dict:
                    self.biases[i].copy_(state_dict[bk])
                elif strict:
                    missing_keys.append(bk)

    def get_new_length(self, length):
        for k, s in zip(self.kernel_sizes, self.strides):
            length = 1 + torch.div(length - k + 2 * (k // 2), s, rounding_mode='floor')
        return length

    @staticmethod
    def _out_len(t: int, k: int, s: int) -> int:
        return 1 + (t + 2 * (k // 2) - k) // s

    def padded_rows(self, t_in: int, layer: int = 0) -> tuple:
        """(rows, front_pad) of the zero-padded input layout layer `layer` reads for an input of t_in frames."""
        k, s = self.kernel_sizes[layer], self.strides[layer]
        t_out = self._out_len(t_in, k, s)
        return max(k // 2 + t_in, (t_out - 1) * s + k), k // 2

    def forward_padded(self, xp: Tensor, t_in: int, length: Tensor):
        """xp: [B, rows, C] already in the padded layout of layer 0 (e.g. written by the fbank kernel)."""
        B = xp.shape[0]
        x = xp
        n = len(self.kernel_sizes)
        t = t_in
        for i, (k, s) in enumerate(zip(self.kernel_sizes, self.strides)):
            t_out = self._out_len(t, k, s)
            if i + 1 < n:
                rows, front = self.padded_rows(t_out, i + 1)
            else:
                rows, front = t_out, 0
            x = ops.ConvGLUFn.apply(x, self.weights[i], self.biases[i], B, t, self.channels[i][0], k, s, rows, front,
                                    i > 0 or x.requires_grad, self.activation)
            t = t_out
        return x, self.get_new_length(length)

    def forward(self, x: Tensor, length: Tensor):
        """x (B, S, D) bf16 -> (B, S', D'), new lengths"""
        B, T, C = x.shape
        rows, front = self.padded_rows(T, 0)
        xp = ops.PadRowsFn.apply(x, rows, front) if x.requires_grad else K.pad_rows(x.contiguous(), B, T, C, rows, front)
        return self.forward_padded(xp, T, length)


class Embedding(nn.Embedding):
    """pasero/models/modules.py:890-947 (no partial freezing on this path). The table is stored with its row count
    rounded up to a multiple of 8 (TMA / 16-byte rows of the tied projection's logits; e.g. Whisper's 51 865 -> 51 872):
    the extra rows are zero, never receive a gradient (their logits are exactly 0 and the loss kernel only reads the real
    `vocab_size` columns) and are trimmed from / padded into the state_dict, which keeps the reference's shapes."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: int):
        self.vocab_size = num_embeddings
        super().__init__((num_embeddings + 7) // 8 * 8, embedding_dim, padding_idx)
        nn.init.normal_(self.weight, mean=0, std=embedding_dim ** -0.5)
        nn.init.constant_(self.weight[padding_idx], 0)
        nn.init.constant_(self.weight[num_embeddings:], 0)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        w = self.weight if keep_vars else self.weight.detach()
        destination[prefix + 'weight'] = w[:self.vocab_size]

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + 'weight'
        if key in state_dict:
            w = state_dict[key]
            if w.shape[0] not in (self.vocab_size, self.weight.shape[0]) or w.shape[1] != self.weight.shape[1]:
                error_msgs.append(f'size mismatch for {key}: {tuple(w.shape)} vs ({self.vocab_size}, {self.weight.shape[1]})')
                return
            with torch.no_grad():
                self.weight[:w.shape[0]].copy_(w)
                self.weight[w.shape[0]:].zero_()
        elif strict:
            missing_keys.append(key)

    def projection(self, input: Tensor) -> Tensor:
        return ops.linear(input, self.weight, None)[..., :self.vocab_size]
