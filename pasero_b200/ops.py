"""Autograd-aware fused ops of the ST training step. Every forward/backward here is a handful of C-ABI kernel
launches (pasero_b200.kernels); torch only provides device memory, the autograd graph and the stream.

Dropout masks are Philox(seed, offset): `DropoutState` hands out a fresh offset per call so that backward can
regenerate the same mask without storing it.
"""
import math
import os
import weakref
from typing import Optional

import torch
from torch import Tensor

from . import kernels as K

BF16 = torch.bfloat16


class DropoutState:
    """Process-wide Philox (seed, running offset). One offset per dropout site per forward."""
    seed = 0
    offset = 0

    @classmethod
    def reseed(cls, seed: int) -> None:
        cls.seed, cls.offset = int(seed) & (2 ** 63 - 1), 0

    @classmethod
    def next(cls) -> tuple:
        cls.offset += 1
        return cls.seed, cls.offset


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


class DirectGrads:
    """Trainer-installed sink for parameter gradients. When active, the backward kernels write each weight / bias
    gradient straight into its slice of the flat gradient arena (overwrite on the first write of an optimizer step,
    accumulate afterwards: tied embeddings, gradient-accumulation micro-batches) and the autograd Functions return None
    for parameters, so no AccumulateGrad `+=` kernels run and the arena needs no per-step zeroing. `on_ready(param)`
    is called once a parameter's gradient for this micro-batch is complete (drives the allreduce buckets)."""

    active: Optional['DirectGrads'] = None

    def __init__(self, views: dict, on_ready=None, views32: Optional[dict] = None, params=()):
        # keyed by id(param): `owners` holds weak references so that a recycled id (a parameter of a model that has been
        # garbage-collected while this sink is still installed) is never mistaken for one of ours
        self.owners = {id(p): weakref.ref(p) for p in params}
        self.views = views            # id(param) -> bf16 arena view
        # id(param) -> fp32 main-gradient view (GEMM weights): weight-gradient GEMMs TMA-reduce-add their (split-K)
        # partial tiles straight into it; the Trainer rounds it into the bf16 arena once per optimizer step
        self.views32 = views32 or {}
        self.written = set()
        self.on_ready = on_ready

    def start_step(self) -> None:
        self.written.clear()

    def owns(self, param) -> bool:
        ref = self.owners.get(id(param))
        return ref is not None and ref() is param

    def sink(self, param) -> tuple:
        """(out tensor, accumulate flag) for this parameter, or (None, False) when not managed."""
        view = self.views.get(id(param)) if self.owns(param) else None
        if view is None:
            return None, False
        acc = id(param) in self.written
        self.written.add(id(param))
        return view, acc

    def ready(self, *params) -> None:
        if self.on_ready is not None:
            for p in params:
                if p is not None and self.owns(p):
                    self.on_ready(p)


class WgradStream:
    """Second CUDA stream for the weight-gradient GEMMs that accumulate into the fp32 main-gradient arena. Nothing reads
    those sums before the gradient flush, so the GEMMs need not sit on backward's critical path: forked onto this stream
    they fill the SMs that the dependent chain (dgrad -> attention backward -> LayerNorm backward ...) leaves idle at
    kernel tails and in the decoder's small launches. Works the same eagerly and under CUDA-graph capture (the fork /
    join become graph edges). Operands are kept referenced until `join()` so the caching allocator cannot hand their
    memory to a later kernel of the main stream while the GEMM still reads it. PB200_WGRAD_STREAM=0 disables it."""
    enabled = os.environ.get('PB200_WGRAD_STREAM', '1') != '0'
    _streams: dict = {}
    _pending: list = []
    _forked = None      # the side stream with un-joined work, if any

    @classmethod
    def run(cls, fn, *operands) -> None:
        if not cls.enabled:
            fn()
            return
        main = torch.cuda.current_stream()
        side = cls._streams.get(main.device)
        if side is None:
            side = cls._streams[main.device] = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)          # operands were produced on the main stream
        with torch.cuda.stream(side):
            fn()
        cls._pending.extend(operands)
        cls._forked = side

    @classmethod
    def run_behind(cls, fn):
        """Runs fn() on the side stream BEHIND everything forked so far and everything already on the current stream, without
        making the current stream wait (the gradient-bucket flush: it needs the bucket's weight-gradient GEMMs, which are all
        on the side stream already, not the other way round — a join here would stall the dependent chain of backward until
        the side stream has drained, once per bucket). Returns the stream fn ran on: the side stream, or the current stream
        when the second stream is disabled."""
        if not cls.enabled:
            fn()
            return torch.cuda.current_stream()
        cls.run(fn)
        return cls._forked

    @classmethod
    def join(cls) -> None:
        """The current stream waits for every forked GEMM (call before anything reads the fp32 main gradients)."""
        if cls._forked is not None:
            torch.cuda.current_stream().wait_stream(cls._forked)
            cls._forked = None
            cls._pending.clear()


def _sink(param):
    d = DirectGrads.active
    return d.sink(param) if (d is not None and param is not None) else (None, False)


def _ready(*params):
    d = DirectGrads.active
    if d is not None:
        d.ready(*params)


def _sink32(param):
    """fp32 main-gradient view of a GEMM weight (None when the Trainer does not provide one)."""
    d = DirectGrads.active
    return d.views32.get(id(param)) if (d is not None and param is not None and d.owns(param)) else None


def _wgrad(dy: Tensor, x: Tensor, param):
    """dW[N, K] = dy^T x of a Linear, routed to the best sink: fp32 main gradient (always accumulate), bf16 arena view
    (overwrite / accumulate), or a fresh tensor handed back to autograd. Returns what autograd should see."""
    v32 = _sink32(param)
    if v32 is not None:
        WgradStream.run(lambda: K.linear_wgrad(dy, x, out=v32, accumulate=True), dy, x)
        return None
    out, acc = _sink(param)
    return _ret(K.linear_wgrad(dy, x, out=out, accumulate=acc), out)


def _wgrad_bias(dy: Tensor, x: Tensor, w_out: Tensor, acc_w: bool, b32: Optional[Tensor], b_out: Optional[Tensor],
                acc_b: bool) -> None:
    """w_out (+)= dy^T x and the bias gradient (column sums of dy) of the same Linear. A bias with an fp32 main gradient
    (b32) gets its sums from the weight-gradient GEMM itself when that one accumulates in fp32 too — dy is then read once
    instead of twice; otherwise a column-sum kernel writes b32 (+=) or the bf16 sink b_out."""
    fuse = b32 is not None and w_out.dtype == torch.float32 and acc_w
    if w_out.dtype == torch.float32 and acc_w:   # fp32 main gradient: off the critical path
        WgradStream.run(lambda: K.linear_wgrad(dy, x, out=w_out, accumulate=True, bias_grad32=b32 if fuse else None), dy, x)
    else:
        K.linear_wgrad(dy, x, out=w_out, accumulate=acc_w)
    if b32 is not None:
        if not fuse:
            K.colsum(dy, out=b32, accumulate=True)
    elif b_out is not None:
        K.colsum(dy, out=b_out, accumulate=acc_b)


def _linear_grads(dy: Tensor, x: Tensor, w_param, b_param, need_w: bool = True, need_b: bool = True):
    """(dW, db) of y = x W^T + b for autograd (None where the gradient went straight into a DirectGrads sink)."""
    dw = db = None
    need_b = need_b and b_param is not None
    b32 = _sink32(b_param) if need_b else None
    w32 = _sink32(w_param) if need_w else None
    if w32 is not None and b32 is not None:
        _wgrad_bias(dy, x, w32, True, b32, None, False)
        return None, None
    if need_w:
        dw = _wgrad(dy, x, w_param)
    if need_b:
        if b32 is not None:
            K.colsum(dy, out=b32, accumulate=True)
        else:
            o, a = _sink(b_param)
            db = _ret(K.colsum(dy, out=o, accumulate=a), o)
    return dw, db


def _ret(grad, out):
    """Gradient to hand back to autograd: None when it already went into the arena."""
    return None if out is not None else grad


# --------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x @ W^T + b   (x: [M, K] bf16, W: [N, K])"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        ctx.save_for_backward(x, weight)
        ctx.params = (weight, bias)
        return K.linear_fwd(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        w_param, b_param = ctx.params
        dy = _c(dy)
        dx = K.linear_dgrad(dy, weight) if ctx.needs_input_grad[0] else None
        dw, db = _linear_grads(dy, x, w_param, b_param, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        _ready(w_param, b_param)
        return dx, dw, db


def _fc1_relu(x: Tensor, w1: Tensor, b1: Optional[Tensor]):
    """h = relu(x W1^T + b1) and, when the shape allows, the 1-bit sign map of h that the fc2 dgrad epilogue uses as
    its ReLU-backward mask (reading h itself there costs 16x the epilogue traffic)."""
    bits = None
    if K.relu_bits_supported(w1.shape[0], b1):
        bits = torch.empty(w1.shape[0] // 32, x.shape[0], dtype=torch.int32, device=x.device)
    return K.linear_fwd(x, w1, b1, act=1, relu_bits_out=bits), bits


class LinearReluFn(torch.autograd.Function):
    """h = relu(x @ W^T + b) with the activation fused in the GEMM epilogue (the encoder's `in_linear`,
    pasero/models/transformer.py:647-651)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        h = K.linear_fwd(x, weight, bias, act=1)
        ctx.save_for_backward(x, weight, h)
        ctx.params = (weight, bias)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, weight, h = ctx.saved_tensors
        w_param, b_param = ctx.params
        dpre = K.relu_bwd(_c(dh), h)
        dx = K.linear_dgrad(dpre, weight) if ctx.needs_input_grad[0] else None
        dw, db = _linear_grads(dpre, x, w_param, b_param, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        _ready(w_param, b_param)
        return dx, dw, db


class FFNFn(torch.autograd.Function):
    """y = relu(x W1^T + b1) W2^T + b2 (pasero/models/transformer.py:999-1019); ReLU fused in the fc1 epilogue, its
    backward mask fused in the fc2 dgrad epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x = _c(x)
        h, bits = _fc1_relu(x, w1, b1)
        y = K.linear_fwd(h, w2, b2)
        ctx.save_for_backward(x, w1, w2, h, bits)
        ctx.params = (w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, h, bits = ctx.saved_tensors
        p_w1, p_b1, p_w2, p_b2 = ctx.params
        dy = _c(dy)
        dh = K.linear_dgrad(dy, w2, relu_mask=h if bits is None else None, relu_bits=bits)
        dw2, db2 = _linear_grads(dy, h, p_w2, p_b2)
        _ready(p_w2, p_b2)
        dx = K.linear_dgrad(dh, w1)
        dw1, db1 = _linear_grads(dh, x, p_w1, p_b1)
        _ready(p_w1, p_b1)
        return dx, dw1, db1, dw2, db2


class AttentionFn(torch.autograd.Function):
    """Multi-head attention block with fused projections (pasero/models/modules.py:579-739, training path).

    in_w: [3D, D] rows ordered (k, v, q) — the reference's parameter creation order — in_b: [3D];
    self-attention: one GEMM produces k|v|q; cross-attention: q from `query`, k|v from `key` in one GEMM.
    """

    @staticmethod
    def forward(ctx, query, key, in_w, in_b, out_w, out_b, kv_len, B, Lq, Lk, H, causal, scale, key_bias=True, attn_p=0.0):
        D = H * 64
        ctx.key_bias = key_bias
        query = _c(query)
        self_attn = key is None
        if self_attn:
            kvq = K.linear_fwd(query, in_w, in_b)                       # [B*L, 3D]
            k2, v2, q2 = kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:]
            key_c = None
        else:
            key_c = _c(key)
            kv = K.linear_fwd(key_c, in_w[:2 * D], in_b[:2 * D])         # [B*Lk, 2D]
            q2 = K.linear_fwd(query, in_w[2 * D:], in_b[2 * D:])         # [B*Lq, D]
            k2, v2 = kv[:, :D], kv[:, D:]
            kvq = None
        drop = _attn_dropout(attn_p, B, H, Lq, Lk, query.device)
        o, lse = K.attention_fwd(q2, k2, v2, B, H, Lq, Lk, kv_len, causal, scale, dropout=drop)
        y = K.linear_fwd(o, out_w, out_b)
        ctx.save_for_backward(query, key_c, in_w, out_w, q2, k2, v2, o, lse, kv_len)
        ctx.drop = drop
        ctx.dims = (B, Lq, Lk, H, causal, scale, self_attn)
        ctx.params = (in_w, in_b, out_w, out_b)
        return y

    @staticmethod
    def backward(ctx, dy):
        query, key_c, in_w, out_w, q2, k2, v2, o, lse, kv_len = ctx.saved_tensors
        B, Lq, Lk, H, causal, scale, self_attn = ctx.dims
        D = H * 64
        p_in_w, p_in_b, p_out_w, p_out_b = ctx.params
        dy = _c(dy)
        d_o = K.linear_dgrad(dy, out_w)
        d_out_w, d_out_b = _linear_grads(dy, o, p_out_w, p_out_b)
        _ready(p_out_w, p_out_b)
        sink_w, acc_w = _sink(p_in_w)
        sink_b, acc_b = _sink(p_in_b)
        w32 = _sink32(p_in_w)
        if w32 is not None:      # fp32 main gradient: always accumulate, nothing to hand back
            sink_w, acc_w = w32, True
        d_in_w = sink_w if sink_w is not None else torch.empty_like(in_w)
        d_in_b = sink_b if sink_b is not None else torch.zeros(3 * D, dtype=dy.dtype, device=dy.device)
        b32 = _sink32(p_in_b)    # fp32 main gradient of the fused bias: filled by the weight-gradient GEMMs themselves
        kb = ctx.key_bias   # False: the k rows of the fused bias are not a parameter of the reference model -> no gradient
        if self_attn:
            dkvq = torch.empty(B * Lq, 3 * D, dtype=dy.dtype, device=dy.device)
            K.attention_bwd(d_o, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale,
                            dkvq[:, 2 * D:], dkvq[:, :D], dkvq[:, D:2 * D], dropout=ctx.drop)
            dquery = K.linear_dgrad(dkvq, in_w)
            if kb:
                _wgrad_bias(dkvq, query, d_in_w, acc_w, b32, d_in_b, acc_b)
            else:
                K.linear_wgrad(dkvq, query, out=d_in_w, accumulate=acc_w)
                K.colsum(dkvq[:, D:], out=(b32 if b32 is not None else d_in_b)[D:], accumulate=acc_b or b32 is not None)
            dkey = None
        else:
            dq = torch.empty(B * Lq, D, dtype=dy.dtype, device=dy.device)
            dkv = torch.empty(B * Lk, 2 * D, dtype=dy.dtype, device=dy.device)
            K.attention_bwd(d_o, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale, dq, dkv[:, :D], dkv[:, D:], dropout=ctx.drop)
            dquery = K.linear_dgrad(dq, in_w[2 * D:])
            dkey = K.linear_dgrad(dkv, in_w[:2 * D])
            if kb:
                _wgrad_bias(dkv, key_c, d_in_w[:2 * D], acc_w, b32[:2 * D] if b32 is not None else None, d_in_b[:2 * D], acc_b)
            else:
                K.linear_wgrad(dkv, key_c, out=d_in_w[:2 * D], accumulate=acc_w)
                K.colsum(dkv[:, D:], out=(b32 if b32 is not None else d_in_b)[D:2 * D], accumulate=acc_b or b32 is not None)
            _wgrad_bias(dq, query, d_in_w[2 * D:], acc_w, b32[2 * D:] if b32 is not None else None, d_in_b[2 * D:], acc_b)
        _ready(p_in_w, p_in_b)
        return (dquery, dkey, _ret(d_in_w, sink_w), _ret(d_in_b, sink_b), d_out_w, d_out_b,
                None, None, None, None, None, None, None, None, None)


def _attn_dropout(p: float, B: int, H: int, Lq: int, Lk: int, device):
    """(p, seed, offset, keep-bit buffer) of one attention call (pasero/models/modules.py:686-731), or None when p == 0."""
    if not p or p <= 0:
        return None
    seed, offset = DropoutState.next()
    return float(p), seed, offset, K.attention_mask_buffer(B, H, Lq, Lk, device)


def _ln_bwd(dy, s, gamma, mean, rstd, p, seed, offset, params, lin_bias, keep=None, need_ln=True, need_lin=True):
    """Shared backward of `LN(residual + dropout(x))`: returns (dx, dresidual, dgamma, dbeta, dlin_bias) with the
    parameter gradients routed through DirectGrads when active (None is returned for those). need_ln / need_lin = False
    (frozen parameters) skip the corresponding column sums."""
    p_gamma, p_beta = params
    D = dy.shape[1]
    if not need_ln or not need_lin:
        dgamma = dbeta = dlin = None
        sg = sb = sl = None
        ag = ab = al = False
        if need_ln:
            sg, ag = _sink(p_gamma); sb, ab = _sink(p_beta)
            dgamma = sg if sg is not None else torch.empty(D, dtype=dy.dtype, device=dy.device)
            dbeta = sb if sb is not None else (torch.empty(D, dtype=dy.dtype, device=dy.device) if p_beta is not None else None)
        if need_lin and lin_bias is not None:
            sl, al = _sink(lin_bias)
            dlin = sl if sl is not None else torch.empty(D, dtype=dy.dtype, device=dy.device)
        # every output is an arena sink (nothing is handed back to autograd): the partials' reduction can leave the chain
        side = WgradStream.run if ((dgamma is None or sg is not None) and (dbeta is None or sb is not None)
                                   and (dlin is None or sl is not None)) else None
        dx, dres = K.dropout_add_ln_bwd(dy, None, s, gamma, mean, rstd, p, seed, offset, True, dgamma, dbeta,
                                        accumulate=ag or ab, dbias_lin=dlin, accumulate_lin=al, keep_mask=keep, side_run=side)
        return dx, dres, _ret(dgamma, sg), _ret(dbeta, sb), _ret(dlin, sl)
    sg, ag = _sink(p_gamma)
    sb, ab = _sink(p_beta)
    sl, al = _sink(lin_bias)
    dgamma = sg if sg is not None else torch.empty(D, dtype=dy.dtype, device=dy.device)
    dbeta = sb if sb is not None else (torch.empty(D, dtype=dy.dtype, device=dy.device) if p_beta is not None else None)
    dlin = sl if sl is not None else (torch.empty(D, dtype=dy.dtype, device=dy.device) if lin_bias is not None else None)
    side = WgradStream.run if (sg is not None and (dbeta is None or sb is not None) and (dlin is None or sl is not None)) else None
    dx, dres = K.dropout_add_ln_bwd(dy, None, s, gamma, mean, rstd, p, seed, offset, True, dgamma, dbeta,
                                    accumulate=ag or ab, dbias_lin=dlin, accumulate_lin=al, keep_mask=keep, side_run=side)
    return dx, dres, _ret(dgamma, sg), _ret(dbeta, sb), _ret(dlin, sl)


FUSE_LINEAR_LN = True    # tests / tools switch it off to compare the fused kernel with the two-kernel path


def _linear_dropout_add_ln(a: Tensor, w: Tensor, b: Optional[Tensor], residual: Tensor, gamma, beta, eps: float, p: float,
                           seed: int, offset: int):
    """(out, sum, mean, rstd, keep bits) of LN(residual + dropout(a w^T + b)): one fused kernel when the shape allows
    (d_model 512, encoder-sized row counts: csrc/gemm_ln.cu), else the GEMM followed by the dropout + residual + LN kernel."""
    M, N = a.shape[0], w.shape[0]
    keep = torch.empty(M * N // 8, dtype=torch.uint8, device=a.device) if p > 0 else None
    if FUSE_LINEAR_LN and K.linear_ln_supported(M, N, a.shape[1], residual, gamma):
        out, s, mean, rstd = K.linear_dropout_add_ln_fwd(a, w, b, residual, gamma, beta, eps, p, seed, offset, keep_out=keep)
    else:
        y = K.linear_fwd(a, w, b)
        out, s, mean, rstd = K.dropout_add_ln_fwd(y, residual, gamma, beta, eps, p, seed, offset, keep_out=keep)
    return out, s, mean, rstd, keep


class GradCell:
    """Running sum of the gradients that several autograd nodes owe ONE tensor — the encoder output, read by the
    cross-attention of every decoder layer. Autograd would add the nodes' contributions with one elementwise kernel per extra
    consumer (five [B*S, D] adds per step for six decoder layers); instead each node's key / value dgrad GEMM takes the running
    sum as its residual operand (added in the GEMM epilogue), the nodes return None for that input, and the last one to run
    hands autograd the total. Counting down `remaining` makes the result independent of the order the engine picks."""

    def __init__(self, consumers: int):
        self.remaining, self.acc = consumers, None

    def contribute(self, fn):
        """fn(residual or None) -> this node's gradient plus the residual. Returns the total for the last node, else None."""
        self.acc = fn(self.acc)
        self.remaining -= 1
        if self.remaining > 0:
            return None
        total, self.acc = self.acc, None
        return total


def _dkey(dkv: Tensor, w_kv: Tensor, cell: Optional[GradCell]):
    if cell is None:
        return K.linear_dgrad(dkv, w_kv)
    return cell.contribute(lambda acc: K.linear_dgrad(dkv, w_kv, residual=acc))


class AttnSublayerFn(torch.autograd.Function):
    """Post-norm attention sub-block in one autograd node: LN(x + dropout(MHA(x, key))) (pasero/models/transformer.py
    :1021-1054, 1246-1289). Compared with AttentionFn followed by DropoutAddLNFn, the backward gets the out-projection
    bias gradient from the LayerNorm-backward pass and adds the skip-connection gradient in the epilogue of the last
    dgrad GEMM (no column-sum pass over dy, no separate add kernel)."""

    @staticmethod
    def forward(ctx, x, key, in_w, in_b, out_w, out_b, gamma, beta, kv_len, B, Lq, Lk, H, causal, scale, p, eps, key_bias=True, attn_p=0.0,
                key_grad_cell=None):
        D = H * 64
        ctx.key_bias = key_bias
        ctx.key_grad_cell = key_grad_cell
        x = _c(x)
        self_attn = key is None
        if self_attn:
            kvq = K.linear_fwd(x, in_w, in_b)
            k2, v2, q2 = kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:]
            key_c = None
        else:
            key_c = _c(key)
            kv = K.linear_fwd(key_c, in_w[:2 * D], in_b[:2 * D])
            q2 = K.linear_fwd(x, in_w[2 * D:], in_b[2 * D:])
            k2, v2 = kv[:, :D], kv[:, D:]
        drop = _attn_dropout(attn_p, B, H, Lq, Lk, x.device)
        ctx.drop = drop
        o, lse = K.attention_fwd(q2, k2, v2, B, H, Lq, Lk, kv_len, causal, scale, dropout=drop)
        seed, offset = DropoutState.next() if p > 0 else (0, 0)
        out, s, mean, rstd, keep = _linear_dropout_add_ln(o, out_w, out_b, x, gamma, beta, eps, p, seed, offset)
        ctx.save_for_backward(x, key_c, in_w, out_w, q2, k2, v2, o, lse, kv_len, s, gamma, mean, rstd, keep)
        ctx.dims = (B, Lq, Lk, H, causal, scale, self_attn, p, seed, offset)
        ctx.params = (in_w, in_b, out_w, out_b, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, key_c, in_w, out_w, q2, k2, v2, o, lse, kv_len, s, gamma, mean, rstd, keep = ctx.saved_tensors
        B, Lq, Lk, H, causal, scale, self_attn, p, seed, offset = ctx.dims
        D = H * 64
        p_in_w, p_in_b, p_out_w, p_out_b, p_gamma, p_beta = ctx.params
        need = ctx.needs_input_grad   # (x, key, in_w, in_b, out_w, out_b, gamma, beta, ...): frozen parameters cost nothing
        dy, dres, dgamma, dbeta, d_out_b = _ln_bwd(_c(dout), s, gamma, mean, rstd, p, seed, offset, (p_gamma, p_beta), p_out_b, keep,
                                                   need_ln=need[6], need_lin=need[5])
        _ready(p_gamma, p_beta)
        d_o = K.linear_dgrad(dy, out_w)
        d_out_w = _wgrad(dy, o, p_out_w) if need[4] else None
        _ready(p_out_w, p_out_b)
        if not need[2]:   # frozen projections: only the input gradients
            if self_attn:
                dkvq = torch.empty(B * Lq, 3 * D, dtype=dy.dtype, device=dy.device)
                K.attention_bwd(d_o, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale,
                                dkvq[:, 2 * D:], dkvq[:, :D], dkvq[:, D:2 * D], dropout=ctx.drop)
                return (K.linear_dgrad(dkvq, in_w, residual=dres), None, None, None, d_out_w, d_out_b, dgamma, dbeta) + (None,) * 12
            dq = torch.empty(B * Lq, D, dtype=dy.dtype, device=dy.device)
            dkv = torch.empty(B * Lk, 2 * D, dtype=dy.dtype, device=dy.device)
            K.attention_bwd(d_o, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale, dq, dkv[:, :D], dkv[:, D:], dropout=ctx.drop)
            dkey = _dkey(dkv, in_w[:2 * D], ctx.key_grad_cell) if need[1] else None
            return (K.linear_dgrad(dq, in_w[2 * D:], residual=dres), dkey, None, None, d_out_w, d_out_b, dgamma, dbeta) + (None,) * 12
        sink_w, acc_w = _sink(p_in_w)
        sink_b, acc_b = _sink(p_in_b)
        w32 = _sink32(p_in_w)
        if w32 is not None:      # fp32 main gradient: always accumulate, nothing to hand back
            sink_w, acc_w = w32, True
        d_in_w = sink_w if sink_w is not None else torch.empty_like(in_w)
        d_in_b = sink_b if sink_b is not None else torch.zeros(3 * D, dtype=dy.dtype, device=dy.device)
        b32 = _sink32(p_in_b)    # fp32 main gradient of the fused bias: filled by the weight-gradient GEMMs themselves
        kb = ctx.key_bias   # False: the k rows of the fused bias are not a parameter of the reference model -> no gradient
        if self_attn:
            dkvq = torch.empty(B * Lq, 3 * D, dtype=dy.dtype, device=dy.device)
            K.attention_bwd(d_o, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale,
                            dkvq[:, 2 * D:], dkvq[:, :D], dkvq[:, D:2 * D], dropout=ctx.drop)
            dx = K.linear_dgrad(dkvq, in_w, residual=dres)
            if kb:
                _wgrad_bias(dkvq, x, d_in_w, acc_w, b32, d_in_b, acc_b)
            else:
                K.linear_wgrad(dkvq, x, out=d_in_w, accumulate=acc_w)
                K.colsum(dkvq[:, D:], out=(b32 if b32 is not None else d_in_b)[D:], accumulate=acc_b or b32 is not None)
            dkey = None
        else:
            dq = torch.empty(B * Lq, D, dtype=dy.dtype, device=dy.device)
            dkv = torch.empty(B * Lk, 2 * D, dtype=dy.dtype, device=dy.device)
            K.attention_bwd(d_o, q2, k2, v2, o, lse, B, H, Lq, Lk, kv_len, causal, scale, dq, dkv[:, :D], dkv[:, D:], dropout=ctx.drop)
            dx = K.linear_dgrad(dq, in_w[2 * D:], residual=dres)
            dkey = _dkey(dkv, in_w[:2 * D], ctx.key_grad_cell) if need[1] else None
            if kb:
                _wgrad_bias(dkv, key_c, d_in_w[:2 * D], acc_w, b32[:2 * D] if b32 is not None else None, d_in_b[:2 * D], acc_b)
            else:
                K.linear_wgrad(dkv, key_c, out=d_in_w[:2 * D], accumulate=acc_w)
                K.colsum(dkv[:, D:], out=(b32 if b32 is not None else d_in_b)[D:2 * D], accumulate=acc_b or b32 is not None)
            _wgrad_bias(dq, x, d_in_w[2 * D:], acc_w, b32[2 * D:] if b32 is not None else None, d_in_b[2 * D:], acc_b)
        _ready(p_in_w, p_in_b)
        return (dx, dkey, _ret(d_in_w, sink_w), _ret(d_in_b, sink_b), d_out_w, d_out_b, dgamma, dbeta) + (None,) * 12


class FFNSublayerFn(torch.autograd.Function):
    """Post-norm feed-forward sub-block in one autograd node: LN(x + dropout(fc2(relu(fc1(x))))) (pasero/models/
    transformer.py:999-1019, 1073-1086); same backward fusions as AttnSublayerFn."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, p, eps):
        x = _c(x)
        h, bits = _fc1_relu(x, w1, b1)
        seed, offset = DropoutState.next() if p > 0 else (0, 0)
        out, s, mean, rstd, keep = _linear_dropout_add_ln(h, w2, b2, x, gamma, beta, eps, p, seed, offset)
        ctx.save_for_backward(x, w1, w2, h, s, gamma, mean, rstd, bits, keep)
        ctx.cfg = (p, seed, offset)
        ctx.params = (w1, b1, w2, b2, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w1, w2, h, s, gamma, mean, rstd, bits, keep = ctx.saved_tensors
        p, seed, offset = ctx.cfg
        p_w1, p_b1, p_w2, p_b2, p_gamma, p_beta = ctx.params
        need = ctx.needs_input_grad    # frozen parameters (adapter training): no weight-gradient GEMMs / column sums
        dy, dres, dgamma, dbeta, db2 = _ln_bwd(_c(dout), s, gamma, mean, rstd, p, seed, offset, (p_gamma, p_beta), p_b2, keep,
                                               need_ln=need[5], need_lin=need[4])
        _ready(p_gamma, p_beta)
        dh = K.linear_dgrad(dy, w2, relu_mask=h if bits is None else None, relu_bits=bits)
        dw2 = _wgrad(dy, h, p_w2) if need[3] else None
        _ready(p_w2, p_b2)
        dx = K.linear_dgrad(dh, w1, residual=dres)
        dw1, db1 = _linear_grads(dh, x, p_w1, p_b1, need[1], need[2])
        _ready(p_w1, p_b1)
        return dx, dw1, db1, dw2, db2, dgamma, dbeta, None, None


class DropoutAddLNFn(torch.autograd.Function):
    """y = LayerNorm(residual + dropout(x)) (gamma given) or residual + dropout(x) (gamma None)."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, p, eps):
        x = _c(x)
        residual = _c(residual) if residual is not None else None
        seed, offset = DropoutState.next() if p > 0 else (0, 0)
        # with LayerNorm the keep bits are stored (one byte per 8 elements): backward then reads them instead of redoing the
        # Philox rounds, which is also what selects csrc/norm.cu's training-shape backward kernel
        keep = (torch.empty(x.numel() // 8, dtype=torch.uint8, device=x.device)
                if p > 0 and gamma is not None and x.shape[1] % 8 == 0 else None)
        y, s, mean, rstd = K.dropout_add_ln_fwd(x, residual, gamma, beta, eps, p, seed, offset, keep_out=keep)
        ctx.save_for_backward(s if gamma is not None else None, gamma, mean, rstd, keep)
        ctx.cfg = (p, seed, offset, residual is not None, beta is not None)
        ctx.params = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        s, gamma, mean, rstd, keep = ctx.saved_tensors
        p, seed, offset, has_res, has_beta = ctx.cfg
        dy = _c(dy)
        D = dy.shape[1]
        p_gamma, p_beta = ctx.params
        sg, ag = _sink(p_gamma)
        sb, ab = _sink(p_beta)
        dgamma = sg if sg is not None else (torch.empty(D, dtype=dy.dtype, device=dy.device) if gamma is not None else None)
        dbeta = sb if sb is not None else (torch.empty(D, dtype=dy.dtype, device=dy.device) if (gamma is not None and has_beta) else None)
        dx, dres = K.dropout_add_ln_bwd(dy, None, s, gamma, mean, rstd, p, seed, offset, has_res, dgamma, dbeta,
                                        accumulate=ag or ab, keep_mask=keep)
        _ready(p_gamma, p_beta)
        return dx, dres, _ret(dgamma, sg), _ret(dbeta, sb), None, None


class EmbedFn(torch.autograd.Function):
    """dropout(bf16(bf16(src * scale) + pos)): src = table[ids] (decoder tokens) or a feature matrix (encoder).
    `pos_table` (optional): the trainable LearnedPositionalEmbedding weight whose rows [shift, shift + L) are `pos`; its
    gradient is the sum over the batch of the gradient reaching the sum (pasero/models/modules.py:460-484)."""

    @staticmethod
    def forward(ctx, src, ids, pos, scale, p, L, padding_idx, pos_table=None, shift=0, notify_ready=True):
        # notify_ready=False: the table is used again later in backward order (an encoder sharing it embeds tokens too);
        # only its last user reports the parameter's gradient complete (data-parallel bucket bookkeeping)
        ctx.notify_ready = notify_ready
        seed, offset = DropoutState.next() if p > 0 else (0, 0)
        M = ids.numel() if ids is not None else src.shape[0]
        D = src.shape[1]
        ids_flat = _c(ids).view(-1) if ids is not None else None
        out = K.embed_fwd(ids_flat, _c(src), pos.detach() if pos is not None else None, M, D, L, scale, p, seed, offset)
        ctx.save_for_backward(ids_flat)
        ctx.cfg = (M, D, scale, p, seed, offset, padding_idx, src.shape[0], L, shift)
        ctx.table = src if ids is not None else None
        ctx.pos_table = pos_table
        return out

    @staticmethod
    def backward(ctx, dy):
        (ids_flat,) = ctx.saved_tensors
        M, D, scale, p, seed, offset, padding_idx, rows, L, shift = ctx.cfg
        dy = _c(dy)
        dpos = None
        dense = None   # gradient reaching bf16(src * scale) + pos, i.e. dropout'(dy)
        if ctx.pos_table is not None:
            dense = K.embed_bwd(dy, None, M, D, 1.0, p, seed, offset)
            so, acc = _sink(ctx.pos_table)
            full = so if so is not None else torch.zeros_like(ctx.pos_table)
            if so is not None and not acc:
                full.zero_()       # rows outside [shift, shift + L) get no gradient this step
            K.colsum(dense.view(M // L, L * D), out=full[shift:shift + L].view(-1), accumulate=acc)
            _ready(ctx.pos_table)
            dpos = _ret(full, so)
        if ids_flat is None:
            dx = dense if (dense is not None and scale == 1.0) else K.embed_bwd(dy, None, M, D, scale, p, seed, offset)
            return dx, None, None, None, None, None, None, dpos, None, None
        t32 = _sink32(ctx.table)
        if t32 is not None:       # scatter-add straight into the fp32 main gradient of the (tied) embedding table
            K.embed_bwd(dy, ids_flat, M, D, scale, p, seed, offset, padding_idx, t32)
            if ctx.notify_ready:
                _ready(ctx.table)
            return None, None, None, None, None, None, None, dpos, None, None
        dtable32 = torch.zeros(rows, D, dtype=torch.float32, device=dy.device)
        K.embed_bwd(dy, ids_flat, M, D, scale, p, seed, offset, padding_idx, dtable32)
        so, acc = _sink(ctx.table)
        dtable = so if so is not None else torch.empty(rows, D, dtype=dy.dtype, device=dy.device)
        K.f32_to_bf16(dtable32, dtable, accumulate=acc)
        if ctx.notify_ready:
            _ready(ctx.table)
        return _ret(dtable, so), None, None, None, None, None, None, dpos, None, None


class FFNGeluFn(torch.autograd.Function):
    """y = gelu(x W1^T + b1) W2^T + b2 (erf GELU; pasero/models/transformer.py:999-1019 with activation_fn = gelu, the
    Whisper family). The pre-activation is kept for the backward pass, the activation is its own elementwise kernel."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x = _c(x)
        pre = K.linear_fwd(x, w1, b1)
        M, F_ = pre.shape
        h = K.gelu_fwd(pre, 1, M, F_, M, 0).view(M, F_)
        y = K.linear_fwd(h, w2, b2)
        ctx.save_for_backward(x, w1, w2, pre, h)
        ctx.params = (w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, pre, h = ctx.saved_tensors
        p_w1, p_b1, p_w2, p_b2 = ctx.params
        dy = _c(dy)
        dh = K.linear_dgrad(dy, w2)
        dw2, db2 = _linear_grads(dy, h, p_w2, p_b2)
        _ready(p_w2, p_b2)
        dpre = K.gelu_bwd(dh, pre)
        dx = K.linear_dgrad(dpre, w1)
        dw1, db1 = _linear_grads(dpre, x, p_w1, p_b1)
        _ready(p_w1, p_b1)
        return dx, dw1, db1, dw2, db2


class PadRowsFn(torch.autograd.Function):
    """[B, T, C] -> zero-padded [B, Tp, C] with the data at rows [front, front + T) (the conv subsampler's input layout);
    differentiable so that an `in_linear` in front of the subsampler receives its gradient."""

    @staticmethod
    def forward(ctx, x, Tp, front):
        B, T, C = x.shape
        ctx.dims = (T, front)
        return K.pad_rows(_c(x), B, T, C, Tp, front)

    @staticmethod
    def backward(ctx, dout):
        T, front = ctx.dims
        return _c(dout[:, front:front + T]), None, None


class ConvGLUFn(torch.autograd.Function):
    """One layer of the conv subsampler: Conv1d(kernel KW, stride, padding KW//2) + GLU over channels
    (pasero/models/modules.py:829-834) as a TMA-described implicit GEMM over the zero-padded [B, Tp, C] input
    (row t of batch b = the KW*C contiguous elements starting at padded frame t*stride) with a fused bias epilogue,
    then GLU written into the next layer's padded layout.

    weight: [Cout, KW * Cin] in tap-major order (w[o, tap * Cin + c] = conv.weight[o, c, tap]).
    """

    @staticmethod
    def forward(ctx, xp, weight, bias, B, Tin, Cin, KW, stride, out_Tp, out_front_pad, need_input_grad, act='glu'):
        Tp = xp.shape[1]
        Cout = weight.shape[0]
        Tout = 1 + (Tin + 2 * (KW // 2) - KW) // stride
        pre = torch.empty(B * Tout, Cout, dtype=xp.dtype, device=xp.device)
        K.gemm_raw(xp.data_ptr(), weight.data_ptr(), pre, Tout, Cout, KW * Cin, lda=stride * Cin, ldb=KW * Cin, ldc=Cout,
                   batch=B, a_batch_stride=Tp * Cin, b_batch_stride=0, c_batch_stride=Tout * Cout, bias=bias, split_k=1,
                   elem_dtype=xp.dtype)
        if act == 'glu':
            out = K.glu_fwd(pre, B, Tout, Cout // 2, out_Tp, out_front_pad)
        else:   # GELU keeps all Cout channels (Whisper-style front end, modules.py:803)
            out = K.gelu_fwd(pre, B, Tout, Cout, out_Tp, out_front_pad)
        ctx.act = act
        ctx.save_for_backward(xp, weight, pre)
        ctx.dims = (B, Tin, Cin, KW, stride, Tout, Tp, out_Tp, out_front_pad, need_input_grad)
        ctx.params = (weight, bias)
        return out

    @staticmethod
    def backward(ctx, dout):
        xp, weight, pre = ctx.saved_tensors
        B, Tin, Cin, KW, stride, Tout, Tp, out_Tp, out_front_pad, need_input_grad = ctx.dims
        Cout = weight.shape[0]
        # gradient w.r.t. the activation's output rows (drop the padding rows of the padded layout)
        if ctx.act == 'glu':
            dy = _c(dout[:, out_front_pad:out_front_pad + Tout]).view(B * Tout, Cout // 2)
            dpre = K.glu_bwd(dy, pre, B * Tout, Cout // 2)
        else:
            dy = _c(dout[:, out_front_pad:out_front_pad + Tout]).view(B * Tout, Cout)
            dpre = K.gelu_bwd(dy, pre)
        p_w, p_b = ctx.params
        so_w, acc_w = _sink(p_w)
        w32 = _sink32(p_w)
        if w32 is not None:
            so_w, acc_w = w32, True
        dw = so_w if so_w is not None else torch.empty_like(weight)
        b32 = _sink32(p_b)
        fuse_b = b32 is not None and w32 is not None   # bias gradient = row sums of the GEMM's A operand (see _wgrad_bias)
        K.gemm_raw(dpre.data_ptr(), xp.data_ptr(), dw, Cout, KW * Cin, Tout, lda=Cout, ldb=stride * Cin, ldc=KW * Cin, batch=B,
                   a_batch_stride=Tout * Cout, b_batch_stride=Tp * Cin, a_mn_major=True, b_mn_major=True, reduce_batch=True,
                   accumulate=acc_w, a_rowsum=b32 if fuse_b else None, elem_dtype=dpre.dtype)
        so_b, acc_b = _sink(p_b)
        if b32 is not None:
            db = b32 if fuse_b else K.colsum(dpre, out=b32, accumulate=True)
        else:
            db = K.colsum(dpre, out=so_b, accumulate=acc_b)
        _ready(p_w, p_b)
        dw, db = _ret(dw, so_w), _ret(db, so_b)
        dxp = None
        if need_input_grad:
            dcol = torch.empty(B * Tout, KW * Cin, dtype=dout.dtype, device=dout.device)
            K.gemm_raw(dpre.data_ptr(), weight.data_ptr(), dcol, B * Tout, KW * Cin, Cout, lda=Cout, ldb=KW * Cin,
                       ldc=KW * Cin, b_mn_major=True, split_k=1, elem_dtype=dpre.dtype)
            dx = K.col2im(dcol, B, Tin, Tout, Cin, KW, stride)            # [B*Tin, Cin]
            front = KW // 2
            dxp = torch.zeros_like(xp)
            dxp[:, front:front + Tin] = dx.view(B, Tin, Cin)
        return dxp, dw, db, None, None, None, None, None, None, None, None, None


class ProjCEFn(torch.autograd.Function):
    """Tied output projection + label-smoothed cross-entropy (pasero/models/modules.py:935-947,
    transformer.py:323-380). Returns the summed loss (fp32 0-dim); `stats` (fp64[8], device) receives loss / nll / token
    count.

    V < kernels.CE_CHUNK_MIN_VOCAB: the logits are produced once in bf16 and overwritten in place by their gradient in the
    same pass that computes the loss; no fp32 copy, no softmax tensor.
    Larger vocabularies (BASELINE config 5: 70 376): the projection is evaluated in chunks of kernels.CE_CHUNK_COLS columns
    that never leave L2 — pass 1 folds each chunk into per-row online log-sum-exp statistics, pass 2 recomputes the chunk,
    turns it into d loss / d logits in place and feeds it straight to the chunk's dgrad / wgrad GEMMs. Neither the logits
    nor their gradient exist as a [B*T, V] tensor (the reference: three fp32 copies, transformer.py:357, 367); the price is
    the projection GEMM a second time (+1/3 of the projection FLOPs)."""

    unit_loss_grad = False   # Trainer sets this when the loss is the root of backward (grad_output == 1)
    loss_scale = 1.0         # float16 training: the Trainer's current loss scale, folded into d loss / d logits where it is written

    @staticmethod
    def forward(ctx, x, table, target, padding_idx, label_smoothing, stats, vocab=None):
        # vocab: number of real rows of `table` (the rest is zero padding up to a multiple of 8: their logits are exactly
        # 0, the loss kernel neither reads nor overwrites them, so they act as zero gradients in backward)
        x = _c(x)
        V = vocab if vocab is not None else table.shape[0]
        ctx.table = table
        before = stats[0].clone()
        ctx.elem_dtype = x.dtype
        if V >= K.CE_CHUNK_MIN_VOCAB:
            ctx.chunked = True
            need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]   # (grad mode is off inside forward: ask autograd)
            dx32, dtable = ProjCEFn._chunked(x, table, target, padding_idx, label_smoothing, stats, V, need_grad)
            ctx.save_for_backward(dx32, dtable)
            return (stats[0] - before).float()
        ctx.chunked = False
        logits = K.linear_fwd(x, table, None)
        K.ls_ce_fwd_bwd(logits, target, padding_idx, label_smoothing, stats, write_grad=True, vocab=vocab,
                        grad_scale=ProjCEFn.loss_scale)
        ctx.save_for_backward(x, table, logits)
        return (stats[0] - before).float()

    @staticmethod
    def _chunked(x, table, target, padding_idx, eps, stats, V, need_grad):
        """Both passes. Returns (dx fp32 [M, D] or None, dtable bf16 [rows, D] or None when it went into the Trainer's
        fp32 main-gradient arena). Gradients are those of the summed loss (grad_output == 1) times ProjCEFn.loss_scale (float16
        training: the scale enters where d loss / d logits is written, like in the one-pass kernel); backward rescales otherwise."""
        M, D = x.shape
        cols = K.CE_CHUNK_COLS
        buf = torch.empty(M, cols, dtype=x.dtype, device=x.device)
        state = torch.empty(4, M, dtype=torch.float32, device=x.device)
        chunks = [(c0, min(cols, V - c0)) for c0 in range(0, V, cols)]

        def project(c0, n):
            # rows [c0, c0 + n8) of the table, n8 = n rounded up to 8 (inside the zero-padded table): logits of this chunk
            n8 = (n + 7) // 8 * 8
            return K.gemm_raw(x.data_ptr(), table[c0:c0 + n8].data_ptr(), buf, M, n8, D, lda=D, ldb=D, ldc=cols, split_k=1,
                              elem_dtype=x.dtype), n8
        for c0, n in chunks:
            project(c0, n)
            K.ce_chunk_stats(buf, target, c0, n, padding_idx, state)
        lse = K.ce_chunk_finish(state, target, V, padding_idx, eps, stats)
        if not need_grad:
            return None, None
        dx32 = torch.zeros(M, D, dtype=torch.float32, device=x.device)
        t32 = _sink32(table)
        dtable = None if t32 is not None else torch.zeros_like(table)
        for c0, n in chunks:
            _, n8 = project(c0, n)
            K.ce_chunk_grad(buf, target, c0, n, V, eps, lse, grad_scale=ProjCEFn.loss_scale)
            g = buf[:, :n8]
            # dx += dlogits_c W_c (fp32 accumulation across the chunks), dW_c = dlogits_c^T x
            K.gemm_raw(g.data_ptr(), table[c0:c0 + n8].data_ptr(), dx32, M, D, n8, lda=cols, ldb=D, ldc=D, b_mn_major=True,
                       accumulate=True, elem_dtype=x.dtype)
            if t32 is not None:
                K.linear_wgrad(g, x, out=t32[c0:c0 + n8], accumulate=True)
            else:
                K.linear_wgrad(g, x, out=dtable[c0:c0 + n8], accumulate=False)
        return dx32, dtable

    @staticmethod
    def backward(ctx, dloss):
        if ctx.chunked:
            dx32, dtable = ctx.saved_tensors
            if dx32 is None:
                raise RuntimeError('ProjCEFn: gradients were not computed in forward (it ran without grad)')
            if not ProjCEFn.unit_loss_grad:
                dx32 = dx32 * dloss
                if dtable is not None:
                    dtable = dtable * dloss.to(dtable.dtype)
                elif float(dloss) != 1.0:
                    raise RuntimeError('ProjCEFn: the table gradient already sits in the gradient arena; scale the loss instead')
            return dx32.to(ctx.elem_dtype), dtable, None, None, None, None, None
        x, table, dlogits = ctx.saved_tensors
        dx = K.linear_dgrad(dlogits, table)
        if ProjCEFn.unit_loss_grad:
            dtable = _wgrad(dlogits, x, ctx.table)
        else:
            dtable = K.linear_wgrad(dlogits, x) * dloss.to(x.dtype)
            dx = dx * dloss.to(dx.dtype)
        return dx, dtable, None, None, None, None, None


# functional front-ends ------------------------------------------------------------------------------
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    shape = x.shape
    y = LinearFn.apply(x.reshape(-1, shape[-1]), weight, bias)
    return y.view(*shape[:-1], weight.shape[0])


def ffn(x: Tensor, w1, b1, w2, b2, activation: str = 'relu') -> Tensor:
    shape = x.shape
    fn = FFNFn if activation == 'relu' else FFNGeluFn
    return fn.apply(x.reshape(-1, shape[-1]), w1, b1, w2, b2).view(shape)


def dropout_add_ln(x: Tensor, residual: Optional[Tensor], gamma, beta, p: float, eps: float) -> Tensor:
    shape = x.shape
    r2 = residual.reshape(-1, shape[-1]) if residual is not None else None
    return DropoutAddLNFn.apply(x.reshape(-1, shape[-1]), r2, gamma, beta, float(p), float(eps)).view(shape)


def attn_sublayer(x: Tensor, key: Optional[Tensor], attn, ln, kv_len, causal: bool, p: float, eps: float,
                  key_grad_cell: Optional[GradCell] = None) -> Tensor:
    """LN(x + dropout(attn(x, key, key))) for a `modules.MultiheadAttention` + nn.LayerNorm pair (post-norm).
    key_grad_cell: shared by every call that reads the same `key` tensor (see GradCell)."""
    B, T, D = x.shape
    S = T if key is None else key.shape[1]
    out = AttnSublayerFn.apply(x.reshape(B * T, D), None if key is None else key.reshape(B * S, D), attn.in_proj_weight,
                               attn.in_proj_bias, attn.out_proj.weight, attn.out_proj.bias, ln.weight, ln.bias, kv_len,
                               B, T, S, attn.num_heads, causal, attention_scale(attn.head_dim, attn.scaled), float(p), float(eps),
                               attn.key_bias, float(attn.dropout) if attn.training else 0.0, key_grad_cell)
    return out.view(B, T, D)


def ffn_sublayer(x: Tensor, fc1, fc2, ln, p: float, eps: float) -> Tensor:
    shape = x.shape
    out = FFNSublayerFn.apply(x.reshape(-1, shape[-1]), fc1.weight, fc1.bias, fc2.weight, fc2.bias, ln.weight, ln.bias,
                              float(p), float(eps))
    return out.view(shape)


def attention_scale(head_dim: int, scaled: bool = True) -> float:
    return 1.0 / math.sqrt(head_dim) if scaled else 1.0
