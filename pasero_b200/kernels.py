"""Raw tensor-level wrappers over the C ABI (no autograd). Each function checks shapes/dtypes on the
host, allocates outputs/workspaces with torch (device memory plumbing) and launches on torch's current
stream. Autograd-aware ops live in `pasero_b200.ops`."""
import math
from ctypes import byref
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import AttnParams, GemmParams, check, ptr, stream_ptr

BF16 = torch.bfloat16
ELEM_DTYPES = (torch.bfloat16, torch.float16)   # the two builds of the kernels (csrc/common.cuh): `dtype: bfloat16 | float16`

# device int64 scalar added (<< 12) to every dropout offset; set by the Trainer when it captures CUDA graphs so that
# replays draw fresh masks. None -> offsets are used as given.
RNG_OFFSET_BASE: Optional[Tensor] = None

# Instrumentation used by bench.py: PROFILE (list) receives (flops, start_event, end_event) per GEMM launch; COUNT
# ({'launches': n}) counts kernel launches issued through this module. Both None in normal operation.
PROFILE = None
PROFILE_TAGS = []          # one "m x n x k flags" string per PROFILE entry
PROFILE_EXTERNAL = False   # True: events become event-record nodes of a CUDA graph being captured (timed on replay)
COUNT = None


def _count(n: int) -> None:
    if COUNT is not None:
        COUNT['launches'] += n


PROFILE_ALL = None   # list of (tag, start_event, end_event) for every non-GEMM wrapper call (tools/step_breakdown.py)


PROFILE_WORK = None  # parallel to PROFILE_ALL: (unit, amount) algorithmic work of the call, e.g. ('flop', 3.1e9), or None


def _bracketed(fn=None, *, work=None):
    """Times a wrapper with external CUDA events when PROFILE_ALL is a list (instrumented graph capture). `work(*args,
    **kwargs) -> (unit, amount)`: the call's algorithmic FLOPs / bytes, for bench.py's per-family roofline table."""
    import functools
    if fn is None:
        return lambda f: _bracketed(f, work=work)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if PROFILE_ALL is None:
            return fn(*args, **kwargs)
        e0 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e1 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e0.record()
        out = fn(*args, **kwargs)
        e1.record()
        shapes = 'x'.join(str(d) for a in args[:2] if torch.is_tensor(a) for d in a.shape)
        PROFILE_ALL.append((f'{fn.__name__} {shapes}', e0, e1))
        if PROFILE_WORK is not None:
            PROFILE_WORK.append(work(*args, **kwargs) if work is not None else None)
        return out
    return wrapper


def _attn_flops(factor):
    # 4 B H Lq Lk d FLOPs forward (QK^T and PV), half of it under a causal mask; backward = 2.5x (five GEMMs instead of two)
    def work(*args, **kwargs):
        B, H, Lq, Lk = (args[3:7] if factor == 1.0 else args[6:10])
        causal = args[8] if factor == 1.0 else args[11]
        return 'flop', factor * 4.0 * B * H * Lq * Lk * 64 * (0.5 if causal else 1.0)
    return work


def _ln_bytes(*args, **kwargs):
    # dropout + residual + LayerNorm, forward or backward: four [M, D] bf16 streams
    M, D = args[0].shape
    return 'byte', 4.0 * M * D * 2


def _ws(nbytes: int, device) -> Optional[Tensor]:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device) if nbytes > 0 else None


# --------------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------------
def gemm_raw(a_ptr: int, b_ptr: int, c: Tensor, m: int, n: int, k: int, *, lda: int, ldb: int, ldc: int,
             batch: int = 1, a_batch_stride: int = 0, b_batch_stride: int = 0, c_batch_stride: int = 0,
             a_mn_major: bool = False, b_mn_major: bool = False, reduce_batch: bool = False,
             bias: Optional[Tensor] = None, residual: Optional[Tensor] = None, relu_mask: Optional[Tensor] = None,
             act: int = 0, accumulate: bool = False, alpha: float = 1.0, split_k: int = 0, tile_n: int = 0,
             cluster_m: int = 0, relu_bits_out: Optional[Tensor] = None, relu_bits: Optional[Tensor] = None,
             a_rowsum: Optional[Tensor] = None, elem_dtype=None) -> Tensor:
    """elem_dtype: the operands' 16-bit type (bfloat16 / float16 pick the library build); defaults to C's when C is 16-bit."""
    if elem_dtype is None:
        assert c.dtype != torch.float32, 'gemm_raw: an fp32 C does not tell the operands\' 16-bit type: pass elem_dtype'
        elem_dtype = c.dtype
    lib = _lib.load(elem_dtype)
    p = GemmParams()
    p.a_rowsum = ptr(a_rowsum)
    p.a, p.b, p.c = a_ptr, b_ptr, c.data_ptr()
    p.m, p.n, p.k, p.batch = m, n, k, batch
    p.lda, p.ldb, p.ldc = lda, ldb, ldc
    p.a_batch_stride, p.b_batch_stride, p.c_batch_stride = a_batch_stride, b_batch_stride, c_batch_stride
    p.a_mn_major, p.b_mn_major, p.reduce_batch = int(a_mn_major), int(b_mn_major), int(reduce_batch)
    p.bias, p.residual, p.relu_mask = ptr(bias), ptr(residual), ptr(relu_mask)
    p.relu_bits_out, p.relu_bits = ptr(relu_bits_out), ptr(relu_bits)
    bits = relu_bits_out if relu_bits_out is not None else relu_bits
    p.relu_bits_ld = bits.stride(0) if bits is not None else 0
    p.act, p.out_fp32, p.accumulate, p.alpha = act, int(c.dtype == torch.float32), int(accumulate), alpha
    p.split_k, p.tile_n, p.cluster_m = split_k, tile_n, cluster_m
    ws = None
    plain = bias is None and residual is None and relu_mask is None and relu_bits is None and act == 0
    direct = plain and accumulate and c.dtype == torch.float32 and alpha == 1.0   # reduce-add straight into C
    if split_k != 1 and plain and not direct:
        zout = 1 if reduce_batch else batch
        tiles = zout * math.ceil(m / 128) * math.ceil(n / 256 if n >= 256 else n / 128)
        sms = lib.pb200_num_sms()
        if split_k > 1 or tiles * 2 <= sms:
            ws = _ws(lib.pb200_gemm_workspace_bytes(m, n, zout, 1), c.device)
    if ws is None and not direct:
        p.split_k = 1
    p.workspace, p.workspace_bytes = ptr(ws), (ws.numel() if ws is not None else 0)
    _count(1 if ws is None else 2)
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e1 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e0.record()
        check(lib.pb200_gemm_bf16(byref(p), stream_ptr()))
        e1.record()
        PROFILE.append((2.0 * m * n * k * (batch if batch > 0 else 1), e0, e1))
        PROFILE_TAGS.append(f'{m}x{n}x{k} z{batch} a_mn{int(a_mn_major)} b_mn{int(b_mn_major)} bias{int(bias is not None)} act{act} '
                            f'mask{int(relu_mask is not None)} res{int(residual is not None)} splitk{int(ws is not None)}')
        return c
    check(lib.pb200_gemm_bf16(byref(p), stream_ptr()))
    return c


def relu_bits_supported(N: int, bias: Optional[Tensor]) -> bool:
    return N % 32 == 0 and bias is not None


def linear_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, act: int = 0,
               residual: Optional[Tensor] = None, out_dtype=None, tile_n: int = 0, cluster_m: int = 0,
               relu_bits_out: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """y[M, N] = act(x[M, K] @ w[N, K]^T + bias) (+ residual). relu_bits_out (int32 [N / 32, M], with act == 1): sign map
    of y for linear_dgrad(..., relu_bits=...)."""
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and x.dtype in ELEM_DTYPES and w.dtype == x.dtype and x.is_contiguous() and w.is_contiguous()
    y = out if out is not None else torch.empty(M, N, dtype=out_dtype or x.dtype, device=x.device)
    assert y.shape == (M, N) and y.is_contiguous()
    return gemm_raw(x.data_ptr(), w.data_ptr(), y, M, N, K, lda=K, ldb=K, ldc=N, bias=bias, act=act,
                    residual=residual, split_k=1, tile_n=tile_n, cluster_m=cluster_m, relu_bits_out=relu_bits_out, elem_dtype=x.dtype)


def linear_dgrad(dy: Tensor, w: Tensor, relu_mask: Optional[Tensor] = None, out: Optional[Tensor] = None,
                 accumulate: bool = False, residual: Optional[Tensor] = None, relu_bits: Optional[Tensor] = None) -> Tensor:
    """dx[M, K] = dy[M, N] @ w[N, K]  (optionally masked by relu_mask > 0, or + residual[M, K] in the epilogue:
    the gradient of the skip connection joins the sub-block's input gradient without a separate add kernel)"""
    M, N = dy.shape
    K = w.shape[1]
    dx = out if out is not None else torch.empty(M, K, dtype=dy.dtype, device=dy.device)
    assert residual is None or (residual.shape == dx.shape and residual.is_contiguous())
    return gemm_raw(dy.data_ptr(), w.data_ptr(), dx, M, K, N, lda=N, ldb=K, ldc=K, b_mn_major=True,
                    relu_mask=relu_mask, residual=residual, accumulate=accumulate, split_k=1, relu_bits=relu_bits, elem_dtype=dy.dtype)


def linear_wgrad(dy: Tensor, x: Tensor, out: Optional[Tensor] = None, accumulate: bool = False,
                 bias_grad32: Optional[Tensor] = None) -> Tensor:
    """dw[N, K] = dy[M, N]^T @ x[M, K]. bias_grad32 (fp32 [N], needs an fp32 `out` with accumulate): += column sums of
    dy, computed inside the same GEMM (one more 16-column MMA per k-step against an all-ones operand)."""
    M, N = dy.shape
    K = x.shape[1]
    dw = out if out is not None else torch.empty(N, K, dtype=dy.dtype, device=dy.device)
    if bias_grad32 is not None:
        assert bias_grad32.dtype == torch.float32 and bias_grad32.numel() == N and bias_grad32.is_contiguous()
        assert dw.dtype == torch.float32 and accumulate, 'fused bias gradient needs the fp32 accumulate path'
    return gemm_raw(dy.data_ptr(), x.data_ptr(), dw, N, K, M, lda=dy.stride(0), ldb=x.stride(0), ldc=dw.stride(0),
                    a_mn_major=True, b_mn_major=True, accumulate=accumulate, a_rowsum=bias_grad32, elem_dtype=dy.dtype)


# --------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------
def _attn_params(q, k, v, o, lse, kv_len, B, H, Lq, Lk, causal, scale, dropout, backward):
    """Fills the C struct; dropout = None or (p, seed, offset, mask uint32 [B, H, Lq, ceil(Lk / 32)]). Returns (params, keepalive)."""
    lib = _lib.load(q.dtype)
    p = AttnParams()
    p.q, p.k, p.v, p.o, p.lse, p.kv_len = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), ptr(kv_len)
    p.ldq, p.ldk, p.ldv, p.ldo = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    p.B, p.H, p.Lq, p.Lk, p.head_dim, p.causal, p.scale = B, H, Lq, Lk, 64, int(causal), scale
    if dropout is not None and dropout[0] > 0:
        p.dropout_p, p.seed, p.offset, p.offset_base, p.drop_mask = dropout[0], dropout[1], dropout[2], ptr(RNG_OFFSET_BASE), dropout[3].data_ptr()
    nbytes = lib.pb200_attention_workspace_bytes(byref(p), int(backward))
    ws = _ws(nbytes, q.device)
    p.workspace, p.workspace_bytes = ptr(ws), (ws.numel() if ws is not None else 0)
    return p, ws


def attention_mask_buffer(B: int, H: int, Lq: int, Lk: int, device='cuda') -> Tensor:
    """Keep-bit buffer of the attention dropout: one int32 word per (b, h, query, 32 keys); written by attention_fwd."""
    return torch.empty(B, H, Lq, (Lk + 31) // 32, dtype=torch.int32, device=device)


@_bracketed(work=_attn_flops(1.0))
def attention_fwd(q: Tensor, k: Tensor, v: Tensor, B: int, H: int, Lq: int, Lk: int, kv_len: Optional[Tensor],
                  causal: bool, scale: float, dropout=None):
    """q/k/v: 2-D bf16 views [B*L, >= H*64] with stride(1) == 1 (e.g. column slices of a fused projection).
    dropout: None or (p, seed, offset, mask) with mask = attention_mask_buffer(...) (filled here, read by attention_bwd)."""
    lib = _lib.load(q.dtype)
    o = torch.empty(B * Lq, H * 64, dtype=q.dtype, device=q.device)
    lse = torch.empty(B, H, Lq, dtype=torch.float32, device=q.device)
    p, ws = _attn_params(q, k, v, o, lse, kv_len, B, H, Lq, Lk, causal, scale, dropout, False)
    _count(1 if ws is None else 2)
    check(lib.pb200_attention_fwd(byref(p), stream_ptr()))
    return o, lse


@_bracketed(work=_attn_flops(2.5))
def attention_bwd(dout: Tensor, q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Tensor, B: int, H: int, Lq: int,
                  Lk: int, kv_len: Optional[Tensor], causal: bool, scale: float, dq: Tensor, dk: Tensor, dv: Tensor,
                  dropout=None):
    lib = _lib.load(q.dtype)
    p, ws = _attn_params(q, k, v, o, lse, kv_len, B, H, Lq, Lk, causal, scale, dropout, True)
    p.dout, p.lddo = dout.data_ptr(), dout.stride(0)
    p.dq, p.dk, p.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    p.lddq, p.lddk, p.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    _count(1 if ws is None else 4)
    check(lib.pb200_attention_bwd(byref(p), stream_ptr()))
    return dq, dk, dv


# --------------------------------------------------------------------------------------------------
# dropout + residual + LayerNorm
# --------------------------------------------------------------------------------------------------
@_bracketed(work=_ln_bytes)
def dropout_add_ln_fwd(x: Tensor, residual: Optional[Tensor], gamma: Optional[Tensor], beta: Optional[Tensor],
                       eps: float, p: float, seed: int, offset: int, need_sum: bool = True, keep_out: Optional[Tensor] = None):
    lib = _lib.load(x.dtype)
    M, D = x.shape
    dev = x.device
    has_ln = gamma is not None
    plain = residual is None and p == 0.0
    sum_out = None if (has_ln and (plain or not need_sum)) else torch.empty(M, D, dtype=x.dtype, device=dev)
    y = torch.empty(M, D, dtype=x.dtype, device=dev) if has_ln else None
    mean = torch.empty(M, dtype=torch.float32, device=dev) if has_ln else None
    rstd = torch.empty(M, dtype=torch.float32, device=dev) if has_ln else None
    _count(1)
    check(lib.pb200_dropout_add_ln_fwd(x.data_ptr(), ptr(residual), ptr(gamma), ptr(beta), M, D, eps, p, seed, offset,
                                       ptr(RNG_OFFSET_BASE), ptr(sum_out), ptr(y), ptr(mean), ptr(rstd), ptr(keep_out), stream_ptr()))
    if sum_out is None:
        sum_out = x  # LN of x alone: the saved pre-norm input is x itself
    return (y if has_ln else sum_out), sum_out, mean, rstd


def linear_ln_supported(M: int, N: int, K: int, residual, gamma) -> bool:
    """The fused Linear + dropout + residual + LayerNorm kernel (csrc/gemm_ln.cu) holds one output row per TMEM lane: d_model
    must be exactly 512 (all 512 accumulator columns). It runs one 256 x 512 tile per CTA pair, which pays off when the
    rows fill the 74 pairs (encoder side); the decoder's 3 072 rows stay on the separate kernels."""
    return N == 512 and K % 64 == 0 and M >= 8192 and M * N < 2 ** 32 and residual is not None and gamma is not None


def linear_dropout_add_ln_fwd(a: Tensor, w: Tensor, bias: Optional[Tensor], residual: Tensor, gamma: Tensor,
                              beta: Optional[Tensor], eps: float, p: float, seed: int, offset: int,
                              keep_out: Optional[Tensor] = None):
    """(y, sum, mean, rstd) of  sum = residual + dropout(a @ w^T + bias),  y = LayerNorm(sum) * gamma + beta  in ONE kernel:
    the Linear's output never reaches HBM. keep_out (uint8 [M * 512 / 8], required when p > 0) receives the dropout keep
    bits in the layout dropout_add_ln_bwd(keep_mask=...) reads."""
    lib = _lib.load(a.dtype)
    M, K = a.shape
    N = w.shape[0]
    assert linear_ln_supported(M, N, K, residual, gamma) and w.shape[1] == K
    assert a.dtype in ELEM_DTYPES and w.dtype == a.dtype and a.is_contiguous() and w.is_contiguous() and residual.is_contiguous()
    assert p == 0.0 or (keep_out is not None and keep_out.numel() == M * N // 8)
    dev = a.device
    y = torch.empty(M, N, dtype=a.dtype, device=dev)
    s = torch.empty(M, N, dtype=a.dtype, device=dev)
    mean = torch.empty(M, dtype=torch.float32, device=dev)
    rstd = torch.empty(M, dtype=torch.float32, device=dev)
    _count(1)
    args = (a.data_ptr(), w.data_ptr(), ptr(bias), residual.data_ptr(), gamma.data_ptr(), ptr(beta), M, K, eps, p, seed, offset,
            ptr(RNG_OFFSET_BASE), s.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), ptr(keep_out), stream_ptr())
    if PROFILE is not None:    # timed with the GEMMs (its FLOPs are the Linear's; the LayerNorm tail rides along)
        e0 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e1 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e0.record()
        check(lib.pb200_linear_dropout_add_ln_fwd(*args))
        e1.record()
        PROFILE.append((2.0 * M * N * K, e0, e1))
        PROFILE_TAGS.append(f'{M}x{N}x{K} linear+dropout+residual+LN fused')
    else:
        check(lib.pb200_linear_dropout_add_ln_fwd(*args))
    return y, s, mean, rstd


@_bracketed(work=_ln_bytes)
def dropout_add_ln_bwd(dy: Tensor, dsum_extra: Optional[Tensor], sum_: Optional[Tensor], gamma: Optional[Tensor],
                       mean, rstd, p: float, seed: int, offset: int, need_dresidual: bool,
                       dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None, accumulate: bool = False,
                       dbias_lin: Optional[Tensor] = None, accumulate_lin: bool = False, keep_mask: Optional[Tensor] = None,
                       side_run=None):
    """dbias_lin (bf16 [D], optional) receives the column sums of dx = the bias gradient of the Linear that produced x.
    side_run(fn, *keepalive): when given, the small kernel that reduces the CTA partials into dgamma / dbeta / dbias_lin is
    handed to it (ops.WgradStream.run: another stream) together with its private partials buffer."""
    lib = _lib.load(dy.dtype)
    M, D = dy.shape
    dev = dy.device
    dx = torch.empty(M, D, dtype=dy.dtype, device=dev)
    dres = torch.empty(M, D, dtype=dy.dtype, device=dev) if (need_dresidual and p > 0.0) else (dx if need_dresidual else None)
    need_ws = (gamma is not None and (dgamma is not None or dbeta is not None)) or dbias_lin is not None
    defer = need_ws and side_run is not None
    ws = None
    if need_ws:   # deferred reduction: the buffer must outlive this call on the launching stream -> not the shared scratch
        nbytes = lib.pb200_ln_bwd_workspace_bytes(M, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev) if defer else _ws(nbytes, dev)
    _count(2 if need_ws else 1)
    flags = int(accumulate) | (int(accumulate_lin) << 1)
    check(lib.pb200_dropout_add_ln_bwd(dy.data_ptr(), ptr(dsum_extra), ptr(sum_), ptr(gamma), ptr(mean), ptr(rstd), M, D,
                                       p, seed, offset, ptr(RNG_OFFSET_BASE), ptr(dres), dx.data_ptr(), ptr(dgamma), ptr(dbeta),
                                       ptr(dbias_lin), flags | (4 if defer else 0), ptr(keep_mask), ptr(ws),
                                       ws.numel() if ws is not None else 0, stream_ptr()))
    if defer:
        has_ln = gamma is not None
        side_run(lambda: check(lib.pb200_ln_bwd_reduce(ws.data_ptr(), M, D, ptr(dgamma) if has_ln else None,
                                                       ptr(dbeta) if has_ln else None, ptr(dbias_lin), flags, stream_ptr())), ws)
    return dx, dres


@_bracketed
def colsum(x: Tensor, out: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
    lib = _lib.load(x.dtype)
    M, N = x.shape
    out = out if out is not None else torch.empty(N, dtype=x.dtype, device=x.device)
    ws = _ws(lib.pb200_colsum_workspace_bytes(M, N), x.device)
    _count(2)
    fn = lib.pb200_colsum_f32 if out.dtype == torch.float32 else lib.pb200_colsum_bf16   # fp32: main gradient of a bias
    check(fn(x.data_ptr(), M, N, x.stride(0), out.data_ptr(), int(accumulate), ws.data_ptr(), ws.numel(), stream_ptr()))
    return out


# --------------------------------------------------------------------------------------------------
# embedding epilogue, conv-subsampler pieces
# --------------------------------------------------------------------------------------------------
@_bracketed
def embed_fwd(ids: Optional[Tensor], src: Tensor, pos: Optional[Tensor], M: int, D: int, L: int, scale: float,
              p: float, seed: int, offset: int) -> Tensor:
    lib = _lib.load(src.dtype)
    out = torch.empty(M, D, dtype=src.dtype, device=src.device)
    _count(1)
    check(lib.pb200_embed_fwd(ptr(ids), src.data_ptr(), ptr(pos), M, D, L, scale, p, seed, offset,
                              ptr(RNG_OFFSET_BASE), src.shape[0] if ids is not None else 0, out.data_ptr(), stream_ptr()))
    return out


@_bracketed
def embed_bwd(dy: Tensor, ids: Optional[Tensor], M: int, D: int, scale: float, p: float, seed: int, offset: int,
              padding_idx: int = -1, dtable: Optional[Tensor] = None) -> Optional[Tensor]:
    lib = _lib.load(dy.dtype)
    dx = torch.empty(M, D, dtype=dy.dtype, device=dy.device) if ids is None else None
    _count(1)
    check(lib.pb200_embed_bwd(dy.data_ptr(), ptr(ids), M, D, scale, p, seed, offset, ptr(RNG_OFFSET_BASE), padding_idx,
                              dtable.shape[0] if dtable is not None else 0, ptr(dx), ptr(dtable), stream_ptr()))
    return dx


@_bracketed
def glu_fwd(x: Tensor, B: int, T: int, C: int, Tp: int, front_pad: int) -> Tensor:
    out = torch.empty(B, Tp, C, dtype=x.dtype, device=x.device)
    _count(1)
    check(_lib.load(x.dtype).pb200_glu_fwd(x.data_ptr(), B, T, C, out.data_ptr(), Tp, front_pad, stream_ptr()))
    return out


@_bracketed
def glu_bwd(dy: Tensor, x: Tensor, M: int, C: int) -> Tensor:
    din = torch.empty(M, 2 * C, dtype=x.dtype, device=x.device)
    _count(1)
    check(_lib.load(x.dtype).pb200_glu_bwd(dy.data_ptr(), x.data_ptr(), M, C, din.data_ptr(), stream_ptr()))
    return din


@_bracketed
def gelu_fwd(x: Tensor, B: int, T: int, C: int, Tp: int, front_pad: int) -> Tensor:
    """gelu(x [B*T, C]) written into rows [front_pad, front_pad + T) of a zero-padded [B, Tp, C] tensor."""
    out = torch.empty(B, Tp, C, dtype=x.dtype, device=x.device)
    _count(1)
    check(_lib.load(x.dtype).pb200_gelu_fwd(x.data_ptr(), B, T, C, out.data_ptr(), Tp, front_pad, stream_ptr()))
    return out


@_bracketed
def gelu_bwd(dy: Tensor, pre: Tensor) -> Tensor:
    """dx = dy * gelu'(pre), elementwise over contiguous bf16 tensors of the same shape."""
    assert dy.is_contiguous() and pre.is_contiguous() and dy.numel() == pre.numel()
    dx = torch.empty_like(dy)
    _count(1)
    check(_lib.load(dy.dtype).pb200_gelu_bwd(dy.data_ptr(), pre.data_ptr(), dy.numel(), dx.data_ptr(), stream_ptr()))
    return dx


@_bracketed
def relu_bwd(dy: Tensor, h: Tensor) -> Tensor:
    """dx = dy * (h > 0) for contiguous bf16 tensors of the same shape (h = the ReLU's output)."""
    assert dy.is_contiguous() and h.is_contiguous() and dy.numel() == h.numel()
    dx = torch.empty_like(dy)
    _count(1)
    check(_lib.load(dy.dtype).pb200_relu_bwd(dy.data_ptr(), h.data_ptr(), dy.numel(), dx.data_ptr(), stream_ptr()))
    return dx


@_bracketed
def col2im(dcol: Tensor, B: int, Tin: int, Tout: int, C: int, KW: int, stride: int) -> Tensor:
    dx = torch.empty(B * Tin, C, dtype=dcol.dtype, device=dcol.device)
    _count(1)
    check(_lib.load(dcol.dtype).pb200_col2im(dcol.data_ptr(), B, Tin, Tout, C, KW, stride, dx.data_ptr(), stream_ptr()))
    return dx


@_bracketed
def pad_rows(x: Tensor, B: int, T: int, C: int, Tp: int, front_pad: int) -> Tensor:
    out = torch.empty(B, Tp, C, dtype=x.dtype, device=x.device)
    _count(1)
    check(_lib.load(x.dtype).pb200_pad_rows(x.data_ptr(), B, T, C, out.data_ptr(), Tp, front_pad, stream_ptr()))
    return out


@_bracketed
def collate_features(rows: Tensor, offsets: Tensor, B: int, Tp: int, front_pad: int = 0, out_dtype=BF16) -> Tensor:
    """rows: fp16 / fp32 [total_rows, C] on the device (utterances back to back), offsets int64 [B + 1] (device, in rows).
    Returns bf16 [B, Tp, C]: utterance b at rows [front_pad, front_pad + len_b), zeros elsewhere."""
    assert rows.dtype in (torch.float16, torch.float32) and rows.is_contiguous() and offsets.dtype == torch.int64
    C = rows.shape[1]
    out = torch.empty(B, Tp, C, dtype=out_dtype, device=rows.device)
    _count(1)
    check(_lib.load(out_dtype).pb200_collate_features(rows.data_ptr(), int(rows.dtype == torch.float16), offsets.data_ptr(), B, C,
                                             out.data_ptr(), Tp, front_pad, stream_ptr()))
    return out


@_bracketed
def f32_to_bf16(src: Tensor, out: Tensor, accumulate: bool = False) -> Tensor:
    _count(1)
    check(_lib.load(out.dtype).pb200_f32_to_bf16(src.data_ptr(), src.numel(), out.data_ptr(), int(accumulate), stream_ptr()))
    return out


# --------------------------------------------------------------------------------------------------
# loss, optimizer, frontend
# --------------------------------------------------------------------------------------------------
def _ce_bytes(logits, target, *args, **kwargs):
    # label-smoothed CE in one pass: the bf16 logits are read once and overwritten with their gradient
    vocab = kwargs.get('vocab') or (args[6] if len(args) > 6 and args[6] else logits.shape[1])
    return 'byte', 2.0 * logits.shape[0] * vocab * 2


@_bracketed(work=_ce_bytes)
def ls_ce_fwd_bwd(logits: Tensor, target: Tensor, padding_idx: int, label_smoothing: float, stats: Tensor,
                  write_grad: bool = True, grad_scale: float = 1.0, row_loss: Optional[Tensor] = None,
                  vocab: Optional[int] = None) -> None:
    """logits [B*T, >= V] bf16 (columns [0, V) overwritten with the gradient), target [B, T] int64 (may be a strided
    view). `vocab` = V when the logits carry padding columns (table rows rounded up to a multiple of 8)."""
    B, T = target.shape
    assert target.stride(1) == 1 and target.dtype == torch.int64
    V = vocab if vocab is not None else logits.shape[1]
    assert V <= logits.shape[1]
    _count(1)
    check(_lib.load(logits.dtype).pb200_ls_ce_fwd_bwd(logits.data_ptr(), target.data_ptr(), target.stride(0), B, T, V,
                                          logits.stride(0), padding_idx, label_smoothing, grad_scale, int(write_grad),
                                          stats.data_ptr(), ptr(row_loss), stream_ptr()))


# vocabularies of at least this many entries take the chunked projection + cross-entropy (ops.ProjCEFn): no [B*T, V] tensor
# in HBM; CE_CHUNK_COLS logits columns are alive at a time (L2 resident at BASELINE batch sizes)
CE_CHUNK_MIN_VOCAB = 32768
CE_CHUNK_COLS = 8192


@_bracketed
def ce_chunk_stats(logits: Tensor, target: Tensor, c0: int, cols: int, padding_idx: int, row_state: Tensor) -> None:
    B, T = target.shape
    _count(1)
    check(_lib.load(logits.dtype).pb200_ce_chunk_stats(logits.data_ptr(), logits.stride(0), target.data_ptr(), target.stride(0), B, T, c0, cols,
                                           padding_idx, row_state.data_ptr(), stream_ptr()))


@_bracketed
def ce_chunk_finish(row_state: Tensor, target: Tensor, V: int, padding_idx: int, label_smoothing: float, stats: Tensor) -> Tensor:
    B, T = target.shape
    lse = torch.empty(B * T, dtype=torch.float32, device=row_state.device)
    _count(1)
    check(_lib.load(None).pb200_ce_chunk_finish(row_state.data_ptr(), target.data_ptr(), target.stride(0), B, T, V, padding_idx,
                                            label_smoothing, stats.data_ptr(), lse.data_ptr(), stream_ptr()))
    return lse


@_bracketed
def ce_chunk_grad(logits: Tensor, target: Tensor, c0: int, cols: int, V: int, label_smoothing: float, lse: Tensor,
                  grad_scale: float = 1.0) -> None:
    B, T = target.shape
    _count(1)
    check(_lib.load(logits.dtype).pb200_ce_chunk_grad(logits.data_ptr(), logits.stride(0), target.data_ptr(), target.stride(0), B, T, c0, cols, V,
                                          label_smoothing, grad_scale, lse.data_ptr(), stream_ptr()))


_SQNORM_WS = {}


@_bracketed
def grad32_flush(acc32: Tensor, grad: Tensor, ranges: Tensor, max_len: int) -> None:
    """bf16 gradient arena <- rounded fp32 main gradients over `ranges` (int64 [n, 2], device); zeroes the source."""
    _count(1)
    check(_lib.load(grad.dtype).pb200_grad32_flush(acc32.data_ptr(), grad.data_ptr(), ranges.data_ptr(), ranges.shape[0], max_len,
                                         stream_ptr()))


@_bracketed
def grad_sqnorm(grad: Tensor, hyper: Tensor, stats: Tensor) -> None:
    lib = _lib.load(grad.dtype)
    ws = _SQNORM_WS.get(grad.device)
    if ws is None:   # persistent, zero-initialised once (the kernel re-arms its ticket)
        ws = _SQNORM_WS[grad.device] = torch.zeros(lib.pb200_grad_sqnorm_workspace_bytes(), dtype=torch.uint8, device=grad.device)
    _count(1)
    check(lib.pb200_grad_sqnorm(grad.data_ptr(), grad.numel(), hyper.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws.numel(),
                                stream_ptr()))


@_bracketed
def adam_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, hyper: Tensor, stats: Tensor,
              master: Optional[Tensor] = None) -> None:
    """master (fp32, same length): the update goes to it and `param` receives its rounded image (MixedPrecisionAdam /
    FlatFP16Adam); None: in-place update with the 16-bit rounding points of optimization.Adam."""
    _count(1)
    check(_lib.load(param.dtype).pb200_adam_step(param.data_ptr(), ptr(master), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                      param.numel(), hyper.data_ptr(), stats.data_ptr(), stream_ptr()))


def fbank_num_frames(num_samples: int) -> int:
    return int(_lib.load(None).pb200_fbank_num_frames(num_samples))


_FBANK_WS = {}


@_bracketed
def fbank_cmvn(pcm: Tensor, sample_offsets: Tensor, frame_offsets: Tensor, batch: int, max_frames: int,
               total_frames: int, pcm_scale: float = 32768.0, cmvn: bool = True, padded_frames: int = 0,
               front_pad: int = 0, keep_fp32: bool = True, out_dtype=BF16):
    """pcm fp32 [total_samples]; offsets int64 [batch+1] on device. Returns (feats fp32 [total_frames, 80],
    bf16 padded [batch, padded_frames, 80] or None). keep_fp32=False (needs the padded output and cmvn): the fp32 rows are
    only the kernel's scratch (raw log-mel), the normalised features exist in the bf16 padded layout alone."""
    lib = _lib.load(out_dtype)
    dev = pcm.device
    assert pcm.dtype == torch.float32 and pcm.is_contiguous()
    feats = torch.empty(total_frames, 80, dtype=torch.float32, device=dev)
    padded = torch.empty(batch, padded_frames, 80, dtype=out_dtype, device=dev) if padded_frames > 0 else None
    # statistics + tile counters: zero on first use, left zeroed by every call (see pb200_fbank_workspace_bytes) -> one
    # persistent buffer per (device, stream, batch), no memset in the step
    key = (dev, stream_ptr(), batch)
    ws = _FBANK_WS.get(key)
    if ws is None:
        ws = _FBANK_WS[key] = torch.zeros(max(int(lib.pb200_fbank_workspace_bytes(batch)), 16), dtype=torch.uint8, device=dev)
    mode = int(cmvn) | (2 if (cmvn and not keep_fp32 and padded is not None) else 0)
    _count(2 if (cmvn or padded is not None) else 1)
    check(lib.pb200_fbank_cmvn(pcm.data_ptr(), sample_offsets.data_ptr(), frame_offsets.data_ptr(), batch, max_frames,
                               pcm_scale, mode, feats.data_ptr(), ptr(padded), padded_frames, front_pad,
                               ws.data_ptr(), ws.numel(), stream_ptr()))
    return feats, padded
