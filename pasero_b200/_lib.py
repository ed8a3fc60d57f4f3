"""ctypes binding of libpasero_b200.so (the C ABI declared in include/pasero_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised with the
library's own error text. Tensors cross the boundary as raw device pointers + sizes; the CUDA stream is
torch's current stream.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_uint64,
                    c_void_p)

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpasero_b200.so')


class GemmParams(Structure):
    _fields_ = [
        ('a', c_void_p), ('b', c_void_p), ('c', c_void_p),
        ('m', c_int64), ('n', c_int64), ('k', c_int64), ('batch', c_int64),
        ('lda', c_int64), ('ldb', c_int64), ('ldc', c_int64),
        ('a_batch_stride', c_int64), ('b_batch_stride', c_int64), ('c_batch_stride', c_int64),
        ('a_mn_major', c_int32), ('b_mn_major', c_int32), ('reduce_batch', c_int32),
        ('bias', c_void_p), ('residual', c_void_p), ('relu_mask', c_void_p),
        ('relu_bits_out', c_void_p), ('relu_bits', c_void_p), ('relu_bits_ld', c_int64),
        ('act', c_int32), ('out_fp32', c_int32), ('accumulate', c_int32), ('alpha', c_float),
        ('split_k', c_int32), ('tile_n', c_int32), ('cluster_m', c_int32),
        ('workspace', c_void_p), ('workspace_bytes', c_size_t),
        ('a_rowsum', c_void_p),
    ]


class AttnParams(Structure):
    _fields_ = [
        ('q', c_void_p), ('k', c_void_p), ('v', c_void_p), ('o', c_void_p),
        ('lse', c_void_p), ('kv_len', c_void_p),
        ('ldq', c_int64), ('ldk', c_int64), ('ldv', c_int64), ('ldo', c_int64),
        ('B', c_int32), ('H', c_int32), ('Lq', c_int32), ('Lk', c_int32), ('head_dim', c_int32),
        ('causal', c_int32), ('scale', c_float),
        ('dout', c_void_p), ('lddo', c_int64), ('delta', c_void_p),
        ('dq', c_void_p), ('dk', c_void_p), ('dv', c_void_p),
        ('lddq', c_int64), ('lddk', c_int64), ('lddv', c_int64),
        ('dropout_p', c_float), ('seed', c_uint64), ('offset', c_uint64), ('offset_base', c_void_p), ('drop_mask', c_void_p),
        ('workspace', c_void_p), ('workspace_bytes', c_size_t),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/pasero_b200.h
_P = c_void_p
PROTOTYPES = {
    'pb200_abi_version': (c_int32, []),
    'pb200_last_error_string': (c_char_p, []),
    'pb200_num_sms': (c_int32, []),
    'pb200_fbank_init': (c_int32, []),
    'pb200_fbank_num_frames': (c_int64, [c_int64]),
    'pb200_fbank_workspace_bytes': (c_size_t, [c_int32]),
    'pb200_fbank_cmvn': (c_int32, [_P, _P, _P, c_int32, c_int32, c_float, c_int32, _P, _P, c_int32, c_int32, _P,
                                   c_size_t, _P]),
    'pb200_gemm_workspace_bytes': (c_size_t, [c_int64, c_int64, c_int64, c_int32]),
    'pb200_gemm_bf16': (c_int32, [POINTER(GemmParams), _P]),
    'pb200_attention_workspace_bytes': (c_size_t, [POINTER(AttnParams), c_int32]),
    'pb200_attention_fwd': (c_int32, [POINTER(AttnParams), _P]),
    'pb200_attention_bwd': (c_int32, [POINTER(AttnParams), _P]),
    'pb200_dropout_add_ln_fwd': (c_int32, [_P, _P, _P, _P, c_int64, c_int64, c_float, c_float, c_uint64, c_uint64,
                                           _P, _P, _P, _P, _P, _P, _P]),
    'pb200_linear_dropout_add_ln_fwd': (c_int32, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_float, c_float, c_uint64, c_uint64,
                                                  _P, _P, _P, _P, _P, _P, _P]),
    'pb200_ln_bwd_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'pb200_ln_bwd_reduce': (c_int32, [_P, c_int64, c_int64, _P, _P, _P, c_int32, _P]),
    'pb200_dropout_add_ln_bwd': (c_int32, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_float, c_uint64, c_uint64,
                                           _P, _P, _P, _P, _P, _P, c_int32, _P, _P, c_size_t, _P]),
    'pb200_colsum_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'pb200_colsum_bf16': (c_int32, [_P, c_int64, c_int64, c_int64, _P, c_int32, _P, c_size_t, _P]),
    'pb200_colsum_f32': (c_int32, [_P, c_int64, c_int64, c_int64, _P, c_int32, _P, c_size_t, _P]),
    'pb200_embed_fwd': (c_int32, [_P, _P, _P, c_int64, c_int64, c_int64, c_float, c_float, c_uint64, c_uint64,
                                  _P, c_int64, _P, _P]),
    'pb200_embed_bwd': (c_int32, [_P, _P, c_int64, c_int64, c_float, c_float, c_uint64, c_uint64, _P, c_int64, c_int64,
                                  _P, _P, _P]),
    'pb200_glu_fwd': (c_int32, [_P, c_int32, c_int32, c_int32, _P, c_int32, c_int32, _P]),
    'pb200_glu_bwd': (c_int32, [_P, _P, c_int64, c_int32, _P, _P]),
    'pb200_gelu_fwd': (c_int32, [_P, c_int32, c_int32, c_int32, _P, c_int32, c_int32, _P]),
    'pb200_gelu_bwd': (c_int32, [_P, _P, c_int64, _P, _P]),
    'pb200_relu_bwd': (c_int32, [_P, _P, c_int64, _P, _P]),
    'pb200_col2im': (c_int32, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    'pb200_pad_rows': (c_int32, [_P, c_int32, c_int32, c_int32, _P, c_int32, c_int32, _P]),
    'pb200_f32_to_bf16': (c_int32, [_P, c_int64, _P, c_int32, _P]),
    'pb200_collate_features': (c_int32, [_P, c_int32, _P, c_int32, c_int32, _P, c_int32, c_int32, _P]),
    'pb200_ls_ce_fwd_bwd': (c_int32, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_float,
                                      c_int32, _P, _P, _P]),
    'pb200_ce_chunk_stats': (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P]),
    'pb200_ce_chunk_finish': (c_int32, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, _P, _P, _P]),
    'pb200_ce_chunk_grad': (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_float, _P, _P]),
    'pb200_grad32_flush': (c_int32, [_P, _P, _P, c_int32, c_int64, _P]),
    'pb200_grad_sqnorm_workspace_bytes': (c_size_t, []),
    'pb200_grad_sqnorm': (c_int32, [_P, c_int64, _P, _P, _P, c_size_t, _P]),
    'pb200_adam_step': (c_int32, [_P, _P, _P, _P, _P, c_int64, _P, _P, _P]),
}

_lib = None


class _HalfBuild:
    """View of the library whose entry points resolve to the float16 build of the kernels (suffix _f16, csrc/f16_symbols.h);
    the library-wide ones (error string, SM count, ABI version) exist once."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name + '_f16', None) if name not in SHARED else None
        fn = fn if fn is not None else getattr(self._lib, name)
        setattr(self, name, fn)
        return fn


SHARED = ('pb200_abi_version', 'pb200_last_error_string', 'pb200_num_sms')
_half = None


def load(dtype=None):
    """Loads the shared library (once) and installs the prototypes. Raises if it is missing. dtype torch.float16 returns the
    float16 build's entry points under the same names; anything else (None, bfloat16, float32) the bfloat16 build's."""
    global _lib, _half
    if _lib is not None:
        return _half if dtype == torch.float16 else _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -m pasero_b200.build` '
            '(pasero_b200 has no CPU or PyTorch fallback for its CUDA kernels)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        for sym in ((name,) if name in SHARED else (name, name + '_f16')):
            fn = getattr(lib, sym)  # AttributeError here = ABI mismatch between header and library
            fn.restype = restype
            fn.argtypes = argtypes
    if lib.pb200_abi_version() != 1:
        raise RuntimeError('libpasero_b200.so ABI version mismatch')
    _lib, _half = lib, _HalfBuild(lib)
    return _half if dtype == torch.float16 else _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().pb200_last_error_string()
        raise RuntimeError(f'pasero_b200: {msg.decode() if msg else "unknown error"} (code {rc})')


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('pasero_b200 kernels need CUDA tensors: there is no CPU fallback')
