"""Audio frontend: batched Kaldi-compatible 80-bin log-mel fbank + utterance CMVN on the GPU.

The reference consumes pre-extracted features (pasero/files.py:103-192); this replaces the offline
torchaudio.compliance.kaldi.fbank + CMVN extraction (fairseq-S2T convention) with one fused kernel pass over the
PCM, and can hand the encoder its input directly in the conv subsampler's zero-padded bf16 layout.
"""
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import kernels as K

NUM_MEL = 80


def num_frames(num_samples: int) -> int:
    return K.fbank_num_frames(num_samples)


class FbankPlan:
    """Offsets for a batch of utterances (host ints -> device int64 arrays), reusable across steps of equal shape."""

    def __init__(self, lengths: Sequence[int], device):
        self.lengths = [int(n) for n in lengths]
        self.frames = [num_frames(n) for n in self.lengths]
        so, fo = [0], [0]
        for n, f in zip(self.lengths, self.frames):
            so.append(so[-1] + n)
            fo.append(fo[-1] + f)
        self.total_samples, self.total_frames = so[-1], fo[-1]
        self.max_frames = max(self.frames) if self.frames else 0
        self.batch = len(self.lengths)
        self.sample_offsets = torch.tensor(so, dtype=torch.int64, device=device)
        self.frame_offsets = torch.tensor(fo, dtype=torch.int64, device=device)
        self.frames_tensor = torch.tensor(self.frames, dtype=torch.int64, device=device)


def fbank(pcm: Tensor, plan: FbankPlan, cmvn: bool = True, pcm_scale: float = 32768.0,
          padded_frames: int = 0, front_pad: int = 0, keep_fp32: bool = True, dtype=torch.bfloat16):
    """pcm: fp32 [total_samples] on the GPU (utterances concatenated, samples in [-1, 1)).
    Returns (features fp32 [total_frames, 80], 16-bit [B, padded_frames, 80] or None). keep_fp32=False: only the padded 16-bit
    layout (`dtype`: bfloat16, or float16 for a `dtype: float16` model) holds the normalised features (what the training step
    consumes); the fp32 tensor is scratch."""
    if not pcm.is_cuda:
        raise RuntimeError('pasero_b200.frontend.fbank needs the PCM on the GPU (no CPU fallback)')
    return K.fbank_cmvn(pcm, plan.sample_offsets, plan.frame_offsets, plan.batch, plan.max_frames, plan.total_frames,
                        pcm_scale=pcm_scale, cmvn=cmvn, padded_frames=padded_frames, front_pad=front_pad, keep_fp32=keep_fp32,
                        out_dtype=dtype)


def fbank_batch(pcm: Tensor, cmvn: bool = True) -> Tensor:
    """Equal-length convenience form: pcm (B, N) -> features (B, F, 80) fp32."""
    B, N = pcm.shape
    plan = FbankPlan([N] * B, pcm.device)
    feats, _ = fbank(pcm.reshape(-1).contiguous(), plan, cmvn=cmvn)
    return feats.view(B, plan.max_frames, NUM_MEL)


def collate_features(packed_rows, offsets, padded_rows: int, front_pad: int = 0, device='cuda', stream=None, dtype=torch.bfloat16):
    """GPU fast path for pre-extracted features (the reference's NumpyFile -> Task.collate -> move_to_device chain,
    pasero/files.py:164-175, pasero/tasks/task.py:531-571, pasero/utils.py:691-736): `packed_rows` / `offsets` as returned by
    `pasero_b200.task.NumpyFile.pack` (ragged rows back to back in the file's dtype, ideally in a pinned buffer). ONE
    host->device copy of the true bytes, then one kernel pads + casts into the bf16 batch tensor — with
    (padded_rows, front_pad) = `ConvolutionSubsampler.padded_rows(max_len)` directly in the conv subsampler's input layout
    (pass `padded_layout=(max_len,)` to the model). Returns (bf16 [B, padded_rows, C], lengths int64 [B] on the device)."""
    rows = torch.as_tensor(packed_rows)
    offs = torch.as_tensor(offsets, dtype=torch.int64)
    rows_d = rows.to(device, non_blocking=True)
    offs_d = offs.to(device, non_blocking=True)
    B = offs.numel() - 1
    return K.collate_features(rows_d, offs_d, B, padded_rows, front_pad, out_dtype=dtype), (offs_d[1:] - offs_d[:-1])
