"""ORACLE (test infrastructure, not product code): CPU restatement of Pasero's length-bucketed batching
and collation for the speech-translation task. Integer work: parity is bit-exact.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

  build_batches          pasero/utils.py:783-843
  task_build_batches     pasero/tasks/task.py:573-610      (encoder_decoder branch; global numpy RNG call order)
  buffered_batching      pasero/datasets.py:708-730
  tokens_as_tensor       pasero/utils.py:709-736
  mask_to_len/len_to_mask pasero/utils.py:246-268
  collate                pasero/tasks/task.py:531-571 + pasero/tasks/translation.py:876-899
  should_skip            pasero/tasks/translation.py:320-328

PARITY PIN: the reference's own docstring examples (utils.py:797-813, :253, :265) and
tests/golden/batching.npz (index lists produced by the unmodified reference on the seeded 5000-sample set of
SURVEY.md §8d, all eight constraint combinations).
"""
import collections
from typing import Callable, Optional

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence


def build_batches(indices, size_fn: Callable, batch_size: int, batch_size_multiple: Optional[int] = None,
                  max_lines: Optional[int] = None) -> list:
    batch_size_multiple = batch_size_multiple or 1
    batches, batch, lengths = [], [], []
    max_len = 0
    for idx in indices:
        size = size_fn(idx)
        if size > batch_size:  # too big to fit alone in a batch: dropped
            continue
        lengths.append(size)
        max_len = max(max_len, size)
        if (max_lines and len(batch) >= max_lines) or max_len * (len(batch) + 1) > batch_size:
            multiple_len = max(batch_size_multiple * (len(batch) // batch_size_multiple), len(batch) % batch_size_multiple)
            batches.append(batch[:multiple_len])
            batch = batch[multiple_len:]
            lengths = lengths[multiple_len:]
            max_len = max(lengths) if lengths else 0
        batch.append(idx)
    if batch:
        batches.append(batch)
    return batches


def task_build_batches(data: list, batch_size: int, batch_size_multiple: Optional[int] = None,
                       lines_per_batch: Optional[int] = None, shuffle: bool = True, sort: bool = True) -> list:
    encoder_length = np.array([len(sample['encoder_input']) for sample in data])
    decoder_length = np.array([len(sample['decoder_input']) for sample in data])
    length = np.maximum(encoder_length, decoder_length)
    indices = np.random.permutation(len(data)) if shuffle else np.arange(len(data))
    if sort:
        indices = indices[np.argsort(decoder_length[indices], kind='stable')]
        indices = indices[np.argsort(encoder_length[indices], kind='stable')]
    batch_indices = build_batches(indices, length.__getitem__, batch_size, batch_size_multiple, lines_per_batch)
    return [[{**data[i], 'index': i} for i in idx] for idx in batch_indices]


def buffered_batching(buffer: list, batch_size: int, batch_by=(), batch_size_multiple=None, lines_per_batch=None,
                      shuffle: bool = True) -> list:
    groups = collections.defaultdict(list)
    for sample in buffer:
        key = tuple(sample['meta'].get(k) for k in batch_by)
        if sample.get('encoder_input') is not None:
            key = key + (sample['encoder_input'].ndim, sample['encoder_input'].dtype)
        groups[key].append(sample)
    batches = []
    for group in groups.values():
        batches += task_build_batches(group, batch_size, batch_size_multiple, lines_per_batch, shuffle=shuffle)
    if shuffle:
        np.random.shuffle(batches)
    return batches


def mask_to_len(mask: torch.Tensor, dim: int = -1) -> torch.Tensor:
    return (mask.flip(dims=(dim,)).cumsum(dim) > 0).sum(dim)


def len_to_mask(lengths: torch.Tensor, size: Optional[int] = None) -> torch.Tensor:
    size = size or lengths.max()
    return torch.arange(size, device=lengths.device).unsqueeze(0) >= lengths.unsqueeze(1)


def tokens_as_tensor(token_list: list, padding_idx: int, dtype: torch.dtype = None):
    token_list = [torch.as_tensor(tokens) for tokens in token_list]
    if token_list[0].is_floating_point():
        tokens = pad_sequence(token_list, batch_first=True, padding_value=0.0).to(dtype)
    elif token_list[0].dtype is torch.bool:
        tokens = pad_sequence(token_list, batch_first=True, padding_value=False).bool()
    else:
        tokens = pad_sequence(token_list, batch_first=True, padding_value=padding_idx).long()
    lengths = torch.LongTensor(list(map(len, token_list)))
    return tokens, lengths


def collate(batch: list, padding_idx: int, dtype: torch.dtype):
    if not batch:
        return None
    decoder_input, _ = tokens_as_tensor([s['decoder_input'] for s in batch], padding_idx, dtype)
    indices = torch.tensor([s['index'] for s in batch])
    prompt_mask, _ = tokens_as_tensor([s['prompt_mask'] for s in batch], padding_idx, dtype)
    prompt_length = mask_to_len(prompt_mask)
    meta = batch[0]['meta']
    meta = {k: v for k, v in meta.items() if all(s['meta'].get(k) == v for s in batch[1:])}
    encoder_input, encoder_input_length = tokens_as_tensor([s['encoder_input'] for s in batch], padding_idx, dtype)
    return {
        'decoder_input': decoder_input, 'meta': meta, 'prompt_mask': prompt_mask, 'prompt_length': prompt_length,
        'indices': indices, 'encoder_input': encoder_input, 'encoder_input_length': encoder_input_length,
    }


def should_skip(source_len: int, target_len: int, max_source_len: int, max_target_len: int, min_len_ratio=None,
                max_len_ratio=None) -> bool:
    ratio = source_len / max(1, target_len)
    return bool(
        source_len == 0 or
        source_len > max_source_len or
        target_len > max_target_len or
        min_len_ratio and ratio < min_len_ratio or
        max_len_ratio and ratio > max_len_ratio
    )
