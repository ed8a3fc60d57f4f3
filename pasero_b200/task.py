"""Host-side feature/batch contract of the speech-translation task (integer work, bit-exact with the reference):
Dictionary ids, NumpyFile feature container, sample binarisation, length bucketing and collation.

Mirrors pasero/preprocessing.py:45-236 (Dictionary), pasero/files.py:103-192 (NumpyFile),
pasero/tasks/speech_translation.py:42-86 and pasero/tasks/translation.py:343-540 (preprocess, language-code / domain
tags), pasero/preprocessing.py:239-269 (tag helpers), pasero/tasks/task.py:531-610 (collate, build_batches),
pasero/utils.py:709-736, 783-843 (tokens_as_tensor, build_batches), pasero/datasets.py:708-730 (buffered_batching).
The numpy global RNG is consumed in the same order as the reference, so seeded runs produce identical batches.
"""
import collections
import itertools
import os
import pickle
import re
from typing import Callable, Iterable, Optional

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence

from .config import TrainingConfig
from .modules import mask_to_len

SEP, PAD, EOS, UNK, BOS = '<sep>', '<pad>', '</s>', '<unk>', '<s>'


class Dictionary:
    """fairseq-style dictionary file ("token count" per line) -> ids; specials sep=0, pad=1, eos=2, unk=3 are
    prepended when the file has none (pasero/preprocessing.py:125-150); bos falls back to eos."""

    def __init__(self, path: Optional[str] = None, tokens: Optional[list] = None):
        if tokens is None:
            tokens = []
            with open(path, encoding='utf-8') as f:
                for line in f:
                    line = line.rstrip('\n')
                    if line:
                        tokens.append(line.split(' ')[0])
        token_set = set(tokens)
        if EOS not in token_set:
            for t in (SEP, BOS, PAD, EOS, UNK):
                assert t not in token_set, 'dictionary has a partial set of special tokens'
            tokens = [SEP, PAD, EOS, UNK] + tokens
        self.tokens = tokens
        self.indices = {w: i for i, w in enumerate(tokens)}
        self.eos_idx = self.indices.get(EOS)
        self.padding_idx = self.indices.get(PAD, self.indices.get(UNK))
        self.bos_idx = self.indices.get(BOS, self.eos_idx)
        self.unk_idx = self.indices.get(UNK, self.padding_idx)
        self.sep_idx = self.indices.get(SEP, self.bos_idx)

    def __len__(self):
        return len(self.tokens)

    def idx(self, token: str) -> int:
        return self.indices.get(token, self.unk_idx)

    def to_indices(self, tokens: list, max_len: Optional[int] = None, append_eos: bool = True,
                   prepend_bos: bool = False, truncate_left: bool = False) -> np.ndarray:
        ids = [self.idx(t) for t in tokens]
        if max_len is not None:
            max_len = max_len - int(append_eos) - int(prepend_bos)
            ids = ids[-max_len:] if truncate_left else ids[:max_len]
        if prepend_bos:
            ids.insert(0, self.bos_idx)
        if append_eos:
            ids.append(self.eos_idx)
        return np.array(ids, dtype=np.int64)


# ---- language-code / domain tags (pasero/preprocessing.py:239-269) ----
_LANG_CODE_REGEX = re.compile(r'<lang:(.+?)>')
_DOMAIN_TAG_REGEX = re.compile(r'<domain:(.+?)>')


def get_lang_code(lang: Optional[str]) -> Optional[str]:
    return f'<lang:{lang}>' if lang else None


def get_domain_tag(domain: Optional[str]) -> Optional[str]:
    return f'<domain:{domain}>' if domain else None


def is_tag(token: str) -> bool:
    return bool(_LANG_CODE_REGEX.fullmatch(token) or _DOMAIN_TAG_REGEX.fullmatch(token))


def split_tags(line: str) -> list:
    """[tag, ..., rest]: the leading tags of a line that already carries some, and the line without them."""
    tags = list(itertools.takewhile(is_tag, line.split()))
    for tag in tags:   # strip this tag and the blanks around it only (the line may span several lines)
        line = re.sub(rf'\s*{re.escape(tag)}\s*', '', line, count=1)
    return [*tags, line]


class NumpyFile:
    """Reader / writer for the reference's feature container (pasero/files.py:103-192): a pickled header
    {'positions', 'lengths', 'dim', 'dtype'} followed by the raw rows of every utterance; header slots whose position is
    0 are unused and skipped (files.py:112-115). Files are interchangeable with the reference's (tests/golden/
    features_ref.npy was written by it).

    B200-side design: the file is memory-mapped once, `f[i]` is a zero-copy view, and `pack()` lays a whole batch out
    in one (pinned) staging buffer — ragged rows back to back plus offsets — which is what the device-side collate
    (`pasero_b200.frontend.collate_features`) consumes: one H2D copy of the true bytes instead of a padded fp32 tensor."""

    def __init__(self, path: str):
        self._path = path
        with open(path, 'rb') as file:
            header = pickle.load(file)
        self._dim = int(header['dim'])
        self._dtype = np.dtype(header['dtype'])
        used = np.asarray(header['positions'], dtype=np.int64) > 0
        self._positions = np.asarray(header['positions'], dtype=np.int64)[used]
        self._lengths = np.asarray(header['lengths'], dtype=np.int64)[used]
        self._map = np.memmap(path, dtype=np.uint8, mode='r') if len(self._positions) else np.zeros(0, dtype=np.uint8)
        self._index = 0

    # ---- writing ---------------------------------------------------------------------------------
    @staticmethod
    def _header(positions, lengths, dim: int, dtype: str) -> bytes:
        return pickle.dumps({'positions': positions, 'lengths': lengths, 'dim': np.array(dim, dtype=np.int64), 'dtype': dtype})

    @classmethod
    def build(cls, path: str, features: Iterable[np.ndarray], dtype: str = 'float16', num_feats: Optional[int] = None):
        """Byte-compatible with the reference's writer: the header is pickled with `num_feats` slots (may exceed the number
        of utterances actually written), rows follow in order."""
        slots = len(features) if num_feats is None else num_feats
        positions, lengths = np.zeros(slots, dtype=np.int64), np.zeros(slots, dtype=np.int64)
        if os.path.dirname(path):
            os.makedirs(os.path.dirname(path), exist_ok=True)
        dim = 0
        with open(path, 'wb') as file:
            cursor = file.write(cls._header(positions, lengths, dim, dtype))    # fixed size: int64 arrays of `slots` entries
            for i, rows in enumerate(features):
                rows = np.ascontiguousarray(rows, dtype=dtype)
                positions[i], lengths[i], dim = cursor, rows.shape[0], rows.shape[1]
                cursor += file.write(rows.tobytes())
            file.seek(0)
            file.write(cls._header(positions, lengths, dim, dtype))
        return cls(path)

    # ---- reading ---------------------------------------------------------------------------------
    def __len__(self):
        return len(self._positions)

    @property
    def dim(self) -> int:
        return self._dim

    @property
    def dtype(self) -> np.dtype:
        return self._dtype

    def get_positions(self):
        return np.arange(len(self._positions)), self._lengths

    def __getitem__(self, index: int) -> np.ndarray:
        """Zero-copy (read-only) view of utterance `index`: (frames, dim), or (frames,) when dim == 1."""
        n = int(self._lengths[index])
        begin = int(self._positions[index])
        x = self._map[begin:begin + n * self._dim * self._dtype.itemsize].view(self._dtype)
        return x.reshape(n, self._dim) if self._dim > 1 else x

    def seek(self, index: int):
        self._index = index

    def tell(self) -> int:
        return self._index

    def __iter__(self):
        return self

    def __next__(self) -> np.ndarray:
        """The reference's iteration protocol (files.py:164-175): an owned copy of the next utterance."""
        if self._index >= len(self._positions):
            raise StopIteration
        x = np.array(self[self._index])
        self._index += 1
        return x

    def pack(self, indices, out: Optional[np.ndarray] = None):
        """Utterances `indices` back to back in one buffer: (rows [sum frames, dim] in the file's dtype, offsets int64
        [len + 1] in frames). `out` (e.g. the numpy view of a pinned tensor, at least sum-frames rows) is filled in place."""
        indices = [int(i) for i in indices]
        offsets = np.zeros(len(indices) + 1, dtype=np.int64)
        np.cumsum(self._lengths[indices], out=offsets[1:])
        total = int(offsets[-1])
        if out is None:
            out = np.empty((total, self._dim), dtype=self._dtype)
        assert out.dtype == self._dtype and out.shape[0] >= total and out.shape[1] == self._dim
        for j, i in enumerate(indices):
            out[offsets[j]:offsets[j + 1]] = self[i].reshape(-1, self._dim)
        return out[:total], offsets

    def close(self):
        self._map = np.zeros(0, dtype=np.uint8)


def tokens_as_tensor(token_list: list, padding_idx: int, dtype: Optional[torch.dtype] = None):
    """pasero/utils.py:709-736: right-pad float -> 0.0 (cast to dtype), bool -> False, int -> padding_idx (int64)."""
    token_list = [torch.as_tensor(tokens) for tokens in token_list]
    if token_list[0].is_floating_point():
        tokens = pad_sequence(token_list, batch_first=True, padding_value=0.0).to(dtype)
    elif token_list[0].dtype is torch.bool:
        tokens = pad_sequence(token_list, batch_first=True, padding_value=False).bool()
    else:
        tokens = pad_sequence(token_list, batch_first=True, padding_value=padding_idx).long()
    lengths = torch.LongTensor(list(map(len, token_list)))
    return tokens, lengths


def build_batches(indices, size_fn: Callable, batch_size: int, batch_size_multiple: Optional[int] = None,
                  max_lines: Optional[int] = None) -> list:
    """pasero/utils.py:783-843 — greedy packing of length-sorted indices under a padded-size budget."""
    batch_size_multiple = batch_size_multiple or 1
    batches, batch, lengths = [], [], []
    max_len = 0
    for idx in indices:
        size = size_fn(idx)
        if size > batch_size:
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


class SpeechTranslationTask:
    """The slice of pasero.tasks.SpeechTranslationTask the training step needs."""

    model_type = 'encoder_decoder'
    find_unused_parameters = False
    encoder_num_embeddings = 0

    def __init__(self, cfg: TrainingConfig, dictionary: Dictionary, target_tags: tuple = ()):
        self.cfg = cfg
        self.dictionary = dictionary
        self.target_tags = tuple(target_tags)   # always prepended (pre-tokenised pipelines that carry their own tags)
        # tag options of TranslationTaskConfig (pasero/config.py:1940-1960), read from the flat config
        opt = cfg.extra.get
        self.lang_code, self.source_lang_code = bool(opt('lang_code')), bool(opt('source_lang_code'))
        self.target_lang_code, self.domain_tag = bool(opt('target_lang_code')), bool(opt('domain_tag'))
        self.cfg_source_tags, self.cfg_target_tags = list(opt('source_tags') or ()), list(opt('target_tags') or ())
        self.prepend_bos = True
        self.max_source_len = cfg.max_source_len
        self.max_target_len = cfg.max_target_len
        self.min_len_ratio = self.max_len_ratio = None

    @property
    def padding_idx(self) -> int:
        return self.dictionary.padding_idx

    @property
    def eos_idx(self) -> int:
        return self.dictionary.eos_idx

    @property
    def bos_idx(self) -> int:
        return self.dictionary.bos_idx

    @property
    def decoder_num_embeddings(self) -> int:
        return len(self.dictionary)

    @classmethod
    def text(cls, cfg: TrainingConfig, dictionary: Dictionary, target_tags: tuple = ()) -> 'SpeechTranslationTask':
        """The text-input flavour (`--task translation`, pasero/tasks/translation.py): same batch contract, but
        `encoder_input` holds token ids of the (shared) dictionary, so the encoder needs an embedding table."""
        task = cls(cfg, dictionary, target_tags)
        task.encoder_num_embeddings = len(dictionary)
        return task

    def get_langs_or_domains(self, key: str) -> set:
        """pasero/tasks/translation.py (TranslationTask.get_langs_or_domains): the values `key` can take in this task's
        corpora, read from the flat config (`source_langs` / `target_langs` / `domains` lists or the single
        `source_lang` / `target_lang` / `domain`)."""
        extra = self.cfg.extra

        def values(name):
            v = extra.get(name + 's') or ([extra[name]] if extra.get(name) else [])
            return set(v)
        if key == 'lang':
            return values('source_lang') | values('target_lang')
        return values(key)

    def get_source_tags(self, meta: dict) -> list:
        """Tags prepended to SOURCE lines (pasero/tasks/translation.py:369-388): configured tags, the TARGET language code
        when `lang_code`, the source language code, the domain tag."""
        tags = list(meta.get('source_tags') or self.cfg_source_tags)
        if self.lang_code:
            assert meta.get('target_lang'), 'missing target language information'
            tags.append(get_lang_code(meta.get('target_lang')))
        if self.source_lang_code:
            assert meta.get('source_lang'), 'missing source language information'
            tags.append(get_lang_code(meta.get('source_lang')))
        if self.domain_tag:
            tags.append(get_domain_tag(meta.get('domain')))
        return [t for t in tags if t]

    def get_target_tags(self, meta: dict) -> list:
        """Tags prepended to TARGET lines and excluded from the loss through the prompt mask (translation.py:390-403)."""
        tags = list(meta.get('target_tags') or self.cfg_target_tags)
        if self.target_lang_code:
            assert meta.get('target_lang'), 'missing target language information'
            tags.append(get_lang_code(meta.get('target_lang')))
        return [t for t in tags if t]

    def check_tags(self) -> None:
        """Every tag this configuration can emit is in the dictionary (translation.py:343-367)."""
        need = []
        if self.domain_tag:
            need += [get_domain_tag(d) for d in self.get_langs_or_domains('domain')]
        if self.source_lang_code:
            need += [get_lang_code(lang) for lang in self.get_langs_or_domains('source_lang')]
        if self.lang_code or self.target_lang_code:
            need += [get_lang_code(lang) for lang in self.get_langs_or_domains('target_lang')]
        for tag in need:
            assert tag in self.dictionary.indices, f'{tag} is OOV'

    def should_skip(self, source_len: int, target_len: int) -> bool:
        ratio = source_len / max(1, target_len)
        return bool(source_len == 0 or source_len > self.max_source_len or target_len > self.max_target_len or
                    self.min_len_ratio and ratio < self.min_len_ratio or self.max_len_ratio and ratio > self.max_len_ratio)

    def preprocess(self, sample: dict, truncate: bool = False, tokenize: bool = False, append_eos: bool = False) -> dict:
        """pasero/tasks/speech_translation.py:42-86 with pre-tokenised targets (BPE itself is CPU string work that
        stays with the reference's tokenizers): features pass through, target -> ids with BOS/EOS, prompt mask."""
        source, target, meta = sample['source'], sample['target'], sample.get('meta') or {}
        source_cutoff = self.max_source_len if truncate else None
        target_cutoff = self.max_target_len if truncate else None
        text_source = isinstance(source, str)     # `--task translation`: the source is a line of text too
        # `tokenize` keeps the reference's meaning for the tags — they are added only then (translation.py:450-452; the
        # speech task always adds the target tags, speech_translation.py:56-58) — but the subword split itself is CPU
        # string work that stays with the reference's tokenizers: lines arrive BPE-split and are split on blanks here.
        tags = list(self.target_tags)
        src_tags = []
        if text_source:
            *src_tags, source = split_tags(source)
            if target is not None:
                *pre, target = split_tags(target)
                tags += pre
            if tokenize:
                src_tags += self.get_source_tags(meta)
                tags += self.get_target_tags(meta)
            # (without `tokenize` the reference drops even the tags the source line carried: translation.py:461)
            source_tok = src_tags + source.split() if tokenize else source.split()
            encoder_input = self.dictionary.to_indices(source_tok, max_len=source_cutoff, prepend_bos=False, append_eos=True)
        else:
            tags += self.get_target_tags(meta)
            encoder_input = source[:source_cutoff]
        target_tok = tags + (target.split() if target else [])
        decoder_input = self.dictionary.to_indices(target_tok, max_len=target_cutoff, prepend_bos=self.prepend_bos,
                                                   append_eos=append_eos)
        prompt_mask = np.zeros_like(decoder_input, dtype=bool)
        prompt_mask[:len(tags) + int(self.prepend_bos)] = True
        if self.should_skip(len(encoder_input), len(decoder_input)):
            assert not truncate
            return {}
        return {'encoder_input': encoder_input, 'decoder_input': decoder_input, 'prompt_mask': prompt_mask, 'meta': meta}

    def build_batches(self, data: list, shuffle: bool = True, sort: bool = True, batch_size: Optional[int] = None) -> list:
        """pasero/tasks/task.py:573-610"""
        encoder_length = np.array([len(sample['encoder_input']) for sample in data])
        decoder_length = np.array([len(sample['decoder_input']) for sample in data])
        length = np.maximum(encoder_length, decoder_length)
        indices = np.random.permutation(len(data)) if shuffle else np.arange(len(data))
        if sort:
            indices = indices[np.argsort(decoder_length[indices], kind='stable')]
            indices = indices[np.argsort(encoder_length[indices], kind='stable')]
        batch_indices = build_batches(indices, length.__getitem__, batch_size or self.cfg.batch_size,
                                      self.cfg.batch_size_multiple, self.cfg.lines_per_batch)
        return [[{**data[i], 'index': i} for i in idx] for idx in batch_indices]

    def buffered_batching(self, buffer: list, batch_by=(), shuffle: bool = True, batch_size: Optional[int] = None) -> list:
        """pasero/datasets.py:708-730"""
        groups = collections.defaultdict(list)
        for sample in buffer:
            key = tuple(sample['meta'].get(k) for k in batch_by)
            if sample.get('encoder_input') is not None:
                key = key + (sample['encoder_input'].ndim, sample['encoder_input'].dtype)
            groups[key].append(sample)
        batches = []
        for group in groups.values():
            batches += self.build_batches(group, shuffle=shuffle, batch_size=batch_size)
        if shuffle:
            np.random.shuffle(batches)
        return batches

    @classmethod
    def collate(cls, batch: list, padding_idx: int, dtype: torch.dtype, model_type: str = 'encoder_decoder'):
        """pasero/tasks/task.py:531-571 + pasero/tasks/translation.py:876-899"""
        if not batch:
            return None
        decoder_input, _ = tokens_as_tensor([s['decoder_input'] for s in batch], padding_idx, dtype)
        indices = torch.tensor([s['index'] for s in batch])
        prompt_mask, _ = tokens_as_tensor([s['prompt_mask'] for s in batch], padding_idx, dtype)
        prompt_length = mask_to_len(prompt_mask)
        meta = batch[0]['meta']
        meta = {k: v for k, v in meta.items() if all(s['meta'].get(k) == v for s in batch[1:])}
        batched = {'decoder_input': decoder_input, 'meta': meta, 'prompt_mask': prompt_mask, 'prompt_length': prompt_length,
                   'indices': indices}
        encoder_input, encoder_input_length = tokens_as_tensor([s['encoder_input'] for s in batch], padding_idx, dtype)
        batched.update({'encoder_input': encoder_input, 'encoder_input_length': encoder_input_length})
        return batched
