"""Seeded synthetic inputs shared by tests/golden/make_golden.py (which feeds them to the reference) and the
tests (which feed them to the oracle and to the CUDA path). numpy.random.RandomState only: reproducible
everywhere."""
import math

import numpy as np
import torch

TINY = dict(vocab_size=200, encoder_layers=2, decoder_layers=2, embed_dim=128, ffn_dim=256, heads=2,
            conv_channels=256)
TRAIN = dict(lr=2e-4, warmup=50, steps=200, clip_norm=1.0, label_smoothing=0.1)
FBANK_CASES = {'noise': 8000, 'tone': 12000, 'single_frame': 400}
BATCHING_SEED = 7
BY_LANGS = {'encoder': ('de', 'it'), 'decoder': ('en', 'fr')}      # per-language adapter fixtures, see below
PAD, EOS, BOS = 1, 2, 2


def waveform(name: str, n: int) -> np.ndarray:
    """float32 PCM in [-1, 1)."""
    rng = np.random.RandomState(1234)
    if name == 'tone':
        t = np.arange(n) / 16000.0
        w = 0.3 * np.sin(2 * math.pi * 440 * t) * np.exp(-t) + 0.01 * rng.randn(n)
    else:
        w = rng.rand(n) - 0.5
    return w.astype(np.float32)


def batching_samples(n: int = 5000, real_features: int = 64) -> list:
    """SURVEY.md §8d batching parity set: F ~ clip(lognormal(6.3, 0.6), 100, 3000), T+1 ~ clip(round(F/20)+U{0..8}, 3, 256)."""
    rng = np.random.RandomState(BATCHING_SEED)
    F = np.clip(rng.lognormal(6.3, 0.6, n), 100, 3000).astype(np.int64)
    T1 = np.clip(np.round(F / 20).astype(np.int64) + rng.randint(0, 9, n), 3, 256)
    zero = np.zeros((1, 80), dtype=np.float16)
    samples = []
    for i in range(n):
        if i < real_features:
            feats = rng.randn(int(F[i]), 80).astype(np.float16)
        else:
            feats = np.broadcast_to(zero, (int(F[i]), 80))  # only len/ndim/dtype matter for bucketing
        dec = np.concatenate([[BOS], rng.randint(4, TINY['vocab_size'], int(T1[i]) - 2), [EOS]]).astype(np.int64)
        pm = np.zeros_like(dec, dtype=bool)
        pm[:1] = True
        samples.append({'encoder_input': feats, 'decoder_input': dec, 'prompt_mask': pm,
                        'meta': {'source_lang': 'de', 'target_lang': 'en'}})
    return samples


def batching_samples_by_lang(n: int = 1200) -> list:
    """The first n samples of the batching parity set, spread over 2 source x 2 target languages (a fixed function of the
    sample number) and tagged with a 'uid' (the `index` field the batcher writes is local to a language group)."""
    out = []
    for i, s in enumerate(batching_samples(n, real_features=0)):
        src, tgt = BY_LANGS['encoder'][(i // 3) % 2], BY_LANGS['decoder'][(i // 7) % 2]
        out.append({**s, 'meta': {'source_lang': src, 'target_lang': tgt, 'uid': i}})
    return out


def _batch(rng, frames, dec_lens) -> dict:
    B, Fm, Tm = len(frames), max(frames), max(dec_lens)
    x = torch.zeros(B, Fm, 80)
    dec = torch.full((B, Tm), PAD, dtype=torch.int64)
    pm = torch.zeros(B, Tm, dtype=torch.bool)
    for b, (f, t) in enumerate(zip(frames, dec_lens)):
        x[b, :f] = torch.from_numpy(rng.randn(f, 80).astype(np.float32)).to(torch.bfloat16).float()
        dec[b, :t] = torch.from_numpy(np.concatenate([[BOS], rng.randint(4, TINY['vocab_size'], t - 2), [EOS]]))
        pm[b, 0] = True
    return {'encoder_input': x, 'encoder_input_length': torch.tensor(frames), 'decoder_input': dec, 'prompt_mask': pm,
            'prompt_length': torch.ones(B, dtype=torch.int64), 'indices': torch.arange(B),
            'meta': {'source_lang': 'de', 'target_lang': 'en'}}


def model_batch() -> dict:
    return _batch(np.random.RandomState(11), [93, 61, 40, 77], [9, 12, 5, 7])


def train_batches(n: int = 16) -> list:
    """Zipf-distributed target tokens: learnable (unigram statistics) but not memorisable within 200 steps, so the
    loss curve is smooth and two correct implementations stay within a fraction of a percent of each other."""
    rng = np.random.RandomState(12)
    probs = 1.0 / np.arange(1, TINY['vocab_size'] - 3)
    probs /= probs.sum()
    out = []
    for _ in range(n):
        B = int(rng.randint(3, 7))
        frames = sorted((int(f) for f in rng.randint(40, 200, B)), reverse=True)
        dec_lens = [int(t) for t in rng.randint(5, 16, B)]
        b = _batch(rng, frames, dec_lens)
        for i, t in enumerate(dec_lens):
            b['decoder_input'][i, 1:t - 1] = torch.from_numpy(4 + rng.choice(len(probs), t - 2, p=probs))
        out.append(b)
    return out


def tiny_param_shapes() -> dict:
    """Parameter names/shapes of the tiny ST model in the reference's state_dict naming."""
    t = TINY
    d, f, c, v = t['embed_dim'], t['ffn_dim'], t['conv_channels'], t['vocab_size']
    shapes = {
        'encoder.subsample.conv_layers.0.weight': (c, 80, 5), 'encoder.subsample.conv_layers.0.bias': (c,),
        'encoder.subsample.conv_layers.1.weight': (2 * d, c // 2, 5), 'encoder.subsample.conv_layers.1.bias': (2 * d,),
    }

    def attn(prefix):
        for p in ('k_proj', 'v_proj', 'q_proj', 'out_proj'):
            shapes[f'{prefix}.{p}.weight'] = (d, d)
            shapes[f'{prefix}.{p}.bias'] = (d,)

    def ln(prefix):
        shapes[f'{prefix}.weight'] = (d,)
        shapes[f'{prefix}.bias'] = (d,)

    def ffn(prefix):
        shapes[f'{prefix}.fc1.weight'] = (f, d)
        shapes[f'{prefix}.fc1.bias'] = (f,)
        shapes[f'{prefix}.fc2.weight'] = (d, f)
        shapes[f'{prefix}.fc2.bias'] = (d,)

    for i in range(t['encoder_layers']):
        p = f'encoder.layers.{i}'
        attn(p + '.self_attn'); ln(p + '.self_attn_layer_norm'); ffn(p); ln(p + '.final_layer_norm')
    shapes['decoder.embed_tokens.weight'] = (v, d)
    for i in range(t['decoder_layers']):
        p = f'decoder.layers.{i}'
        attn(p + '.self_attn'); ln(p + '.self_attn_layer_norm')
        attn(p + '.encoder_attn'); ln(p + '.encoder_attn_layer_norm'); ffn(p); ln(p + '.final_layer_norm')
    return shapes


def tiny_state_dict() -> dict:
    """Seeded, reference-independent weights (bf16-representable fp32) for the tiny model."""
    rng = np.random.RandomState(99)
    sd = {}
    for name, shape in tiny_param_shapes().items():
        if 'layer_norm' in name:
            w = (1.0 + 0.1 * rng.randn(*shape)) if name.endswith('weight') else 0.05 * rng.randn(*shape)
        elif name.endswith('bias'):
            w = 0.02 * rng.randn(*shape)
        elif name == 'decoder.embed_tokens.weight':
            w = rng.randn(*shape) * shape[1] ** -0.5
            w[PAD] = 0
        else:
            fan_in = int(np.prod(shape[1:]))
            w = rng.randn(*shape) * fan_in ** -0.5
        sd[name] = torch.from_numpy(w.astype(np.float32)).to(torch.bfloat16).float()
    return sd


# ---- Whisper-style variant (BASELINE config 4 shape family, pasero/config.py:2539-2566): pre-norm, GELU, learned positions,
# conv k=3,3 strides 1,2 with GELU, no embedding scale, no key bias, final encoder / decoder LayerNorm
WSP = dict(vocab_size=200, encoder_layers=2, decoder_layers=2, embed_dim=128, ffn_dim=256, heads=2, conv_channels=128,
           encoder_max_len=400, decoder_max_len=64)


def wsp_param_shapes() -> dict:
    t = WSP
    d, f, c, v = t['embed_dim'], t['ffn_dim'], t['conv_channels'], t['vocab_size']
    shapes = {
        'encoder.subsample.conv_layers.0.weight': (c, 80, 3), 'encoder.subsample.conv_layers.0.bias': (c,),
        'encoder.subsample.conv_layers.1.weight': (d, c, 3), 'encoder.subsample.conv_layers.1.bias': (d,),
        'encoder.embed_positions.weight': (t['encoder_max_len'] // 2, d),
    }

    def attn(prefix):
        for p in ('k_proj', 'v_proj', 'q_proj', 'out_proj'):
            shapes[f'{prefix}.{p}.weight'] = (d, d)
            if p != 'k_proj':            # attention_key_bias=False (modules.py:516-522)
                shapes[f'{prefix}.{p}.bias'] = (d,)

    def ln(prefix):
        shapes[f'{prefix}.weight'] = (d,)
        shapes[f'{prefix}.bias'] = (d,)

    def ffn(prefix):
        shapes[f'{prefix}.fc1.weight'] = (f, d)
        shapes[f'{prefix}.fc1.bias'] = (f,)
        shapes[f'{prefix}.fc2.weight'] = (d, f)
        shapes[f'{prefix}.fc2.bias'] = (d,)

    for i in range(t['encoder_layers']):
        p = f'encoder.layers.{i}'
        attn(p + '.self_attn'); ln(p + '.self_attn_layer_norm'); ffn(p); ln(p + '.final_layer_norm')
    ln('encoder.layer_norm')
    shapes['decoder.embed_tokens.weight'] = (v, d)
    shapes['decoder.embed_positions.weight'] = (t['decoder_max_len'], d)
    for i in range(t['decoder_layers']):
        p = f'decoder.layers.{i}'
        attn(p + '.self_attn'); ln(p + '.self_attn_layer_norm')
        attn(p + '.encoder_attn'); ln(p + '.encoder_attn_layer_norm'); ffn(p); ln(p + '.final_layer_norm')
    ln('decoder.layer_norm')
    return shapes


def wsp_state_dict() -> dict:
    rng = np.random.RandomState(77)
    sd = {}
    for name, shape in wsp_param_shapes().items():
        if 'layer_norm' in name:
            w = (1.0 + 0.1 * rng.randn(*shape)) if name.endswith('weight') else 0.05 * rng.randn(*shape)
        elif name.endswith('bias'):
            w = 0.02 * rng.randn(*shape)
        elif name.endswith('embed_positions.weight'):
            w = 0.05 * rng.randn(*shape)
        elif name == 'decoder.embed_tokens.weight':
            w = rng.randn(*shape) * shape[1] ** -0.5
            w[PAD] = 0
        else:
            fan_in = int(np.prod(shape[1:]))
            w = rng.randn(*shape) * fan_in ** -0.5
        sd[name] = torch.from_numpy(w.astype(np.float32)).to(torch.bfloat16).float()
    return sd


# ---- adapter variant (pasero/models/adapters.py, `--arch adapter_transformer`): the tiny ST model + a bottleneck adapter
# after every layer; only the adapters are trained (base parameters frozen)
ADAPTER_DIM = 16


def adapter_param_shapes() -> dict:
    t = TINY
    d = t['embed_dim']
    shapes = dict(tiny_param_shapes())
    for side, n in (('encoder', t['encoder_layers']), ('decoder', t['decoder_layers'])):
        for i in range(n):
            p = f'{side}.layers.{i}.adapters.default'
            shapes[p + '.down.weight'] = (ADAPTER_DIM, d)
            shapes[p + '.down.bias'] = (ADAPTER_DIM,)
            shapes[p + '.up.weight'] = (d, ADAPTER_DIM)
            shapes[p + '.up.bias'] = (d,)
            shapes[p + '.layer_norm.weight'] = (d,)
            shapes[p + '.layer_norm.bias'] = (d,)
    return shapes


def adapter_state_dict() -> dict:
    sd = dict(tiny_state_dict())
    rng = np.random.RandomState(55)
    for name, shape in adapter_param_shapes().items():
        if name in sd:
            continue
        if 'layer_norm' in name:
            w = (1.0 + 0.1 * rng.randn(*shape)) if name.endswith('weight') else 0.05 * rng.randn(*shape)
        elif name.endswith('bias'):
            w = 0.02 * rng.randn(*shape)
        else:
            w = rng.randn(*shape) * shape[1] ** -0.5
        sd[name] = torch.from_numpy(w.astype(np.float32)).to(torch.bfloat16).float()
    return sd


# ---- per-language adapters (`--encoder-adapters-by source_lang --decoder-adapters-by target_lang`, adapters.py:74-105,
# 186-229): one adapter set per source language in the encoder and per target language in the decoder; a batch only uses (and
# only trains) the sets of its own language pair, so every step leaves some parameters without a gradient
# language pair of each training step (cycled): 'fr' first appears at step 3, 'it' at step 2, and (it, fr) only once per cycle,
# so the adapter sets' Adam step counts drift apart
BY_SCHEDULE = [('de', 'en'), ('de', 'en'), ('it', 'en'), ('de', 'fr'), ('de', 'en'), ('it', 'fr'), ('de', 'en'), ('it', 'en')]
BY_TRAIN = dict(lr=1e-3, warmup=10, steps=40)


def by_adapter_param_shapes() -> dict:
    t = TINY
    d = t['embed_dim']
    shapes = dict(tiny_param_shapes())
    for side, n in (('encoder', t['encoder_layers']), ('decoder', t['decoder_layers'])):
        for i in range(n):
            for lang in BY_LANGS[side]:
                p = f'{side}.layers.{i}.adapters.lang:{lang}'
                shapes[p + '.down.weight'] = (ADAPTER_DIM, d)
                shapes[p + '.down.bias'] = (ADAPTER_DIM,)
                shapes[p + '.up.weight'] = (d, ADAPTER_DIM)
                shapes[p + '.up.bias'] = (d,)
                shapes[p + '.layer_norm.weight'] = (d,)
                shapes[p + '.layer_norm.bias'] = (d,)
    return shapes


def by_adapter_state_dict() -> dict:
    sd = dict(tiny_state_dict())
    rng = np.random.RandomState(56)
    for name, shape in by_adapter_param_shapes().items():
        if name in sd:
            continue
        if 'layer_norm' in name:
            w = (1.0 + 0.1 * rng.randn(*shape)) if name.endswith('weight') else 0.05 * rng.randn(*shape)
        elif name.endswith('bias'):
            w = 0.02 * rng.randn(*shape)
        else:
            w = rng.randn(*shape) * shape[1] ** -0.5
        sd[name] = torch.from_numpy(w.astype(np.float32)).to(torch.bfloat16).float()
    return sd


_TRAIN_BATCHES = None


def by_train_batch(step: int) -> dict:
    """The batch of training step `step`: the seeded batches of train_batches(), tagged with the step's language pair."""
    global _TRAIN_BATCHES
    if _TRAIN_BATCHES is None:
        _TRAIN_BATCHES = train_batches()
    batches = _TRAIN_BATCHES
    b = dict(batches[step % len(batches)])
    src, tgt = BY_SCHEDULE[step % len(BY_SCHEDULE)]
    b['meta'] = {'source_lang': src, 'target_lang': tgt}
    return b


# ---- text-input variant (BASELINE configs 1 / 5 family: `--task translation`, token-id encoder sharing its embedding
# table with the decoder and the output projection; pasero/models/transformer.py:610-752)
def text_param_shapes() -> dict:
    shapes = {k: v for k, v in tiny_param_shapes().items() if not k.startswith('encoder.subsample')}
    shapes['encoder.embed_tokens.weight'] = shapes['decoder.embed_tokens.weight']
    return shapes


def text_state_dict() -> dict:
    sd = {k: v for k, v in tiny_state_dict().items() if not k.startswith('encoder.subsample')}
    sd['encoder.embed_tokens.weight'] = sd['decoder.embed_tokens.weight']     # shared_embeddings (config.py:1064-1067)
    return sd


def text_batch() -> dict:
    rng = np.random.RandomState(21)
    src_lens, dec_lens = [17, 9, 12, 5], [9, 12, 5, 7]
    B, Sm, Tm = len(src_lens), max(src_lens), max(dec_lens)
    src = torch.full((B, Sm), PAD, dtype=torch.int64)
    dec = torch.full((B, Tm), PAD, dtype=torch.int64)
    pm = torch.zeros(B, Tm, dtype=torch.bool)
    for b, (s_, t) in enumerate(zip(src_lens, dec_lens)):
        src[b, :s_] = torch.from_numpy(np.concatenate([rng.randint(4, TINY['vocab_size'], s_ - 1), [EOS]]))
        dec[b, :t] = torch.from_numpy(np.concatenate([[BOS], rng.randint(4, TINY['vocab_size'], t - 2), [EOS]]))
        pm[b, 0] = True
    return {'encoder_input': src, 'encoder_input_length': torch.tensor(src_lens), 'decoder_input': dec, 'prompt_mask': pm,
            'prompt_length': torch.ones(B, dtype=torch.int64), 'indices': torch.arange(B),
            'meta': {'source_lang': 'de', 'target_lang': 'en'}}


def summarize_grad(g: torch.Tensor) -> dict:
    """Compact but discriminating digest of a gradient tensor (full tensor when small)."""
    g = g.detach().float()
    if g.numel() <= 4096:
        return {'full': g.clone()}
    g2 = g.reshape(g.shape[0], -1)
    return {'rows': g2[:4].clone(), 'row_sum': g2.sum(1), 'col_sum': g2.sum(0), 'abs_sum': g2.abs().sum()}


# ---- BASELINE-sized variants (tests/test_real_shapes_gpu.py): the same parameter layouts at the true dimensions of
# BASELINE configs 2 / 4 / 5. No reference-generated fixture is committed for these (hundreds of MB): the CUDA path is
# compared with the oracle (itself pinned to the reference on the tiny fixtures) on weights and batches seeded here.
REAL_ST = dict(vocab_size=8032, encoder_layers=12, decoder_layers=6, embed_dim=512, ffn_dim=2048, heads=8, conv_channels=1024)
REAL_WSP = dict(vocab_size=51865, encoder_layers=12, decoder_layers=12, embed_dim=768, ffn_dim=3072, heads=12,
                conv_channels=768, encoder_max_len=3000, decoder_max_len=448)
REAL_BIG = dict(vocab_size=70376, encoder_layers=6, decoder_layers=6, embed_dim=1024, ffn_dim=4096, heads=16, adapter_dim=64)


def sized_param_shapes(t: dict, family: str = 'st') -> dict:
    """Reference-named parameter shapes for dimension dict `t`; family: 'st' (conv k=5,5 + GLU, sinusoidal, post-norm),
    'wsp' (Whisper family) or 'text' (shared token embedding, optional adapters when t has 'adapter_dim')."""
    d, f, v = t['embed_dim'], t['ffn_dim'], t['vocab_size']
    c = t.get('conv_channels')
    shapes = {}
    if family == 'st':
        shapes.update({'encoder.subsample.conv_layers.0.weight': (c, 80, 5), 'encoder.subsample.conv_layers.0.bias': (c,),
                       'encoder.subsample.conv_layers.1.weight': (2 * d, c // 2, 5), 'encoder.subsample.conv_layers.1.bias': (2 * d,)})
    elif family == 'wsp':
        shapes.update({'encoder.subsample.conv_layers.0.weight': (c, 80, 3), 'encoder.subsample.conv_layers.0.bias': (c,),
                       'encoder.subsample.conv_layers.1.weight': (d, c, 3), 'encoder.subsample.conv_layers.1.bias': (d,),
                       'encoder.embed_positions.weight': (t['encoder_max_len'] // 2, d)})
    key_bias = family != 'wsp'

    def attn(prefix):
        for p in ('k_proj', 'v_proj', 'q_proj', 'out_proj'):
            shapes[f'{prefix}.{p}.weight'] = (d, d)
            if p != 'k_proj' or key_bias:
                shapes[f'{prefix}.{p}.bias'] = (d,)

    def ln(prefix):
        shapes[f'{prefix}.weight'] = (d,)
        shapes[f'{prefix}.bias'] = (d,)

    def ffn(prefix):
        shapes[f'{prefix}.fc1.weight'] = (f, d)
        shapes[f'{prefix}.fc1.bias'] = (f,)
        shapes[f'{prefix}.fc2.weight'] = (d, f)
        shapes[f'{prefix}.fc2.bias'] = (d,)

    def adapter(prefix):
        a = t.get('adapter_dim')
        if a:
            p = prefix + '.adapters.default'
            shapes[p + '.down.weight'] = (a, d)
            shapes[p + '.down.bias'] = (a,)
            shapes[p + '.up.weight'] = (d, a)
            shapes[p + '.up.bias'] = (d,)
            ln(p + '.layer_norm')

    for i in range(t['encoder_layers']):
        p = f'encoder.layers.{i}'
        attn(p + '.self_attn'); ln(p + '.self_attn_layer_norm'); ffn(p); ln(p + '.final_layer_norm'); adapter(p)
    if family == 'wsp':
        ln('encoder.layer_norm')
    shapes['decoder.embed_tokens.weight'] = (v, d)
    if family == 'wsp':
        shapes['decoder.embed_positions.weight'] = (t['decoder_max_len'], d)
    for i in range(t['decoder_layers']):
        p = f'decoder.layers.{i}'
        attn(p + '.self_attn'); ln(p + '.self_attn_layer_norm')
        attn(p + '.encoder_attn'); ln(p + '.encoder_attn_layer_norm'); ffn(p); ln(p + '.final_layer_norm'); adapter(p)
    if family == 'wsp':
        ln('decoder.layer_norm')
    return shapes


def sized_state_dict(t: dict, family: str = 'st', seed: int = 3) -> dict:
    """Seeded bf16-representable fp32 weights for `sized_param_shapes(t, family)` (xavier-like scales so that 12 + 6
    post-norm layers stay well conditioned; adapters get a visible scale instead of the reference's 1e-6 init)."""
    rng = np.random.RandomState(seed)
    sd = {}
    for name, shape in sized_param_shapes(t, family).items():
        if 'layer_norm' in name:
            w = (1.0 + 0.1 * rng.randn(*shape)) if name.endswith('weight') else 0.05 * rng.randn(*shape)
        elif name.endswith('bias'):
            w = 0.02 * rng.randn(*shape)
        elif name.endswith('embed_positions.weight'):
            w = 0.05 * rng.randn(*shape)
        elif name == 'decoder.embed_tokens.weight':
            w = rng.randn(*shape) * shape[1] ** -0.5
            w[PAD] = 0
        else:
            fan_in = int(np.prod(shape[1:]))
            w = rng.randn(*shape) * fan_in ** -0.5
        if name.endswith('out_proj.weight'):
            # fan-in-scaled random attention blocks mix the tokens with near-uniform weights and, 12 post-norm layers deep, all
            # positions collapse onto one vector: the q / k gradients are then ~0 and any flash-style backward (ours, PyTorch's
            # fused SDPA alike: delta = rowsum(dO o O) from the bf16 output) returns noise for them. A smaller attention
            # branch keeps the token identity in the residual stream, as trained models do.
            w = w * 0.3
        if family == 'st' and name == 'encoder.subsample.conv_layers.1.weight':
            # random fan-in-scaled conv weights give |embed * sqrt(d)| ~ 10 and first-layer attention logits of ~ +-100: a
            # one-hot softmax that no 8-bit mantissa resolves (neither ours nor PyTorch's). Trained models sit at O(1).
            w = w * 0.15
        sd[name] = torch.from_numpy(w.astype(np.float32)).to(torch.bfloat16).float()
    if family == 'text':
        sd['encoder.embed_tokens.weight'] = sd['decoder.embed_tokens.weight']
    return sd


def sized_batch(seed: int, frames, dec_lens, vocab: int, text: bool = False) -> dict:
    """Ragged batch: `frames` feature frames (or source tokens when text) and `dec_lens` decoder-input lengths per line."""
    rng = np.random.RandomState(seed)
    B, Tm = len(frames), max(dec_lens)
    dec = torch.full((B, Tm), PAD, dtype=torch.int64)
    pm = torch.zeros(B, Tm, dtype=torch.bool)
    if text:
        x = torch.full((B, max(frames)), PAD, dtype=torch.int64)
    else:
        x = torch.zeros(B, max(frames), 80)
    for b, (f, t) in enumerate(zip(frames, dec_lens)):
        if text:
            x[b, :f] = torch.from_numpy(np.concatenate([rng.randint(4, vocab, f - 1), [EOS]]))
        else:
            x[b, :f] = torch.from_numpy(rng.randn(f, 80).astype(np.float32)).to(torch.bfloat16).float()
        dec[b, :t] = torch.from_numpy(np.concatenate([[BOS], rng.randint(4, vocab, t - 2), [EOS]]))
        pm[b, 0] = True
    return {'encoder_input': x, 'encoder_input_length': torch.tensor(frames), 'decoder_input': dec, 'prompt_mask': pm,
            'prompt_length': torch.ones(B, dtype=torch.int64), 'indices': torch.arange(B),
            'meta': {'source_lang': 'de', 'target_lang': 'en'}}


# --------------------------------------------------------------------------------------------------
# decoding: a deterministic stand-in decoder with the reference's call signature, for pinning the SEARCH algorithm (the model's
# incremental forward is pinned separately): next-token scores are a seeded function of the last two tokens, the prefix length
# and a per-sentence key read from the encoder output; EOS gets likelier with length, so hypotheses finish at different steps.
# --------------------------------------------------------------------------------------------------
MOCK_VOCAB = 40


class MockDecoder(torch.nn.Module):
    padding_idx, eos_idx, bos_idx = PAD, EOS, BOS

    def __init__(self, seed: int = 0, max_len: int = 32, eos_slope: float = 0.9):
        super().__init__()
        rng = np.random.RandomState(100 + seed)
        self.a = torch.from_numpy(rng.randn(MOCK_VOCAB, MOCK_VOCAB).astype(np.float32) * 1.5)
        self.b = torch.from_numpy(rng.randn(MOCK_VOCAB, MOCK_VOCAB).astype(np.float32))
        self.c = torch.from_numpy(rng.randn(MOCK_VOCAB).astype(np.float32))
        self._max_len, self.eos_slope = max_len, eos_slope

    @property
    def max_len(self) -> int:
        return self._max_len

    def forward(self, encoder_out, encoder_mask, decoder_input, state=None, meta=None, return_layers=(), **kwargs):
        prefix = decoder_input if not state or 'prefix' not in state else torch.cat([state['prefix'], decoder_input], dim=1)
        if state is not None:
            state['prefix'] = prefix
        T = decoder_input.size(1)
        key = encoder_out[:, 0, 0].float()
        rows = []
        for t in range(prefix.size(1) - T, prefix.size(1)):
            last = prefix[:, t].clamp(min=0)
            prev = prefix[:, t - 1].clamp(min=0) if t > 0 else last
            logits = self.a[last] + 0.7 * self.b[prev] + key.unsqueeze(1) * self.c.unsqueeze(0)
            logits[:, self.eos_idx] += self.eos_slope * (t - 2)
            logits[:, self.padding_idx] = -20.0
            rows.append(logits)
        return torch.stack(rows, dim=1), {}


def mock_encoder_out(batch_size: int, seed: int = 0):
    """(encoder_out (B, 3, 4), encoder_mask (B, 3)): only entry [b, 0, 0], the sentence key, matters to MockDecoder."""
    rng = np.random.RandomState(200 + seed)
    enc = torch.from_numpy(rng.randn(batch_size, 3, 4).astype(np.float32))
    return enc, torch.zeros(batch_size, 3, dtype=torch.bool)
