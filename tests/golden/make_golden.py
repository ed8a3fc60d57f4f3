"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference (naver/pasero,
imported from /root/reference) and torchaudio in the build container. Run once, commit the outputs:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so its own outputs on seeded inputs
are the pin. Two import stubs (tests/golden/_stubs: sacrebleu, stopes) and one harness-side shim (torch's
LRScheduler no longer accepts the positional `verbose` the reference passes, pasero/optimization.py:38) are
needed to import it; no reference file is modified or copied.
"""
import json
import os
import shutil
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('PASERO_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
from tests.golden import synth  # noqa: E402  (seeded synthetic inputs shared with the tests)


def import_reference():
    warnings.filterwarnings('ignore')
    sys.path.insert(0, os.path.join(HERE, '_stubs'))
    sys.path.insert(0, REF)
    orig = torch.optim.lr_scheduler.LRScheduler.__init__

    def init(self, optimizer, last_epoch=-1, verbose=None):
        orig(self, optimizer, last_epoch)
    torch.optim.lr_scheduler.LRScheduler.__init__ = init
    import pasero.tasks, pasero.datasets, pasero.models  # noqa: F401,E401  (fills the registries, cli/train.py:28-33)
    import pasero
    return pasero


def make_task_and_model(pasero, tmp, extra=(), adapters=False, by_langs=False):
    from pasero.config import TrainingConfig, get_task_class, get_architecture
    dict_path = os.path.join(tmp, 'dict.txt')
    with open(dict_path, 'w') as f:
        for i in range(synth.TINY['vocab_size'] - 4):
            f.write(f'tok{i} 1\n')
    shutil.copy(os.path.join(REF, 'examples/TED/de-en/bpecodes'), os.path.join(tmp, 'bpecodes'))
    t = synth.TINY
    args = ['--task', 'speech_translation', '--arch', 'transformer',
            '--encoder-layers', str(t['encoder_layers']), '--decoder-layers', str(t['decoder_layers']),
            '--embed-dim', str(t['embed_dim']), '--encoder-ffn-dim', str(t['ffn_dim']), '--decoder-ffn-dim', str(t['ffn_dim']),
            '--encoder-attention-heads', str(t['heads']), '--decoder-attention-heads', str(t['heads']),
            '--input-dim', '80', '--conv-input-dim', '80', '--conv-channels', str(t['conv_channels']),
            '--conv-kernel-sizes', '5', '5', '--data-dir', tmp, '--dict', 'dict.txt', '--tokenizer-path', 'bpecodes',
            '--model-dir', os.path.join(tmp, 'model'),
            '--encoder-max-len', '3000', '--batch-size', '64000', '--dropout', '0', '--label-smoothing', '0.1',
            '--metrics', '--early-stopping-metric', 'nll_loss', '--max-steps', '1000', '--valid-interval', '1000',
            '--dataset-type', 'simple', '--dtype', 'float32', '--lr', str(synth.TRAIN['lr']),
            '--warmup', str(synth.TRAIN['warmup']), '--seed', '1234']
    if by_langs:
        args += ['--source-langs', *synth.BY_LANGS['encoder'], '--target-langs', *synth.BY_LANGS['decoder']]
    else:
        args += ['--source-lang', 'de', '--target-lang', 'en']
    args += list(extra)       # (later occurrences of an option win)
    cfg = TrainingConfig(*args)
    task = get_task_class(cfg.task)(cfg.data_dir, cfg.task_cfg)
    task.setup_for_model(cfg.model_cfg)
    cfg.dp_size, cfg.dp_rank, cfg.dp_local_size, cfg.dp_local_rank, cfg.start_rank = 1, 0, 1, 0, 0  # utils.setup_distributed
    torch.manual_seed(1234)
    model = get_architecture(cfg.model_cfg)(cfg.model_cfg, cfg, task)
    # seeded, reference-independent weights (bf16-representable), so the fixtures need not carry them
    sd = synth.by_adapter_state_dict() if by_langs else synth.adapter_state_dict() if adapters else synth.tiny_state_dict()
    assert {k: tuple(v.shape) for k, v in model.named_parameters()} == (synth.by_adapter_param_shapes() if by_langs else
                                                                        synth.adapter_param_shapes() if adapters else synth.tiny_param_shapes())
    missing = model.load_state_dict(sd, strict=False)
    assert all('_float_tensor' in k for k in missing.missing_keys) and not missing.unexpected_keys, missing
    return cfg, task, model


def make_whisper_model(pasero, tmp):
    """Whisper-style tiny model (`--arch whisper_base` scaled down; pasero/config.py:2539-2566) on the same dictionary."""
    from pasero.config import TrainingConfig, get_task_class, get_architecture
    t = synth.WSP
    args = ['--task', 'speech_translation', '--arch', 'whisper_base',
            '--encoder-layers', str(t['encoder_layers']), '--decoder-layers', str(t['decoder_layers']),
            '--embed-dim', str(t['embed_dim']), '--encoder-ffn-dim', str(t['ffn_dim']), '--decoder-ffn-dim', str(t['ffn_dim']),
            '--encoder-attention-heads', str(t['heads']), '--decoder-attention-heads', str(t['heads']),
            '--conv-channels', str(t['conv_channels']), '--data-dir', tmp, '--dict', 'dict.txt', '--tokenizer-path', 'bpecodes',
            '--source-lang', 'de', '--target-lang', 'en', '--model-dir', os.path.join(tmp, 'model_wsp'),
            '--encoder-max-len', str(t['encoder_max_len']), '--decoder-max-len', str(t['decoder_max_len']),
            '--batch-size', '64000', '--dropout', '0', '--label-smoothing', '0.1', '--metrics',
            '--early-stopping-metric', 'nll_loss', '--max-steps', '1000', '--valid-interval', '1000', '--dataset-type', 'simple',
            '--dtype', 'float32', '--seed', '1234', '--padding-idx', '1', '--eos-idx', '2', '--bos-idx', '2']
    cfg = TrainingConfig(*args)
    task = get_task_class(cfg.task)(cfg.data_dir, cfg.task_cfg)
    task.setup_for_model(cfg.model_cfg)
    cfg.dp_size, cfg.dp_rank, cfg.dp_local_size, cfg.dp_local_rank, cfg.start_rank = 1, 0, 1, 0, 0
    torch.manual_seed(1234)
    model = get_architecture(cfg.model_cfg)(cfg.model_cfg, cfg, task)
    assert {k: tuple(v.shape) for k, v in model.named_parameters()} == synth.wsp_param_shapes(), \
        {k: tuple(v.shape) for k, v in model.named_parameters() if synth.wsp_param_shapes().get(k) != tuple(v.shape)}
    missing = model.load_state_dict(synth.wsp_state_dict(), strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    return cfg, task, model


def golden_model_adapters(pasero, tmp, out):
    """`--arch adapter_transformer` on the tiny ST model (pasero/models/adapters.py): adapters trained, base frozen."""
    cfg, task, model = make_task_and_model(pasero, tmp, ('--arch', 'adapter_transformer', '--encoder-adapter-dim',
                                                                    str(synth.ADAPTER_DIM), '--decoder-adapter-dim',
                                                                    str(synth.ADAPTER_DIM)), adapters=True)
    model.train()
    batch = synth.model_batch()
    enc_out, enc_mask, _ = model.encoder(batch['encoder_input'].clone(), batch['encoder_input_length'], meta=batch['meta'])
    logits, _ = model.decoder(enc_out, enc_mask, batch['decoder_input'][:, :-1], meta=batch['meta'])
    loss, logs = model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
    model.zero_grad()
    loss.backward()
    trainable = [k for k, p in model.named_parameters() if p.requires_grad]
    assert all('adapters' in k for k in trainable) and trainable
    torch.save({
        'loss': loss.item(), 'logs': logs, 'encoder_out': enc_out.detach(), 'logits': logits.detach(),
        'trainable': trainable, 'param_names': [k for k, _ in model.named_parameters()],
        'grads': {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None},
    }, os.path.join(out, 'st_adapter_model.pt'))
    print('adapter model golden: loss', loss.item(), logs, 'trainable', len(trainable), 'of', len(list(model.parameters())))


def golden_model_text(pasero, tmp, out):
    """`--task translation --arch transformer` (token-id encoder, shared embeddings) on the tiny dimensions."""
    from pasero.config import TrainingConfig, get_task_class, get_architecture
    t = synth.TINY
    args = ['--task', 'translation', '--arch', 'transformer',
            '--encoder-layers', str(t['encoder_layers']), '--decoder-layers', str(t['decoder_layers']),
            '--embed-dim', str(t['embed_dim']), '--encoder-ffn-dim', str(t['ffn_dim']), '--decoder-ffn-dim', str(t['ffn_dim']),
            '--encoder-attention-heads', str(t['heads']), '--decoder-attention-heads', str(t['heads']),
            '--data-dir', tmp, '--dict', 'dict.txt', '--tokenizer-path', 'bpecodes', '--source-lang', 'de', '--target-lang', 'en',
            '--model-dir', os.path.join(tmp, 'model_text'), '--batch-size', '4000', '--dropout', '0', '--label-smoothing', '0.1',
            '--metrics', '--early-stopping-metric', 'nll_loss', '--max-steps', '1000', '--valid-interval', '1000',
            '--dataset-type', 'simple', '--dtype', 'float32', '--seed', '1234']
    cfg = TrainingConfig(*args)
    task = get_task_class(cfg.task)(cfg.data_dir, cfg.task_cfg)
    task.setup_for_model(cfg.model_cfg)
    cfg.dp_size, cfg.dp_rank, cfg.dp_local_size, cfg.dp_local_rank, cfg.start_rank = 1, 0, 1, 0, 0
    torch.manual_seed(1234)
    model = get_architecture(cfg.model_cfg)(cfg.model_cfg, cfg, task)
    assert model.encoder.embed_tokens is model.decoder.embed_tokens
    want = synth.text_param_shapes()
    assert {k: tuple(v.shape) for k, v in model.state_dict().items() if '_float_tensor' not in k} == want
    missing = model.load_state_dict(synth.text_state_dict(), strict=False)
    assert all('_float_tensor' in k for k in missing.missing_keys) and not missing.unexpected_keys, missing
    model.train()
    batch = synth.text_batch()
    names = ['enc_0', f"enc_{t['encoder_layers'] - 1}", f"dec_{t['decoder_layers'] - 1}"]
    enc_out, enc_mask, enc_layers = model.encoder(batch['encoder_input'].clone(), batch['encoder_input_length'], return_layers=names)
    logits, dec_layers = model.decoder(enc_out, enc_mask, batch['decoder_input'][:, :-1], return_layers=names)
    loss, logs = model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
    model.zero_grad()
    loss.backward()
    torch.save({
        'loss': loss.item(), 'logs': logs, 'encoder_out': enc_out.detach(), 'encoder_mask': enc_mask, 'logits': logits.detach(),
        'layers': {k: v.detach() for k, v in {**enc_layers, **dec_layers}.items()},
        'grads': {k: synth.summarize_grad(p.grad) for k, p in model.named_parameters()},
        'param_names': [k for k, _ in model.named_parameters()],
    }, os.path.join(out, 'st_text_model.pt'))
    print('text model golden: loss', loss.item(), logs)


def golden_preprocess_tags(pasero, tmp, out):
    """`TranslationTask.preprocess` / `SpeechTranslationTask.preprocess` with language-code and domain tags
    (pasero/tasks/translation.py:369-540, speech_translation.py:42-86; `lang_code: True` is what the TED-top20 recipe of
    BASELINE config 5 trains with). The fixture carries the dictionary, the BPE-split lines (subword splitting stays with
    the reference's tokenizer) and the arrays the reference produced from them."""
    import json
    from pasero.config import TrainingConfig, get_task_class
    from pasero.preprocessing import split_tags
    lines = [('Guten Morgen , wie geht es Ihnen heute ?', 'Good morning , how are you today ?'),
             ('Das ist ein kleiner Test .', 'This is a small test .'),
             ('<lang:en> Hallo Welt', '<lang:en> <domain:ted> Hello world'),
             ('Eine sehr lange Zeile ' + 'mit vielen Worten ' * 6, 'A very long line ' + 'with many words ' * 6),
             ('Ja', 'Yes indeed it is')]
    d = os.path.join(tmp, 'tags')
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(REF, 'examples/TED/de-en/bpecodes'), os.path.join(d, 'bpecodes'))

    def make(task_name, flags, dict_tokens):
        with open(os.path.join(d, 'dict.txt'), 'w') as f:
            for tok in dict_tokens:
                f.write(f'{tok} 1\n')
        args = ['--task', task_name, '--arch', 'transformer', '--data-dir', d, '--dict', 'dict.txt', '--tokenizer-path', 'bpecodes',
                '--source-lang', 'de', '--target-lang', 'en', '--model-dir', os.path.join(d, 'model'), '--metrics',
                '--early-stopping-metric', 'nll_loss', '--dataset-type', 'simple', '--dtype', 'float32', '--max-steps', '1000',
                '--valid-interval', '1000', '--batch-size', '4000'] + list(flags)
        if task_name == 'speech_translation':
            args += ['--input-dim', '80', '--conv-input-dim', '80', '--conv-channels', '64', '--conv-kernel-sizes', '5', '5',
                     '--encoder-max-len', '3000']
        cfg = TrainingConfig(*args)
        task = get_task_class(cfg.task)(cfg.data_dir, cfg.task_cfg)
        task.setup_for_model(cfg.model_cfg)
        return cfg, task

    # vocabulary = the subwords of the lines above (minus a few, to exercise <unk>) + the tags
    with open(os.path.join(REF, 'examples/TED/de-en/dict.txt')) as f:
        real_vocab = [line.split()[0] for line in f if line.strip()]
    _, probe = make('translation', [], real_vocab)
    subwords = []
    for src, tgt in lines:
        for side, text in (('src', src), ('tgt', tgt)):
            *_, rest = split_tags(text)
            for tok in (probe.src_preprocessor if side == 'src' else probe.tgt_preprocessor).tokenize(rest):
                if tok not in subwords:
                    subwords.append(tok)
    dropped = subwords[3::11]
    tags = ['<lang:de>', '<lang:en>', '<lang:fr>', '<domain:ted>', '<domain:news>']
    dict_tokens = [t for t in subwords if t not in dropped] + tags

    configs = {
        'lang_code': ['--lang-code'],
        'both_codes': ['--source-lang-code', '--target-lang-code'],
        'domain_and_cfg_tags': ['--domain-tag', '--target-tags', '<lang:fr>', '--source-tags', '<domain:news>'],
        'plain': [],
        'short': ['--lang-code', '--target-lang-code', '--max-source-len', '7', '--max-target-len', '6'],
    }
    metas = [{'source_lang': 'de', 'target_lang': 'en', 'domain': 'ted'},
             {'source_lang': 'de', 'target_lang': 'en', 'domain': 'ted', 'target_tags': ['<domain:ted>', '<lang:en>']}]
    cases = []
    for name, flags in configs.items():
        for task_name in ('translation', 'speech_translation'):
            cfg, task = make(task_name, flags, dict_tokens)
            for li, (src, tgt) in enumerate(lines):
                for meta in metas:
                    for tokenize, truncate in ((True, False), (False, False), (True, True)):
                        *src_tags, src_rest = split_tags(src)
                        *tgt_tags, tgt_rest = split_tags(tgt)
                        if task_name == 'speech_translation':
                            frames = 20 + 13 * li
                            source_ref = np.zeros((frames, 80), dtype=np.float16)
                            tgt_text, tgt_pre = tgt_rest, ' '.join(task.tgt_preprocessor.tokenize(tgt_rest))
                            if not tokenize:
                                tgt_text = tgt_pre     # tokenize=False: the reference splits the given line on blanks
                            res = task.preprocess({'source': source_ref, 'target': tgt_text, 'meta': dict(meta)},
                                                  truncate=truncate, tokenize=tokenize, append_eos=True)
                            case = {'task': task_name, 'config': name, 'flags': flags, 'meta': meta, 'tokenize': tokenize,
                                    'truncate': truncate, 'frames': frames, 'target_bpe': tgt_pre}
                            if res:
                                case['encoder_len'] = int(len(res['encoder_input']))
                        else:
                            src_pre = ' '.join(src_tags + task.src_preprocessor.tokenize(src_rest))
                            tgt_pre = ' '.join(tgt_tags + task.tgt_preprocessor.tokenize(tgt_rest))
                            s_in, t_in = (src, tgt) if tokenize else (src_pre, tgt_pre)
                            res = task.preprocess({'source': s_in, 'target': t_in, 'meta': dict(meta)},
                                                  truncate=truncate, tokenize=tokenize, append_eos=True)
                            case = {'task': task_name, 'config': name, 'flags': flags, 'meta': meta, 'tokenize': tokenize,
                                    'truncate': truncate, 'source_bpe': src_pre, 'target_bpe': tgt_pre}
                            if res:
                                case['encoder_input'] = res['encoder_input'].tolist()
                        case['skipped'] = not res
                        if res:
                            case['decoder_input'] = res['decoder_input'].tolist()
                            case['prompt_mask'] = res['prompt_mask'].astype(int).tolist()
                        case['max_source_len'] = int(task.max_source_len)
                        case['max_target_len'] = int(task.max_target_len)
                        cases.append(case)
    with open(os.path.join(out, 'preprocess_tags.json'), 'w') as f:
        json.dump({'dict_tokens': dict_tokens, 'cases': cases}, f)
    print('preprocess/tags golden:', len(cases), 'cases,', sum(c['skipped'] for c in cases), 'skipped,', len(dict_tokens), 'dictionary tokens')


def golden_model_whisper(pasero, cfg, task, model, out):
    model.train()
    batch = synth.model_batch()
    names = ['enc_0', f"enc_{synth.WSP['encoder_layers'] - 1}", 'dec_0', f"dec_{synth.WSP['decoder_layers'] - 1}"]
    enc_out, enc_mask, enc_layers = model.encoder(batch['encoder_input'].clone(), batch['encoder_input_length'], return_layers=names)
    logits, dec_layers = model.decoder(enc_out, enc_mask, batch['decoder_input'][:, :-1], return_layers=names)
    loss, logs = model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    torch.save({
        'cfg': synth.WSP, 'state_dict_keys': sorted(model.state_dict().keys()),
        'loss': loss.item(), 'logs': logs, 'encoder_out': enc_out.detach(), 'encoder_mask': enc_mask, 'logits': logits.detach(),
        'layers': {k: v.detach() for k, v in {**enc_layers, **dec_layers}.items()},
        'grads': {k: synth.summarize_grad(v) for k, v in grads.items()},
        'param_names': [k for k, _ in model.named_parameters()],
    }, os.path.join(out, 'st_wsp_model.pt'))
    print('whisper-style model golden: loss', loss.item(), logs, 'params', sum(p.numel() for p in model.parameters()))


def golden_fbank(out):
    import torchaudio
    res = {}
    for name, n in synth.FBANK_CASES.items():
        w = synth.waveform(name, n)
        wt = torch.from_numpy(w)
        ref32 = torchaudio.compliance.kaldi.fbank((wt * 2 ** 15)[None], num_mel_bins=80, sample_frequency=16000.0)
        ref64 = torchaudio.compliance.kaldi.fbank((wt.double() * 2 ** 15)[None], num_mel_bins=80, sample_frequency=16000.0)
        res[f'{name}_f32'] = ref32.numpy().astype(np.float32)
        res[f'{name}_f64'] = ref64.numpy().astype(np.float64)
    np.savez_compressed(os.path.join(out, 'fbank_torchaudio.npz'), **res)
    print('fbank goldens:', {k: v.shape for k, v in res.items()})


def golden_batching(pasero, task, out):
    from pasero import utils, datasets
    res = {}
    # docstring known-answer vectors of the reference (utils.py:797-813, :253, :265)
    sizes = [1000, 2000, 4000, 500, 1000, 1000]
    kat = utils.build_batches([0, 1, 2, 3, 4, 5], sizes.__getitem__, 4000, 1, 2)
    assert kat == [[0, 1], [2], [3, 4], [5]], kat
    assert utils.mask_to_len(torch.tensor([[0, 1, 0, 1, 0], [1, 1, 0, 0, 0]]).bool()).tolist() == [4, 2]
    assert utils.len_to_mask(torch.tensor([0, 3, 1])).long().tolist() == [[1, 1, 1], [0, 0, 0], [0, 1, 1]]
    samples = synth.batching_samples()
    combos = []
    for bs in (4000, 40000):
        for mult in (None, 8):
            for lines in (None, 50):
                task.cfg.batch_size_multiple = mult
                task.cfg.lines_per_batch = lines
                ds = types.SimpleNamespace(batch_by=[], task=task, shuffle=True, batch_size=bs)
                np.random.seed(synth.BATCHING_SEED)
                batches = datasets.TrainingDataset.buffered_batching(ds, list(samples))
                flat = np.array([s['index'] for b in batches for s in b], dtype=np.int64)
                sizes_ = np.array([len(b) for b in batches], dtype=np.int64)
                key = f'bs{bs}_m{mult or 0}_l{lines or 0}'
                res[key + '_flat'], res[key + '_sizes'] = flat, sizes_
                combos.append((key, len(batches), int(flat.size)))
    task.cfg.batch_size_multiple = None
    task.cfg.lines_per_batch = None
    # collated tensors of one small batch (float16 features -> model dtype, padding values, lengths, prompt length)
    np.random.seed(synth.BATCHING_SEED)
    small = task.build_batches(list(samples[:64]), shuffle=True, batch_size=4000)[3]
    col = task.collate(small, padding_idx=task.padding_idx, dtype=torch.float32, model_type='encoder_decoder')
    res['collate_indices_in'] = np.array([s['index'] for s in small], dtype=np.int64)
    for k in ('decoder_input', 'prompt_mask', 'prompt_length', 'indices', 'encoder_input_length'):
        res['collate_' + k] = col[k].numpy()
    res['collate_encoder_input_sum'] = col['encoder_input'].double().sum(dim=(1, 2)).numpy()
    res['collate_encoder_input_shape'] = np.array(col['encoder_input'].shape)
    np.savez_compressed(os.path.join(out, 'batching.npz'), **res)
    print('batching goldens:', combos)


def golden_batching_by(pasero, task, out):
    """TrainingDataset.buffered_batching with `batch_by` (pasero/datasets.py:708-730; set by cli/train.py:275 from the model's
    `batch_by` when adapters are selected per language): every batch holds one (source, target) language pair, and the
    collated batch keeps exactly the meta entries its samples share (tasks/task.py:531-571)."""
    from pasero import datasets
    samples = synth.batching_samples_by_lang()
    ds = types.SimpleNamespace(batch_by={'source_lang', 'target_lang'}, task=task, shuffle=True, batch_size=4000)
    np.random.seed(synth.BATCHING_SEED)
    batches = datasets.TrainingDataset.buffered_batching(ds, list(samples))
    langs = [{(s['meta']['source_lang'], s['meta']['target_lang']) for s in b} for b in batches]
    assert all(len(x) == 1 for x in langs)
    col = task.collate(batches[0], padding_idx=task.padding_idx, dtype=torch.float32, model_type='encoder_decoder')
    assert 'uid' not in col['meta'] and set(col['meta']) == {'source_lang', 'target_lang'}, col['meta']
    np.savez_compressed(os.path.join(out, 'batching_by.npz'),
                        flat=np.array([s['meta']['uid'] for b in batches for s in b], dtype=np.int64),
                        index=np.array([s['index'] for b in batches for s in b], dtype=np.int64),
                        sizes=np.array([len(b) for b in batches], dtype=np.int64))
    print('batching-by-language golden:', len(batches), 'batches,', sorted({next(iter(x)) for x in langs}))


def golden_model(pasero, cfg, task, model, out):
    model.train()
    batch = synth.model_batch()
    names = ['enc_0', f"enc_{synth.TINY['encoder_layers'] - 1}", 'dec_0', f"dec_{synth.TINY['decoder_layers'] - 1}"]
    enc_out, enc_mask, enc_layers = model.encoder(batch['encoder_input'].clone(), batch['encoder_input_length'], return_layers=names)
    logits, dec_layers = model.decoder(enc_out, enc_mask, batch['decoder_input'][:, :-1], return_layers=names)
    loss, logs = model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    torch.save({
        'tiny': synth.TINY, 'state_dict_keys': sorted(model.state_dict().keys()),
        'loss': loss.item(), 'logs': logs, 'encoder_out': enc_out.detach(), 'encoder_mask': enc_mask,
        'logits': logits.detach(),
        'layers': {k: v.detach() for k, v in {**enc_layers, **dec_layers}.items()},
        'grads': {k: synth.summarize_grad(v) for k, v in grads.items()},
        'param_names': [k for k, _ in model.named_parameters()],
    }, os.path.join(out, 'st_tiny_model.pt'))
    print('model golden: loss', loss.item(), logs, 'params', sum(p.numel() for p in model.parameters()))


def golden_training_curve(pasero, cfg, task, model, out):
    """200 optimizer steps through the reference Trainer.train_step; per-step logs captured by wrapping Metrics.update."""
    from pasero.training import Trainer
    from pasero import utils
    os.makedirs(cfg.model_dir, exist_ok=True)
    metrics = utils.Metrics()
    rows = []
    current = {}
    orig_update = metrics.update

    def update(name, value, *a, **kw):
        current[name] = value
        return orig_update(name, value, *a, **kw)
    metrics.update = update
    trainer = Trainer(cfg, task, model, metrics, None)
    batches = synth.train_batches()

    def it():
        i = 0
        while True:
            b = batches[i % len(batches)]
            yield {k: (v.clone() if torch.is_tensor(v) else dict(v)) for k, v in b.items()}
            i += 1
    iterator = it()
    for step in range(synth.TRAIN['steps']):
        current.clear()
        trainer.train_step(iterator)
        rows.append({k: float(current[k]) for k in ('loss', 'nll_loss', 'num_tokens', 'num_lines', 'gnorm', 'lr')})
        if step == 2:
            sd3 = {k: v.detach().clone() for k, v in model.state_dict().items()}
            opt3 = {n: {kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in trainer.optimizer.state[p].items()}
                    for n, p in model.named_parameters() if p in trainer.optimizer.state}
            trainer.save()
            files = sorted(os.listdir(cfg.model_dir))
            layout = {}
            for fn in files:
                path = os.path.join(cfg.model_dir, fn)
                if fn.endswith('.bin') and not os.path.islink(path):
                    ck = torch.load(path, map_location='cpu', weights_only=False)
                    layout[fn] = {k: (sorted(v.keys()) if isinstance(v, dict) else type(v).__name__) for k, v in ck.items()}
                    if fn.startswith('optimizer'):
                        st = ck['optimizer']['state']
                        first = next(iter(st))
                        layout[fn]['optimizer.state.first_key'] = str(first)
                        layout[fn]['optimizer.state.entry_keys'] = sorted(st[first].keys())
                        layout[fn]['optimizer.keys'] = sorted(ck['optimizer'].keys())
                elif os.path.islink(path):
                    layout[fn] = 'symlink -> ' + os.readlink(path)
            with open(os.path.join(out, 'checkpoint_layout.json'), 'w') as f:
                json.dump(layout, f, indent=1, default=str)
    torch.save({
        'train': synth.TRAIN, 'curve': rows,
        'state_dict_step3': {k: synth.summarize_grad(v) for k, v in sd3.items() if v.is_floating_point() and v.numel() > 1},
        'optimizer_step3': {n: {'step': s['step'], 'exp_avg_sum': s['exp_avg'].double().sum().item(),
                                'exp_avg_sq_sum': s['exp_avg_sq'].double().sum().item()} for n, s in opt3.items()},
        'state_dict_final_checksum': {k: v.double().sum().item() for k, v in model.state_dict().items()},
    }, os.path.join(out, 'st_tiny_train.pt'))
    print('training golden: first', rows[0], 'last', rows[-1])


def golden_training_adapters_by(pasero, tmp, out):
    """`--encoder-adapters-by source_lang --decoder-adapters-by target_lang` through the reference's Trainer.train_step (CPU,
    fp32): the language pair changes from step to step (synth.BY_SCHEDULE), so every step leaves some adapter sets without a
    gradient. Recorded: per-step logs, which parameters each step changed, the optimizer's per-parameter step counts and
    moments at the end."""
    from pasero.training import Trainer
    from pasero import utils
    bt = synth.BY_TRAIN
    cfg, task, model = make_task_and_model(pasero, tmp, ('--arch', 'adapter_transformer', '--encoder-adapter-dim', str(synth.ADAPTER_DIM),
                                                         '--decoder-adapter-dim', str(synth.ADAPTER_DIM),
                                                         '--encoder-adapters-by', 'source_lang', '--decoder-adapters-by', 'target_lang',
                                                         '--lr', str(bt['lr']), '--warmup', str(bt['warmup']),
                                                         '--model-dir', os.path.join(tmp, 'model_by')), by_langs=True)
    assert model.find_unused_parameters
    os.makedirs(cfg.model_dir, exist_ok=True)
    metrics = utils.Metrics()
    current = {}
    orig_update = metrics.update

    def update(name, value, *a, **kw):
        current[name] = value
        return orig_update(name, value, *a, **kw)
    metrics.update = update
    trainer = Trainer(cfg, task, model, metrics, None)
    rows, changed = [], []
    for step in range(bt['steps']):
        b = synth.by_train_batch(step)
        before = {k: v.detach().clone() for k, v in model.named_parameters() if v.requires_grad}
        current.clear()
        trainer.train_step(iter([{k: (v.clone() if torch.is_tensor(v) else dict(v)) for k, v in b.items()}]))
        rows.append({k: float(current[k]) for k in ('loss', 'nll_loss', 'num_tokens', 'num_lines', 'gnorm', 'lr')})
        changed.append(sorted({adapter_set(k) for k, v in model.named_parameters()
                               if v.requires_grad and not torch.equal(before[k], v.detach())}))
    opt = {n: trainer.optimizer.state[p] for n, p in model.named_parameters() if p in trainer.optimizer.state and trainer.optimizer.state[p]}
    trainer.save()
    ck = torch.load(os.path.join(cfg.model_dir, f"optimizer_{bt['steps']}.bin"), map_location='cpu', weights_only=False)
    torch.save({
        'train': bt, 'schedule': synth.BY_SCHEDULE, 'curve': rows, 'changed': changed,
        'optimizer_steps': {n: int(s['step']) for n, s in opt.items()},
        'optimizer_checkpoint_steps': {str(k): int(v['step']) for k, v in ck['optimizer']['state'].items()},
        'exp_avg': {n: synth.summarize_grad(s['exp_avg']) for n, s in opt.items()},
        'params_final': {k: synth.summarize_grad(v.detach()) for k, v in model.named_parameters() if v.requires_grad},
        'trainable': [k for k, v in model.named_parameters() if v.requires_grad],
    }, os.path.join(out, 'st_adapters_by_train.pt'))
    print('adapters_by training golden: first', rows[0], 'last', rows[-1])
    print('  changed per step:', changed[:8])
    print('  step counts:', sorted({(adapter_set(n), s) for n, s in ((n, int(s['step'])) for n, s in opt.items())}))


def adapter_set(name: str) -> str:
    parts = name.split('.')
    return f"{parts[0]}/{parts[parts.index('adapters') + 1]}" if 'adapters' in parts else 'base'


DECODE_CASES = [dict(name='beam5', beam_size=5, max_output_len=12, prompt=False, len_penalty=1.0),
                dict(name='greedy', beam_size=1, max_output_len=12, prompt=False, len_penalty=1.0),
                dict(name='beam3_prompt', beam_size=3, max_output_len=9, prompt=True, len_penalty=0.7),
                dict(name='beam4_short', beam_size=4, max_output_len=3, prompt=False, len_penalty=1.0)]


def decode_prompts(batch_size: int) -> torch.Tensor:
    """BOS (+ forced target tags for some sentences), right-padded: the `decoder_input` a validation batch carries after
    Trainer.inference_step cut it back to `prompt_length`."""
    rows = [[synth.BOS, 7, 9], [synth.BOS], [synth.BOS, 11], [synth.BOS, 7, 9]]
    rows = [rows[i % len(rows)] for i in range(batch_size)]
    width = max(len(r) for r in rows)
    return torch.tensor([r + [synth.PAD] * (width - len(r)) for r in rows], dtype=torch.long)


def golden_decoding(pasero, cfg, task, model, out):
    """pasero.decoding.beam_search (the unmodified reference, CPU, fp32) on the tiny ST model and the seeded model batch:
    tokens, scores and per-position scores of every finished hypothesis, for the cases of DECODE_CASES."""
    import json
    from pasero import decoding
    model.eval()
    batch = synth.model_batch()
    with torch.no_grad():
        enc_out, enc_mask, _ = model.encoder(batch['encoder_input'].clone(), batch['encoder_input_length'])
        cases = {}
        for case in DECODE_CASES:
            # (decoder_input=None would reach EnsembleDecoder.bos_idx, which the reference's wrapper lacks: always pass BOS)
            dec_in = decode_prompts(enc_out.size(0)) if case['prompt'] else torch.full((enc_out.size(0), 1), synth.BOS, dtype=torch.long)
            res = decoding.beam_search(model.decoder, enc_out, enc_mask, case['max_output_len'], case['beam_size'], {},
                                       decoder_input=dec_in, return_scores=True, len_penalty=case['len_penalty'])
            cases[case['name']] = [[{'tokens': [int(t) for t in h['tokens']], 'score': float(h['score']),
                                     'normalized_score': float(h['normalized_score']),
                                     'pos_scores': [float(x) for x in h['pos_scores']]} for h in nbest] for nbest in res]
            print(case['name'], [[h['tokens'] for h in nbest][:2] for nbest in cases[case['name']]][:2])
    # the search algorithm alone, on the deterministic stand-in decoder (synth.MockDecoder): hypotheses that end at different
    # steps, sentences that finish early, forced prompts, the length limit
    mock = {}
    for seed in range(4):
        dec = synth.MockDecoder(seed, max_len=32, eos_slope=0.9 if seed < 3 else 0.05)
        enc, mask = synth.mock_encoder_out(5, seed)
        for case in DECODE_CASES:
            dec_in = decode_prompts(5) if case['prompt'] else torch.full((5, 1), synth.BOS, dtype=torch.long)
            with torch.no_grad():
                res = decoding.beam_search(dec, enc, mask, case['max_output_len'], case['beam_size'], {}, decoder_input=dec_in,
                                           return_scores=True, len_penalty=case['len_penalty'])
            mock[f"{case['name']}_seed{seed}"] = [[{'tokens': [int(t) for t in h['tokens']], 'score': float(h['score']),
                                                   'normalized_score': float(h['normalized_score']),
                                                   'pos_scores': [float(x) for x in h['pos_scores']]} for h in nbest] for nbest in res]
    print('mock lengths', sorted({len(h['tokens']) for r in mock.values() for nb in r for h in nb}),
          'nbest sizes', sorted({len(nb) for r in mock.values() for nb in r}))
    with open(os.path.join(out, 'beam_search.json'), 'w') as f:
        json.dump({'cases': DECODE_CASES, 'results': cases, 'mock': mock}, f)


def numpy_file_features():
    """Seeded utterances of the NumpyFile fixture (float32 in, stored as float16 like examples/*/extract-features.py)."""
    rng = np.random.RandomState(31)
    return [rng.randn(n, 80).astype(np.float32) for n in (7, 1, 23, 40, 3)]


def golden_numpy_file(pasero, out):
    """A feature file written by the reference's own NumpyFile.build (pasero/files.py:122-159), with two unused slots
    (num_feats > number of utterances: positions stay 0 and the reader must skip them, files.py:112-115), plus what the
    reference's reader returns for it."""
    from pasero.files import NumpyFile
    path = os.path.join(out, 'features_ref.npy')
    f = NumpyFile.build(path, numpy_file_features(), dtype='float16', num_feats=7)
    idx, lengths = f.get_positions()
    rows = [next(f) for _ in range(len(idx))]
    f.seek(2)
    third = next(f)
    f.close()
    np.savez(os.path.join(out, 'features_ref_read.npz'), lengths=lengths, third=third,
             checksum=np.array([float(r.astype(np.float64).sum()) for r in rows]),
             **{f'row{i}': r for i, r in enumerate(rows)})
    print('numpy file golden:', os.path.getsize(path), 'bytes,', len(rows), 'utterances, lengths', lengths.tolist())


def reference_option_names(pasero):
    """Names of every option the reference's config classes declare (pasero/config.py: `argument` class attributes of all
    Config subclasses). pasero_b200.config uses the list to tell an option of the reference that this path does not read
    (kept, ignored) from a misspelt one (an error, as in pasero/config.py:1748-1750). Names only, no code."""
    import inspect
    from pasero import config as C
    names = set()
    for _, cls in inspect.getmembers(C, inspect.isclass):
        if issubclass(cls, C.Config):
            for klass in cls.__mro__:
                for k, v in list(vars(klass).items()):
                    if isinstance(v, C.argument) or (k in getattr(klass, '__annotations__', {}) and not k.startswith('_')):
                        names.add(k)
                    if isinstance(v, C.argument):      # `--domain` for `domains`, `--source-lang` for `source_langs`, ...
                        for alias in (getattr(v, 'aliases', None) or ()):
                            names.add(alias.lstrip('-').replace('-', '_'))
    path = os.path.join(ROOT, 'pasero_b200', 'reference_options.json')
    with open(path, 'w') as f:
        json.dump(sorted(names), f, indent=0)
    print('reference option names:', len(names))


def main():
    out = HERE
    pasero = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == 'options':      # only the option-name list
        reference_option_names(pasero)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'decode':       # only the beam-search fixture
        tmp = tempfile.mkdtemp()
        try:
            cfg, task, model = make_task_and_model(pasero, tmp)
            golden_decoding(pasero, cfg, task, model, out)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'numpyfile':    # only the NumpyFile fixture
        golden_numpy_file(pasero, out)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'text':         # only the text-encoder fixture
        tmp = tempfile.mkdtemp()
        try:
            make_task_and_model(pasero, tmp)                  # writes dict.txt / bpecodes
            golden_model_text(pasero, tmp, out)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'tags':         # only the preprocess / tags fixture
        tmp = tempfile.mkdtemp()
        try:
            golden_preprocess_tags(pasero, tmp, out)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'adapters':     # only the adapter fixture
        tmp = tempfile.mkdtemp()
        try:
            golden_model_adapters(pasero, tmp, out)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'adapters_by':  # only the per-language adapter fixtures (batching + training)
        tmp = tempfile.mkdtemp()
        try:
            cfg, task, model = make_task_and_model(pasero, tmp)
            golden_batching_by(pasero, task, out)
            golden_training_adapters_by(pasero, tmp, out)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'whisper':      # only the Whisper-style fixture
        tmp = tempfile.mkdtemp()
        try:
            make_task_and_model(pasero, tmp)                  # writes dict.txt / bpecodes
            cfg3, task3, model3 = make_whisper_model(pasero, tmp)
            golden_model_whisper(pasero, cfg3, task3, model3, out)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return
    golden_fbank(out)
    tmp = tempfile.mkdtemp()
    try:
        cfg, task, model = make_task_and_model(pasero, tmp)
        golden_batching(pasero, task, out)
        golden_batching_by(pasero, task, out)
        golden_model(pasero, cfg, task, model, out)
        golden_decoding(pasero, cfg, task, model, out)
        cfg2, task2, model2 = make_task_and_model(pasero, tmp)
        golden_training_curve(pasero, cfg2, task2, model2, out)
        cfg3, task3, model3 = make_whisper_model(pasero, tmp)
        golden_model_whisper(pasero, cfg3, task3, model3, out)
        golden_model_adapters(pasero, tmp, out)
        golden_training_adapters_by(pasero, tmp, out)
        golden_model_text(pasero, tmp, out)
        golden_preprocess_tags(pasero, tmp, out)
        golden_numpy_file(pasero, out)
        reference_option_names(pasero)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    main()
