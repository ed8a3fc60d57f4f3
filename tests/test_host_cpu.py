"""CPU-side checks of the host layer: batch contract bit-exact against the reference's golden vectors, config
surface, state_dict key layout, and that the C-ABI library loads and exports every declared symbol."""
import os
import re

import numpy as np
import pytest
import torch

from pasero_b200 import TrainingConfig, build_model
from pasero_b200.task import Dictionary, NumpyFile, SpeechTranslationTask, build_batches
from pasero_b200.modules import len_to_mask, mask_to_len
from tests.golden import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def tiny_training_cfg(**kw) -> TrainingConfig:
    t = synth.TINY
    opts = dict(encoder_layers=t['encoder_layers'], decoder_layers=t['decoder_layers'], embed_dim=t['embed_dim'],
                encoder_ffn_dim=t['ffn_dim'], decoder_ffn_dim=t['ffn_dim'], encoder_attention_heads=t['heads'],
                decoder_attention_heads=t['heads'], input_dim=80, conv_input_dim=80, conv_channels=t['conv_channels'],
                conv_kernel_sizes=[5, 5], encoder_max_len=3000, batch_size=64000, dropout=0.0, label_smoothing=0.1,
                dtype='bfloat16', lr=synth.TRAIN['lr'], warmup=synth.TRAIN['warmup'], max_steps=1000)
    opts.update(kw)
    return TrainingConfig.from_dict(opts)


def tiny_task(cfg=None) -> SpeechTranslationTask:
    cfg = cfg or tiny_training_cfg()
    d = Dictionary(tokens=[f'tok{i}' for i in range(synth.TINY['vocab_size'] - 4)])
    return SpeechTranslationTask(cfg, d)


def test_library_exports_every_declared_symbol():
    from pasero_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'pasero_b200.h')).read()
    declared = set(re.findall(r'\b(pb200_\w+)\s*\(', header))
    declared.discard('pb200_gemm_params')
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pb200_abi_version() == 1
    assert lib.pb200_fbank_num_frames(160000) == 998 and lib.pb200_fbank_num_frames(399) == 0


def test_no_cpu_fallback():
    """Compute entry points must fail loudly without a GPU rather than fall back."""
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pasero_b200 import frontend
    with pytest.raises(RuntimeError):
        frontend.fbank_batch(torch.zeros(1, 16000))
    from pasero_b200.training import Trainer
    cfg = tiny_training_cfg()
    with pytest.raises(RuntimeError):
        Trainer(cfg, tiny_task(cfg), build_model(cfg, tiny_task(cfg)))


def test_dictionary_ids():
    d = Dictionary(os.path.join(GOLD, '..', '..', 'tests', 'golden', 'dict_small.txt')) if False else \
        Dictionary(tokens=['a', 'b', 'c'])
    assert (d.padding_idx, d.eos_idx, d.unk_idx, d.bos_idx, d.sep_idx) == (1, 2, 3, 2, 0)
    assert len(d) == 7
    assert d.to_indices(['a', 'zzz', 'c'], append_eos=True, prepend_bos=True).tolist() == [2, 4, 3, 6, 2]
    assert d.to_indices(['a', 'b', 'c'], max_len=3, append_eos=True, prepend_bos=True).tolist() == [2, 4, 2]


def test_build_batches_docstring_vectors():
    sizes = [1000, 2000, 4000, 500, 1000, 1000]
    assert build_batches([0, 1, 2, 3, 4, 5], sizes.__getitem__, 4000, 1, 2) == [[0, 1], [2], [3, 4], [5]]
    assert mask_to_len(torch.tensor([[0, 1, 0, 1, 0], [1, 1, 0, 0, 0]]).bool()).tolist() == [4, 2]
    assert len_to_mask(torch.tensor([0, 3, 1])).long().tolist() == [[1, 1, 1], [0, 0, 0], [0, 1, 1]]


@pytest.mark.parametrize('bs', [4000, 40000])
@pytest.mark.parametrize('mult', [None, 8])
@pytest.mark.parametrize('lines', [None, 50])
def test_bucketing_bit_exact_with_reference(bs, mult, lines):
    g = np.load(os.path.join(GOLD, 'batching.npz'))
    cfg = tiny_training_cfg(batch_size_multiple=mult, lines_per_batch=lines)
    task = tiny_task(cfg)
    np.random.seed(synth.BATCHING_SEED)
    batches = task.buffered_batching(list(synth.batching_samples()), batch_size=bs)
    key = f'bs{bs}_m{mult or 0}_l{lines or 0}'
    assert np.array_equal(np.array([s['index'] for b in batches for s in b]), g[key + '_flat'])
    assert np.array_equal(np.array([len(b) for b in batches]), g[key + '_sizes'])


def test_batch_by_language_bit_exact_with_reference():
    """`batch_by` (what `*_adapters_by` models ask the batcher for, cli/train.py:275): batches of one language pair each, in the
    reference's order; the collated batch carries the shared meta entries (the adapters are selected from them) and only those."""
    g = np.load(os.path.join(GOLD, 'batching_by.npz'))
    task = tiny_task()
    np.random.seed(synth.BATCHING_SEED)
    batches = task.buffered_batching(list(synth.batching_samples_by_lang()), batch_by=('source_lang', 'target_lang'), batch_size=4000)
    assert np.array_equal(np.array([s['meta']['uid'] for b in batches for s in b]), g['flat'])
    assert np.array_equal(np.array([s['index'] for b in batches for s in b]), g['index'])
    assert np.array_equal(np.array([len(b) for b in batches]), g['sizes'])
    for b in batches[:20]:
        col = task.collate(b, padding_idx=task.padding_idx, dtype=torch.float32)
        shared = {k: v for k, v in b[0]['meta'].items() if all(s['meta'].get(k) == v for s in b)}
        assert col['meta'] == shared and {'source_lang', 'target_lang'} <= set(shared) and (len(b) == 1 or 'uid' not in shared)


def test_collate_bit_exact_with_reference():
    g = np.load(os.path.join(GOLD, 'batching.npz'))
    task = tiny_task()
    np.random.seed(synth.BATCHING_SEED)
    small = task.build_batches(list(synth.batching_samples()[:64]), shuffle=True, batch_size=4000)[3]
    col = task.collate(small, padding_idx=task.padding_idx, dtype=torch.float32)
    for k in ('decoder_input', 'prompt_mask', 'prompt_length', 'indices', 'encoder_input_length'):
        assert np.array_equal(col[k].numpy(), g['collate_' + k]), k
    assert np.array_equal(col['encoder_input'].double().sum(dim=(1, 2)).numpy(), g['collate_encoder_input_sum'])
    # ... and element-wise: row i is sample i's feature matrix followed by zeros up to the batch's longest utterance
    # (utils.tokens_as_tensor, pasero/utils.py:709-736), in the order the golden `indices` fix
    assert col['encoder_input'].shape == (len(small), max(len(s['encoder_input']) for s in small), 80)
    for i, sample in enumerate(small):
        n = len(sample['encoder_input'])
        assert col['indices'][i] == sample['index'] and col['encoder_input_length'][i] == n
        assert np.array_equal(col['encoder_input'][i, :n].numpy(), np.asarray(sample['encoder_input'], dtype=np.float32))
        assert not col['encoder_input'][i, n:].any()
    assert task.collate([], 1, torch.float32) is None


def test_preprocess_and_skip():
    cfg = tiny_training_cfg(max_source_len=100, max_target_len=6)
    task = tiny_task(cfg)
    feats = np.zeros((50, 80), dtype=np.float16)
    s = task.preprocess({'source': feats, 'target': 'tok1 tok2 nope', 'meta': {}}, append_eos=True)
    assert s['decoder_input'].tolist() == [2, 5, 6, 3, 2] and s['prompt_mask'].tolist() == [True, False, False, False, False]
    assert task.preprocess({'source': np.zeros((101, 80), np.float16), 'target': 'tok1', 'meta': {}}, append_eos=True) == {}
    assert task.preprocess({'source': feats, 'target': 'tok1 ' * 6, 'meta': {}}, append_eos=True) == {}
    t = task.preprocess({'source': np.zeros((300, 80), np.float16), 'target': 'tok1 ' * 9, 'meta': {}}, truncate=True, append_eos=True)
    assert len(t['encoder_input']) == 100 and len(t['decoder_input']) == 6 and t['decoder_input'][-1] == 2


def test_preprocess_tags_bit_exact_with_reference():
    """Language-code / domain tags (`lang_code: True` of the TED-top20 recipe, `source_lang_code`, `target_lang_code`,
    `domain_tag`, configured and per-line tags, lines that already carry tags) for the text and the speech task:
    encoder ids, decoder ids and prompt masks equal the arrays the reference's `preprocess` produced
    (tests/golden/make_golden.py: golden_preprocess_tags). Subword splitting stays with the reference's tokenizer: the
    fixture carries the BPE-split lines."""
    import json
    with open(os.path.join(GOLD, 'preprocess_tags.json')) as f:
        g = json.load(f)
    d = Dictionary(tokens=g['dict_tokens'])
    tasks = {}
    n_checked = 0
    for c in g['cases']:
        key = (c['task'], c['config'])
        if key not in tasks:
            opts = {}
            flags = list(c['flags'])
            while flags:
                name = flags.pop(0)[2:].replace('-', '_')
                vals = []
                while flags and not flags[0].startswith('--'):
                    vals.append(flags.pop(0))
                opts[name] = (True if not vals else (int(vals[0]) if vals[0].isdigit() else vals))
            kw = {k: opts.pop(k) for k in ('max_source_len', 'max_target_len') if k in opts}
            if c['task'] == 'translation':
                kw.update(encoder_max_len=256, input_dim=None, conv_input_dim=None)
                cfg = tiny_training_cfg(task='translation', **kw, **opts)
                task = SpeechTranslationTask.text(cfg, d)
            else:
                cfg = tiny_training_cfg(**kw, **opts)
                task = SpeechTranslationTask(cfg, d)
            assert (task.max_source_len, task.max_target_len) == (c['max_source_len'], c['max_target_len'])
            tasks[key] = task
        task = tasks[key]
        if c['task'] == 'translation':
            source = c['source_bpe']
        else:
            source = np.zeros((c['frames'], 80), dtype=np.float16)
        out = task.preprocess({'source': source, 'target': c['target_bpe'], 'meta': dict(c['meta'])},
                              truncate=c['truncate'], tokenize=c['tokenize'], append_eos=True)
        assert (not out) == c['skipped'], c
        if out:
            if c['task'] == 'translation':
                assert out['encoder_input'].tolist() == c['encoder_input'], c
            else:
                assert len(out['encoder_input']) == c['encoder_len']
            assert out['decoder_input'].tolist() == c['decoder_input'], c
            assert out['prompt_mask'].astype(int).tolist() == c['prompt_mask'], c
            n_checked += 1
    assert n_checked > 200
    # every tag the configuration can emit must be in the dictionary (translation.py:343-367)
    cfg = tiny_training_cfg(lang_code=True, target_langs=['en', 'xx'])
    with pytest.raises(AssertionError, match='OOV'):
        SpeechTranslationTask(cfg, d).check_tags()
    SpeechTranslationTask(tiny_training_cfg(lang_code=True, domain_tag=True, target_lang='en', domain='ted'), d).check_tags()


def test_fbank_plan_offsets_match_oracle_frame_counts():
    """Host side of the frontend: frame counts (snip_edges, 25 ms / 10 ms at 16 kHz) and the concatenated-batch offsets
    of a ragged batch, including utterances shorter than one window (0 frames)."""
    from oracle import fbank as of
    from pasero_b200 import frontend
    lengths = [160000, 400, 399, 0, 16000, 480000, 561, 560, 559]
    plan = frontend.FbankPlan(lengths, 'cpu')
    want = [of.num_frames(n) for n in lengths]
    assert plan.frames == want == [998, 1, 0, 0, 98, 2998, 2, 2, 1]
    assert plan.sample_offsets.tolist() == np.concatenate([[0], np.cumsum(lengths)]).tolist()
    assert plan.frame_offsets.tolist() == np.concatenate([[0], np.cumsum(want)]).tolist()
    assert (plan.batch, plan.max_frames, plan.total_frames, plan.total_samples) == (9, 2998, sum(want), sum(lengths))


def test_numpy_file_roundtrip(tmp_path):
    rng = np.random.RandomState(0)
    feats = [rng.randn(n, 80).astype(np.float32) for n in (5, 17, 1)]
    f = NumpyFile.build(str(tmp_path / 'train.npy'), feats, dtype='float16')
    assert len(f) == 3 and f.get_positions()[1].tolist() == [5, 17, 1]
    for got, want in zip(f, feats):
        assert got.dtype == np.float16 and np.array_equal(got, want.astype(np.float16))
    f.seek(1)
    assert next(f).shape == (17, 80)


def test_numpy_file_written_by_the_reference():
    """tests/golden/features_ref.npy was written by the reference's NumpyFile.build (make_golden.py: 5 utterances in a
    7-slot header); features_ref_read.npz is what the reference's own reader returned for it. Our reader must return the
    same arrays, and our writer must produce the same bytes."""
    from tests.golden.make_golden import numpy_file_features
    want = np.load(os.path.join(GOLD, 'features_ref_read.npz'))
    f = NumpyFile(os.path.join(GOLD, 'features_ref.npy'))
    assert len(f) == 5 and f.dim == 80 and f.dtype == np.float16
    assert f.get_positions()[0].tolist() == [0, 1, 2, 3, 4] and np.array_equal(f.get_positions()[1], want['lengths'])
    rows = list(f)
    for i, r in enumerate(rows):
        assert r.dtype == np.float16 and np.array_equal(r, want[f'row{i}']) and np.array_equal(f[i], want[f'row{i}'])
    f.seek(2)
    assert np.array_equal(next(f), want['third']) and f.tell() == 3
    packed, offsets = f.pack([3, 0, 4])
    assert offsets.tolist() == [0, 40, 47, 50] and np.array_equal(packed[40:47], want['row0'])


def test_numpy_file_writer_is_byte_compatible(tmp_path):
    from tests.golden.make_golden import numpy_file_features
    path = str(tmp_path / 'ours.npy')
    NumpyFile.build(path, numpy_file_features(), dtype='float16', num_feats=7)
    assert open(path, 'rb').read() == open(os.path.join(GOLD, 'features_ref.npy'), 'rb').read()


def test_state_dict_keys_match_reference():
    import json
    gold = torch.load(os.path.join(GOLD, 'st_tiny_model.pt'), weights_only=False)
    cfg = tiny_training_cfg()
    model = build_model(cfg, tiny_task(cfg))
    sd = model.state_dict()
    assert sorted(sd.keys()) == gold['state_dict_keys']
    shapes = synth.tiny_param_shapes()
    for k, shape in shapes.items():
        assert tuple(sd[k].shape) == shape, k
    # loading a reference-format state dict fills the fused / tap-major storage
    ref_sd = synth.tiny_state_dict()
    missing = model.load_state_dict(ref_sd, strict=False)
    assert not missing.unexpected_keys
    sd2 = model.state_dict()
    for k, v in ref_sd.items():
        assert torch.equal(sd2[k].float(), v), k
    attn = model.encoder.layers[0].self_attn
    D = cfg.model_cfg.embed_dim
    assert torch.equal(attn.in_proj_weight[2 * D:].detach(), ref_sd['encoder.layers.0.self_attn.q_proj.weight'])
    layout = json.load(open(os.path.join(GOLD, 'checkpoint_layout.json')))
    assert sorted(layout['model_3.bin']['model']) == sorted(sd.keys())


def test_params_regex_matches_reference_names():
    """--train-params-regex / --freeze-params-regex act on the reference's parameter names (cli/train.py:230-242), also
    where this package stores several of them in one tensor."""
    from pasero_b200 import reference_param_names
    task = tiny_task()
    model = build_model(tiny_training_cfg(), task)
    names = reference_param_names(model)
    ref_keys = {k for refs in names.values() for k in refs}
    import json
    with open(os.path.join(GOLD, 'checkpoint_layout.json')) as f:
        layout = json.load(f)
    want = {k for k in layout['model_3.bin']['model'] if '_float_tensor' not in k}
    assert ref_keys == want, (sorted(ref_keys - want)[:5], sorted(want - ref_keys)[:5])
    fused = [refs for refs in names.values() if len(refs) == 3]
    assert fused and all({r.split('.')[-2] for r in refs} == {'k_proj', 'v_proj', 'q_proj'} for refs in fused)

    model = build_model(tiny_training_cfg(train_params_regex=r'.*(layer_norm|fc1|self_attn\.[kvq]_proj)\.'), task)
    for name, p in model.named_parameters():
        ref = names[name][0]
        assert p.requires_grad == bool(re.match(r'.*(layer_norm|fc1|self_attn\.[kvq]_proj)\.', ref)), (name, ref)
    assert any(p.requires_grad for p in model.parameters()) and not all(p.requires_grad for p in model.parameters())
    model = build_model(tiny_training_cfg(freeze_params_regex=r'encoder\.'), task)
    assert all(p.requires_grad != n.startswith('encoder.') for n, p in model.named_parameters())
    with pytest.raises(NotImplementedError):   # a regex that cuts a fused projection apart
        build_model(tiny_training_cfg(train_params_regex=r'.*q_proj'), task)


def test_config_surface():
    cfg = TrainingConfig.from_dict({'task': 'speech_translation', 'arch': 'transformer', 'encoder_layers': 12,
                                    'input_dim': 80, 'conv_input_dim': 80, 'conv_channels': 1024, 'conv_kernel_sizes': [5, 5],
                                    'encoder_max_len': 3000, 'dtype': 'bfloat16', 'batch_size': 64000, 'dict': 'de-en/dict.txt',
                                    'source_lang': 'de', 'metrics': [], 'max_steps': 200, 'valid_interval': 100})
    m = cfg.model_cfg
    assert (m.encoder_layers, m.decoder_layers, m.embed_dim, m.label_smoothing, m.decoder_max_len) == (12, 6, 512, 0.1, 256)
    assert m.conv_strides == (2, 2) and cfg.extra['dict'] == 'de-en/dict.txt' and cfg.max_source_len == 3000
    with pytest.raises(NotImplementedError):
        TrainingConfig.from_dict({'dtype': 'bfloat16', 'fsdp': True})
    with pytest.raises(NotImplementedError):
        TrainingConfig.from_dict({'dtype': 'float32'})
    # the IWSLT2023 recipe's precision options (examples/IWSLT2023/taq-fr-contrastive-1.yaml:37-38)
    f16 = TrainingConfig.from_dict({'dtype': 'float16', 'memory_efficient_fp16': False, 'flat_fp16': True})
    assert (f16.dtype, f16.memory_efficient_fp16, f16.flat_fp16) == ('float16', False, True)
    assert cfg.as_dict()['encoder_layers'] == 12


def test_config_argv_surface_and_unknown_options(tmp_path):
    """CLI > YAML > defaults like the reference (pasero/config.py:1711-1760); a key that neither this path nor the reference
    declares is an error (config.py:1748-1750), a reference option this path does not read is kept in `extra`."""
    import yaml
    path = tmp_path / 'cfg.yaml'
    yaml.safe_dump(dict(task='speech_translation', arch='transformer', encoder_layers=12, conv_kernel_sizes=[5, 5], input_dim=80,
                        conv_channels=1024, dtype='bfloat16', dict='de-en/dict.txt', metrics=[], dataset_type='simple'), open(path, 'w'))
    cfg = TrainingConfig.from_argv(['-c', str(path), '--decoder-layers', '3', '--conv-kernel-sizes', '3', '3', '--lr', '1e-4',
                                    '--dropout', '0.2'])
    assert (cfg.model_cfg.encoder_layers, cfg.model_cfg.decoder_layers, cfg.model_cfg.conv_kernel_sizes) == (12, 3, (3, 3))
    assert cfg.lr == 1e-4 and cfg.model_cfg.dropout == 0.2 and cfg.extra['dict'] == 'de-en/dict.txt'
    with pytest.raises(ValueError, match='encoder_layerz'):
        TrainingConfig.from_argv(['--encoder-layerz', '3'])
    with pytest.raises(ValueError, match='not_an_option'):
        TrainingConfig.from_dict({'not_an_option': 1})
    with pytest.raises(NotImplementedError):
        TrainingConfig.from_dict({'fsdp': True})


def test_roofline_work_accounting():
    """The algorithmic FLOPs / bytes bench.py divides the in-situ kernel times by (pasero_b200/kernels.py: _attn_flops,
    _ln_bytes, _ce_bytes) against the figures DESIGN.md section 4 states, at the config-2 shapes."""
    from pasero_b200 import kernels as K
    B, H, Lq, Lk = 64, 8, 250, 250
    q = torch.empty(B * Lq, H * 64)
    # forward: 4 B H Lq Lk 64 FLOPs; backward 2.5 x; a causal mask halves it
    assert K._attn_flops(1.0)(q, q, q, B, H, Lq, Lk, None, False, 0.125) == ('flop', 4.0 * B * H * Lq * Lk * 64)
    assert K._attn_flops(1.0)(q, q, q, B, H, 48, 48, None, True, 0.125) == ('flop', 2.0 * B * H * 48 * 48 * 64)
    bwd = K._attn_flops(2.5)(q, q, q, q, q, q, B, H, Lq, Lk, None, False, 0.125, q, q, q)
    assert bwd == ('flop', 10.0 * B * H * Lq * Lk * 64)
    x = torch.empty(16000, 512)
    assert K._ln_bytes(x) == ('byte', 4.0 * 16000 * 512 * 2)            # four [M, D] 16-bit streams
    logits, target = torch.empty(3072, 8032), torch.empty(64, 48, dtype=torch.int64)
    assert K._ce_bytes(logits, target, 1, 0.1, None) == ('byte', 2.0 * 3072 * 8032 * 2)     # read once, gradient written over it
    padded = torch.empty(3072, 8040)
    assert K._ce_bytes(padded, target, 1, 0.1, None, vocab=8032) == ('byte', 2.0 * 3072 * 8032 * 2)
    assert K._ce_bytes(padded, target, 1, 0.1, None, True, 1.0, None, 8032) == ('byte', 2.0 * 3072 * 8032 * 2)
