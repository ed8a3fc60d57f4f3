"""The drop-in boundary, exercised against the UNMODIFIED reference where it is available (the build container:
/root/reference; skipped on the GPU box, which does not have it). This does exactly what INTEGRATION.md §2 tells a Pasero
maintainer to do — register `pasero_b200.transformer.Transformer` under a new architecture name in the reference's own
registry — then builds the model the way `cli/train.py:220` does, from the reference's own config and task objects, and
checks that checkpoints are interchangeable."""
import os
import shutil

import pytest
import torch

from tests.golden import synth

REF = os.environ.get('PASERO_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'pasero')), reason='the reference tree is not on this box')


@pytest.fixture(scope='module')
def pasero():
    from tests.golden import make_golden as mg
    ref = mg.import_reference()
    from pasero.config import register_model, register_model_config, TransformerConfig, AdapterTransformerConfig
    import pasero_b200.transformer as b200

    # ---- INTEGRATION.md §2, verbatim -------------------------------------------------------------
    @register_model_config('transformer_b200')
    class TransformerB200Config(TransformerConfig):
        pass

    @register_model('transformer_b200')
    class TransformerB200(b200.Transformer):
        """--arch transformer_b200 --task speech_translation --dtype bfloat16"""

    @register_model_config('adapter_transformer_b200')
    class AdapterTransformerB200Config(AdapterTransformerConfig):
        pass

    @register_model('adapter_transformer_b200')
    class AdapterTransformerB200(b200.Transformer):
        """--arch adapter_transformer_b200: adapters are selected by the config class, as in models/adapters.py:62-119"""
    # -----------------------------------------------------------------------------------------------
    return ref


def reference_args(tmp, arch, extra=()):
    t = synth.TINY
    with open(os.path.join(tmp, 'dict.txt'), 'w') as f:
        for i in range(t['vocab_size'] - 4):
            f.write(f'tok{i} 1\n')
    shutil.copy(os.path.join(REF, 'examples/TED/de-en/bpecodes'), os.path.join(tmp, 'bpecodes'))
    return ['--task', 'speech_translation', '--arch', arch,
            '--encoder-layers', str(t['encoder_layers']), '--decoder-layers', str(t['decoder_layers']),
            '--embed-dim', str(t['embed_dim']), '--encoder-ffn-dim', str(t['ffn_dim']), '--decoder-ffn-dim', str(t['ffn_dim']),
            '--encoder-attention-heads', str(t['heads']), '--decoder-attention-heads', str(t['heads']),
            '--input-dim', '80', '--conv-input-dim', '80', '--conv-channels', str(t['conv_channels']),
            '--conv-kernel-sizes', '5', '5', '--data-dir', tmp, '--dict', 'dict.txt', '--tokenizer-path', 'bpecodes',
            '--source-lang', 'de', '--target-lang', 'en', '--model-dir', os.path.join(tmp, 'model_' + arch),
            '--encoder-max-len', '3000', '--batch-size', '64000', '--dropout', '0.1', '--label-smoothing', '0.1',
            '--metrics', '--early-stopping-metric', 'nll_loss', '--max-steps', '1000', '--valid-interval', '1000',
            '--dataset-type', 'simple', '--dtype', 'float32', '--lr', '0.0002', '--warmup', '50', '--seed', '1234'] + list(extra)


def build_through_registry(tmp, arch, extra=()):
    """cli/train.py:172-220: config -> task -> `get_architecture(cfg.model_cfg)(cfg.model_cfg, cfg, task)`."""
    from pasero.config import TrainingConfig, get_task_class, get_architecture
    cfg = TrainingConfig(*reference_args(tmp, arch, extra))
    task = get_task_class(cfg.task)(cfg.data_dir, cfg.task_cfg)
    task.setup_for_model(cfg.model_cfg)
    cfg.dp_size, cfg.dp_rank, cfg.dp_local_size, cfg.dp_local_rank, cfg.start_rank = 1, 0, 1, 0, 0   # utils.setup_distributed
    torch.manual_seed(1234)
    model = get_architecture(cfg.model_cfg)(cfg.model_cfg, cfg, task)
    return cfg, task, model


def shapes(sd):
    return {k: tuple(v.shape) for k, v in sd.items()}


@pytest.mark.parametrize('arch', ['transformer', 'adapter_transformer'])
def test_registry_binding_builds_and_exchanges_checkpoints(pasero, tmp_path, arch):
    import pasero_b200.transformer as b200
    tmp = str(tmp_path)
    cfg_r, task_r, ref_model = build_through_registry(tmp, arch)
    cfg_o, task_o, our_model = build_through_registry(tmp, arch + '_b200')
    assert isinstance(our_model, b200.Transformer) and type(our_model).__name__.endswith('B200')
    assert our_model.padding_idx == ref_model.padding_idx
    assert our_model.total_param_count == ref_model.total_param_count
    # same keys and shapes as the reference's model
    assert shapes(our_model.state_dict()) == shapes(ref_model.state_dict())
    assert bool(our_model.cfg.adapters) == (arch == 'adapter_transformer')
    # the same parameters are trainable (adapter_transformer freezes the base, models/adapters.py:115-119)
    from pasero_b200 import reference_param_names
    names = reference_param_names(our_model)
    ours_trainable = {r for n, p in our_model.named_parameters() if p.requires_grad for r in names[n]}
    ref_trainable = {n for n, p in ref_model.named_parameters() if p.requires_grad}
    assert ours_trainable == ref_trainable
    # a checkpoint WRITTEN BY THE REFERENCE's Trainer (training.py:918-939) loads into our model ...
    from pasero.training import Trainer
    from pasero import utils
    os.makedirs(cfg_r.model_dir, exist_ok=True)
    trainer = Trainer(cfg_r, task_r, ref_model, utils.Metrics(), None)
    batches = synth.train_batches()
    for step in range(3):
        trainer.train_step(iter([{k: (v.clone() if torch.is_tensor(v) else dict(v)) for k, v in batches[step].items()}]))
    trainer.save()
    ckpt = torch.load(os.path.join(cfg_r.model_dir, 'model_3.bin'), map_location='cpu', weights_only=False)
    result = our_model.load_state_dict(ckpt['model'], strict=True)
    assert not result.missing_keys and not result.unexpected_keys
    ours, want = our_model.state_dict(), ref_model.state_dict()
    assert all(torch.equal(ours[k].float(), want[k].float()) for k in want if not k.endswith('_float_tensor'))
    # ... and ours loads back into the reference's
    result = ref_model.load_state_dict(our_model.state_dict(), strict=True)
    assert not result.missing_keys and not result.unexpected_keys


def test_registry_binding_per_language_adapters(pasero, tmp_path):
    """`--encoder-adapters-by source_lang --decoder-adapters-by target_lang` through the reference's registry and its own task
    object (whose corpora define the languages, models/adapters.py:74-105): the same adapter sets under the same keys, the same
    `batch_by` / `find_unused_parameters` the reference's `AdapterTransformer` hands to cli/train.py:275 and to DDP, and the
    adapters of the batch's language pair selected from `meta` (adapters.py:186-197, 219-229)."""
    from pasero_b200.training import adapter_group
    from pasero_b200.transformer import _select_adapters
    tmp = str(tmp_path)
    extra = ('--source-langs', 'de', 'it', '--target-langs', 'en', 'fr', '--encoder-adapters-by', 'source_lang',
             '--decoder-adapters-by', 'target_lang', '--encoder-adapter-dim', '16', '--decoder-adapter-dim', '16')
    cfg_r, task_r, ref_model = build_through_registry(tmp, 'adapter_transformer', extra)
    cfg_o, task_o, our_model = build_through_registry(tmp, 'adapter_transformer_b200', extra)
    assert shapes(our_model.state_dict()) == shapes(ref_model.state_dict())
    assert our_model.find_unused_parameters is True and ref_model.find_unused_parameters is True
    assert sorted(our_model.batch_by) == sorted(ref_model.batch_by) == ['source_lang', 'target_lang']
    trainable = {n for n, p in our_model.named_parameters() if p.requires_grad}
    assert trainable == {n for n, p in ref_model.named_parameters() if p.requires_grad}
    assert {adapter_group(n) for n in trainable} == {('encoder', 'lang:de'), ('encoder', 'lang:it'), ('decoder', 'lang:en'),
                                                     ('decoder', 'lang:fr')}
    meta = {'source_lang': 'it', 'target_lang': 'fr'}
    _select_adapters(our_model.encoder.layers, our_model.cfg.encoder_adapters_by, meta)
    _select_adapters(our_model.decoder.layers, our_model.cfg.decoder_adapters_by, meta)
    assert all(tuple(layer.adapters_in_use) == ('lang:it',) for layer in our_model.encoder.layers)
    assert all(tuple(layer.adapters_in_use) == ('lang:fr',) for layer in our_model.decoder.layers)


def test_reference_training_config_is_adopted(pasero, tmp_path):
    """`Trainer(cfg, ...)` receives the reference's own TrainingConfig at cli/train.py:287."""
    from pasero.config import TrainingConfig as RefTrainingConfig
    from pasero_b200 import TrainingConfig
    ref_cfg = RefTrainingConfig(*reference_args(str(tmp_path), 'transformer_b200', ('--dtype', 'bfloat16', '--clip-norm', '0.5')))
    cfg = TrainingConfig.adopt(ref_cfg)
    assert (cfg.lr, cfg.warmup, cfg.clip_norm, cfg.max_steps, cfg.seed) == (0.0002, 50, 0.5, 1000, 1234)
    assert tuple(cfg.adam_betas) == tuple(ref_cfg.adam_betas) and cfg.virtual_dp_size == ref_cfg.virtual_dp_size
    assert cfg.model_cfg.conv_kernel_sizes == (5, 5) and cfg.model_cfg.label_smoothing == 0.1
    assert cfg.as_dict() == ref_cfg.as_dict()
    with pytest.raises(NotImplementedError):
        TrainingConfig.adopt(RefTrainingConfig(*reference_args(str(tmp_path), 'transformer_b200', ('--dtype', 'bfloat16', '--tp-size', '2'))))
