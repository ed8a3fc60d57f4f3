"""Incremental decoding on the B200: `TransformerDecoder.forward(..., state=...)` on the sm_100a kernels against the
teacher-forced decoder, the beam search on top of it against the reference's golden hypotheses, and Trainer.inference_step.
(The search algorithm itself is pinned bit-exactly on the CPU, tests/test_decoding_cpu.py.)"""
import json
import os

import pytest
import torch

from tests.golden import synth
from tests.test_decoding_cpu import CASES, GOLD, prompts
from tests.test_parity_gpu import BF16, build, rel_to_max, to_cuda, wsp_training_cfg
from tests.test_host_cpu import tiny_task

pytestmark = pytest.mark.gpu


def _models():
    from pasero_b200 import build_model
    cfg, task, model = build()
    yield 'post-norm relu sinusoidal', model
    wcfg = wsp_training_cfg()
    wmodel = build_model(wcfg, tiny_task(wcfg))
    wmodel.load_state_dict(synth.wsp_state_dict(), strict=False)
    yield 'pre-norm gelu learned positions', wmodel


def test_incremental_decoder_matches_teacher_forcing():
    """Prompt of three tokens in one call, then one token per call: the logits of every position must be the teacher-forced
    ones (same kernels, same bf16 rounding points except the attention's tile shapes: 2 % of the largest logit)."""
    batch = synth.model_batch()
    for name, model in _models():
        model.to(BF16).cuda().eval()
        cb = to_cuda(batch)
        with torch.no_grad():
            enc_out, enc_mask, _ = model.encoder(cb['encoder_input'], cb['encoder_input_length'])
            dec_in = cb['decoder_input'][:, :-1].clamp(min=0)
            dec_in = torch.where(dec_in == model.padding_idx, torch.full_like(dec_in, 5), dec_in)   # no padding inside the prefix
            full, _ = model.decoder(enc_out, enc_mask, dec_in)
            state, parts = {}, []
            logits, _ = model.decoder(enc_out, enc_mask, dec_in[:, :3], state=state)
            parts.append(logits)
            for t in range(3, dec_in.size(1)):
                logits, _ = model.decoder(enc_out, enc_mask, dec_in[:, t:t + 1], state=state)
                parts.append(logits)
            inc = torch.cat(parts, dim=1)
        assert state['offset'] == dec_in.size(1) and state['dec_0_self_attn_key'].shape[:2] == (dec_in.size(0), dec_in.size(1))
        assert inc.shape == full.shape
        assert rel_to_max(inc, full) < 2e-2, (name, rel_to_max(inc, full))


@pytest.mark.parametrize('case', ['beam5', 'greedy', 'beam3_prompt', 'beam4_short'])
def test_beam_search_on_cuda_decoder_vs_reference_golden(case):
    """The reference's hypotheses (fp32, CPU) against the search on the bf16 CUDA decoder: same best hypothesis, score within 2 %."""
    from pasero_b200 import decoding
    c = CASES[case]
    cfg, task, model = build()
    model.to(BF16).cuda().eval()
    cb = to_cuda(synth.model_batch())
    with torch.no_grad():
        enc_out, enc_mask, _ = model.encoder(cb['encoder_input'], cb['encoder_input_length'])
        out = decoding.search(model.decoder, enc_out, enc_mask, max_output_len=c['max_output_len'], beam_size=c['beam_size'],
                              decoder_input=prompts(enc_out.size(0), c['prompt']).cuda(), len_penalty=c['len_penalty'])
    for nbest, want in zip(out, GOLD['results'][case]):
        assert nbest[0]['tokens'].tolist() == want[0]['tokens']
        assert abs(float(nbest[0]['score']) - want[0]['score']) / want[0]['score'] < 2e-2


def test_trainer_inference_step():
    from pasero_b200.training import Trainer
    cfg, task, model = build()
    trainer = Trainer(cfg, task, model)
    batch = dict(synth.model_batch())
    B = batch['decoder_input'].size(0)
    batch['prompt_length'] = torch.ones(B, dtype=torch.long)
    batch['indices'] = torch.arange(40, 40 + B)
    hyps = trainer.inference_step(batch, beam_size=5, max_output_len=12)
    assert [h['idx'] for h in hyps] == list(range(40, 40 + B)) and trainer.model.training
    for h, want in zip(hyps, GOLD['results']['beam5']):
        assert h['tokens'].tolist() == want[0]['tokens'] and abs(h['score'] - want[0]['score']) / want[0]['score'] < 2e-2
    assert trainer.inference_step(None) == []
    # and a training step still works afterwards (the decoder's state path leaves no autograd / dropout state behind)
    logs = trainer.train_step(iter([synth.model_batch()]))
    assert logs['num_tokens'] > 0
