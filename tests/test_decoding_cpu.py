"""pasero_b200.decoding.beam_search against the reference's pasero/decoding.py:1225-1656.

The search is written against a decoder's call signature only, so it is pinned on the CPU:
  * golden vectors (tests/golden/beam_search.json, written by `make_golden.py decode` from the UNMODIFIED reference): every
    finished hypothesis — tokens, score, normalised score, per-position scores — of the reference's beam search (a) on a
    deterministic stand-in decoder whose hypotheses end at different steps (synth.MockDecoder: early EOS, sentences that finish
    before the others, forced prompts, the length limit) and (b) on the tiny ST model, here decoded through the oracle;
  * where the reference tree is present, the two implementations are also run side by side on fresh seeds."""
import json
import os

import pytest
import torch

from pasero_b200 import decoding
from tests.golden import synth

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'beam_search.json')))
CASES = {c['name']: c for c in GOLD['cases']}


def prompts(batch_size: int, prompt: bool) -> torch.Tensor:
    rows = [[synth.BOS, 7, 9], [synth.BOS], [synth.BOS, 11], [synth.BOS, 7, 9]] if prompt else [[synth.BOS]]
    rows = [rows[i % len(rows)] for i in range(batch_size)]
    width = max(len(r) for r in rows)
    return torch.tensor([r + [synth.PAD] * (width - len(r)) for r in rows], dtype=torch.long)


def check(ours: list, want: list, tol: float) -> None:
    assert len(ours) == len(want)
    for nbest, wbest in zip(ours, want):
        assert [h['tokens'].tolist() for h in nbest] == [w['tokens'] for w in wbest]
        for h, w in zip(nbest, wbest):
            assert abs(float(h['score']) - w['score']) <= tol * max(1.0, abs(w['score']))
            assert abs(float(h['normalized_score']) - w['normalized_score']) <= tol * max(1.0, abs(w['normalized_score']))
            assert torch.allclose(h['pos_scores'].float(), torch.tensor(w['pos_scores']), atol=tol * 10, rtol=tol)


@pytest.mark.parametrize('seed', range(4))
@pytest.mark.parametrize('case', sorted(CASES))
def test_search_matches_reference_golden_on_stand_in_decoder(case, seed):
    c = CASES[case]
    dec = synth.MockDecoder(seed, max_len=32, eos_slope=0.9 if seed < 3 else 0.05)
    enc, mask = synth.mock_encoder_out(5, seed)
    out = decoding.beam_search(dec, enc, mask, c['max_output_len'], c['beam_size'], {}, decoder_input=prompts(5, c['prompt']),
                               return_scores=True, len_penalty=c['len_penalty'])
    check(out, GOLD['mock'][f'{case}_seed{seed}'], 1e-6)


class OracleDecoder:
    """The oracle's teacher-forced decoder behind the incremental interface: the prefix lives in the state, every call
    recomputes it (CPU, fp32)."""

    def __init__(self, cfg, sd):
        self.cfg, self.sd = cfg, sd
        self.padding_idx, self.eos_idx, self.bos_idx, self.max_len = synth.PAD, synth.EOS, synth.BOS, cfg.decoder_max_len

    def __call__(self, enc, mask, decoder_input, state=None, meta=None, **kwargs):
        from oracle import model as om
        prefix = decoder_input if 'prefix' not in state else torch.cat([state['prefix'], decoder_input], dim=1)
        state['prefix'] = prefix
        return om.decode(self.cfg, self.sd, enc, mask, prefix)[:, -decoder_input.size(1):], {}


@pytest.mark.parametrize('case', sorted(CASES))
def test_search_on_tiny_model_matches_reference_golden(case):
    from oracle import model as om
    from tests.test_oracle_cpu import tiny_cfg
    c = CASES[case]
    sd = synth.tiny_state_dict()
    batch = synth.model_batch()
    with torch.no_grad():
        enc, mask = om.encode(tiny_cfg(), sd, batch['encoder_input'], batch['encoder_input_length'])
        out = decoding.beam_search(OracleDecoder(tiny_cfg(), sd), enc, mask, c['max_output_len'], c['beam_size'], {},
                                   decoder_input=prompts(enc.size(0), c['prompt']), return_scores=True, len_penalty=c['len_penalty'])
    check(out, GOLD['results'][case], 1e-4)


@pytest.mark.skipif(not os.path.isdir('/root/reference/pasero'), reason='the reference tree is not on this box')
@pytest.mark.parametrize('seed', [11, 12, 13])
def test_search_side_by_side_with_the_reference(seed):
    from tests.golden import make_golden as mg
    mg.import_reference()
    from pasero import decoding as ref
    dec = synth.MockDecoder(seed, max_len=20, eos_slope=0.6)
    enc, mask = synth.mock_encoder_out(7, seed)
    for beam, maxlen, prompt, lp in [(5, 15, False, 1.0), (2, 6, True, 0.5), (8, 18, True, 1.3), (1, 10, False, 1.0)]:
        dec_in = prompts(7, prompt)
        with torch.no_grad():
            want = ref.beam_search(dec, enc, mask, maxlen, beam, {}, decoder_input=dec_in.clone(), return_scores=True, len_penalty=lp)
        ours = decoding.beam_search(dec, enc, mask, maxlen, beam, {}, decoder_input=dec_in.clone(), return_scores=True, len_penalty=lp)
        want = [[{'tokens': h['tokens'].tolist(), 'score': float(h['score']), 'normalized_score': float(h['normalized_score']),
                  'pos_scores': [float(x) for x in h['pos_scores']]} for h in nbest] for nbest in want]
        check(ours, want, 1e-6)
