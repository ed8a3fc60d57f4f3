"""Decoding between training steps: incremental decoder forward with a key/value `state` on the sm_100a kernels,
and the beam search `Trainer.inference_step` runs on top of it.

  * `incremental_forward`  — `TransformerDecoder.forward(..., state=...)` of the reference (pasero/models/transformer.py:830-898,
    :1246-1289; pasero/models/modules.py:621-641): `decoder_input` holds only the newest tokens, `state['offset']` the number
    of positions already consumed, `state['dec_{i}_self_attn_key' / '_value']` the cached projections, shaped (B, S, H, 64) like
    the reference's so that a search loop can reorder them with `index_select(0, beams)`. Unlike the reference (which
    re-projects the encoder output at every step) the cross-attention keys / values are projected once and kept in the state
    (`dec_{i}_cross_kv`). Inference only: no autograd graph, kernels called directly.
  * `beam_search` / `search` — the reference's search (pasero/decoding.py:1225-1656, fairseq's algorithm): top 2K candidates of
    the K x V continuations per sentence and step; those among the first K that end in EOS are finished hypotheses (score = sum
    of log-probabilities); the first K that do not become the new beams; a sentence is done once it has K finished hypotheses;
    hypotheses are ranked by score / (length + 1) ** len_penalty. Prompt tokens (`decoder_input`: BOS + target tags) are forced.
    Written against the decoder call signature only, so it runs on this package's CUDA decoder and on any callable with the
    reference's interface (the parity tests drive it with the reference's own CPU decoder and with the oracle).
Scores are returned the reference's way: positive losses, lower is better."""
import math
from typing import Callable, Optional

import torch
from torch import Tensor

from . import kernels as K
from . import modules


# --------------------------------------------------------------------------------------------------
# incremental decoder forward (CUDA kernels, no autograd)
# --------------------------------------------------------------------------------------------------
def _ln(x: Tensor, residual: Optional[Tensor], ln, eps: float) -> Tensor:
    """LayerNorm(x + residual) (ln given) or x + residual (ln None), through the fused kernel with p = 0."""
    out, _, _, _ = K.dropout_add_ln_fwd(x, residual, ln.weight if ln is not None else None, ln.bias if ln is not None else None,
                                        eps, 0.0, 0, 0, need_sum=False)
    return out


def _sublayer(x: Tensor, ln, prenorm: bool, eps: float, fn: Callable[[Tensor], Tensor]) -> Tensor:
    """residual + fn(norm?(x)), normalised afterwards when post-norm (pasero/models/transformer.py:1056-1099)."""
    if prenorm:
        return _ln(fn(_ln(x, None, ln, eps)), x, None, eps)
    return _ln(fn(x), x, ln, eps)


@torch.no_grad()
def incremental_forward(decoder, encoder_out: Tensor, encoder_mask: Optional[Tensor], decoder_input: Tensor, state: dict,
                        encoder_lengths: Optional[Tensor] = None, meta: Optional[dict] = None):
    """Logits (B, T, V) of the T newest target positions; `state` is updated in place."""
    from .transformer import _select_adapters
    cfg = decoder.cfg
    _select_adapters(decoder.layers, cfg.decoder_adapters_by if cfg.adapters else (), meta)
    B, T = decoder_input.shape
    D, H = cfg.embed_dim, cfg.decoder_attention_heads
    eps = cfg.norm_eps
    offset = int(state.get('offset', 0))
    state['offset'] = offset + T
    table = decoder.embed_tokens.weight
    pos = decoder.embed_positions.table(T, offset)
    x = K.embed_fwd(decoder_input.reshape(-1).contiguous(), table, pos.detach().contiguous(), B * T, D, T, float(decoder.embed_scale),
                    0.0, 0, 0)
    if encoder_lengths is None:
        encoder_lengths = modules.MultiheadAttention.key_lengths_of(encoder_mask)
    S = encoder_out.shape[1]
    enc2 = encoder_out.reshape(B * S, D)
    for i, layer in enumerate(decoder.layers):
        sa, ca = layer.self_attn, layer.encoder_attn
        scale = 1.0 / math.sqrt(sa.head_dim) if sa.scaled else 1.0
        key = f'dec_{i}_self_attn'

        def self_attention(h, sa=sa, key=key):
            kvq = K.linear_fwd(h, sa.in_proj_weight, sa.in_proj_bias)                      # [B*T, 3D] (k | v | q)
            k_new, v_new = kvq[:, :D].reshape(B, T, H, 64), kvq[:, D:2 * D].reshape(B, T, H, 64)
            if key + '_key' in state:                                                       # modules.py:621-641
                k_all = torch.cat([state[key + '_key'], k_new], dim=1)
                v_all = torch.cat([state[key + '_value'], v_new], dim=1)
            else:
                k_all, v_all = k_new.contiguous(), v_new.contiguous()
            state[key + '_key'], state[key + '_value'] = k_all, v_all
            L = k_all.shape[1]
            q = kvq[:, 2 * D:]
            # causal over the cached prefix: query t (global position L - T + t) sees keys [0, L - T + t]
            o, _ = K.attention_fwd(q, k_all.view(B * L, D), v_all.view(B * L, D), B, H, T, L, None, T > 1, scale)
            return K.linear_fwd(o, sa.out_proj.weight, sa.out_proj.bias)

        def cross_attention(h, ca=ca, i=i):
            ckey = f'dec_{i}_cross_kv'
            if ckey not in state:      # projected once per sentence, reordered with the beams like every other state tensor
                state[ckey] = K.linear_fwd(enc2, ca.in_proj_weight[:2 * D], ca.in_proj_bias[:2 * D]).view(B, S, 2 * D)
            kv = state[ckey].view(B * S, 2 * D)
            q = K.linear_fwd(h, ca.in_proj_weight[2 * D:], ca.in_proj_bias[2 * D:])
            o, _ = K.attention_fwd(q, kv[:, :D], kv[:, D:], B, H, T, S, encoder_lengths, False, scale)
            return K.linear_fwd(o, ca.out_proj.weight, ca.out_proj.bias)

        def ffn(h, layer=layer):
            act = 1 if cfg.activation_fn == 'relu' else 2
            return K.linear_fwd(K.linear_fwd(h, layer.fc1.weight, layer.fc1.bias, act=act), layer.fc2.weight, layer.fc2.bias)

        x = _sublayer(x, layer.self_attn_layer_norm, layer.prenorm, eps, self_attention)
        x = _sublayer(x, layer.encoder_attn_layer_norm, layer.prenorm, eps, cross_attention)
        x = _sublayer(x, layer.final_layer_norm, layer.prenorm, eps, ffn)
        x = layer.apply_adapters(x.view(B, T, D)).reshape(B * T, D)
    if decoder.layer_norm is not None:
        x = _ln(x, None, decoder.layer_norm, eps)
    logits = K.linear_fwd(x, table, None)[:, :decoder.embed_tokens.vocab_size]
    return logits.view(B, T, -1), {}


def reorder_state(state: dict, beams: Tensor) -> None:
    """Decoder.reorder_state of the reference: every tensor of the incremental state follows its beam."""
    for k, v in state.items():
        if torch.is_tensor(v) and v.dim() > 0:
            state[k] = v.index_select(0, beams.to(v.device))


# --------------------------------------------------------------------------------------------------
# search
# --------------------------------------------------------------------------------------------------
def _attr(decoder, name: str, default=None):
    v = getattr(decoder, name, default)
    if v is None:
        raise AttributeError(f'the decoder passed to search() has no `{name}`')
    return v


@torch.no_grad()
def beam_search(decoder, encoder_out: Tensor, encoder_mask: Optional[Tensor], max_output_len: int, beam_size: int = 5,
                meta: Optional[dict] = None, *, decoder_input: Optional[Tensor] = None, len_penalty: float = 1.0,
                blacklist=(), return_scores: bool = False, **kwargs) -> list:
    """Returns, per sentence, its finished hypotheses best first: dicts with 'tokens' (LongTensor: the hypothesis after its
    first token, EOS included), 'score' (negative log-probability), 'normalized_score' and, with return_scores, 'pos_scores'."""
    bos, eos, pad = _attr(decoder, 'bos_idx'), _attr(decoder, 'eos_idx'), _attr(decoder, 'padding_idx')
    B = encoder_out.size(0)
    Kb = beam_size
    device = encoder_out.device
    meta = meta or {}
    if decoder_input is None:
        decoder_input = torch.full((B, 1), bos, dtype=torch.long, device=device)
    decoder_input = decoder_input.to(device)
    prompt_len = (decoder_input != pad).long().sum(dim=1)      # prompts are left-aligned (utils.mask_to_len)
    if not bool((prompt_len > 0).all()):
        raise ValueError('every sentence needs at least one prompt token (BOS)')
    start, max_prompt = int(prompt_len.min()), int(prompt_len.max())
    max_len = min(int(_attr(decoder, 'max_len')), max_prompt + max_output_len)
    tokens = torch.full((B * Kb, max_len), eos, dtype=torch.long, device=device)      # EOS doubles as the padding of `tokens`
    rep = decoder_input.repeat_interleave(Kb, dim=0)
    tokens[:, :rep.size(1)] = rep[:, :max_len]
    step_input = rep[:, :start]
    enc = encoder_out.repeat_interleave(Kb, dim=0)
    mask = encoder_mask.repeat_interleave(Kb, dim=0) if encoder_mask is not None else None
    hyps = [[] for _ in range(B)]
    done = torch.zeros(B, dtype=torch.bool, device=device)
    hyp_scores = None                     # (B * K,) cumulative log-probabilities of the live beams
    cum = torch.zeros(B * Kb, max_len, device=device) if return_scores else None
    rank = torch.arange(2 * Kb, device=device)
    row0 = (Kb * torch.arange(B, device=device)).unsqueeze(1)
    state = {}
    for step in range(start, max_len):
        logits, _ = decoder(enc, mask, step_input, state=state, meta=meta)
        logp = logits.float().log_softmax(dim=-1)                       # (B * K, T, V)
        scores = logp[:, -1]
        V = scores.size(-1)
        prompt_scores = None
        if logp.size(1) > 1:              # the common prompt prefix is scored in one call at the first step
            forced_prefix = step_input[:, 1:]
            prompt_scores = logp[:, :-1].gather(2, forced_prefix.unsqueeze(2)).squeeze(2) * (forced_prefix != pad)
            scores = scores + prompt_scores.sum(dim=1, keepdim=True)
        for t in blacklist:
            scores[:, t] = -math.inf
        if step < max_prompt:             # longer prompts: their next token is forced
            forced = tokens[:, step]
            rows = torch.arange(scores.size(0), device=device)
            keep = scores[rows, forced].clone()
            is_forced = (forced != pad) & (forced != eos)
            scores = scores.masked_fill(is_forced.unsqueeze(1), -math.inf)
            scores[rows, forced] = keep
        if hyp_scores is None:
            scores = scores[::Kb]         # all K copies are identical at the first step
        else:
            scores = scores + hyp_scores.unsqueeze(-1)
        top = scores.reshape(B, -1).topk(2 * Kb, dim=1)
        beam_of = top.indices.div(V, rounding_mode='floor')
        cand_beam = beam_of + row0                                      # rows of `tokens`
        cand_tok = top.indices.fmod(V)
        cand_score = top.values
        ends = cand_tok.eq(eos) & ~cand_score.isinf() & (step >= prompt_len).unsqueeze(1)
        # finished hypotheses: candidates ending in EOS among the first K, for sentences that still want some
        fin = ends[:, :Kb] & ~done.unsqueeze(1)
        if bool(fin.any()):
            for b, j in fin.nonzero().tolist():
                if len(hyps[b]) >= Kb:
                    continue
                row = int(cand_beam[b, j])
                hyp = {'tokens': tokens[row, 1:step + 1].clone(), 'score': -cand_score[b, j]}
                if return_scores:
                    c = cum[row, :step].clone()
                    c[step - 1] = cand_score[b, j]
                    c[1:] = c[1:] - c[:-1].clone()
                    hyp['pos_scores'] = -c
                hyps[b].append(hyp)
            done = torch.tensor([len(h) >= Kb for h in hyps], device=device)
            if bool(done.all()):
                break
        # the new beams: the first K candidates that do not end here
        _, pick = torch.topk(rank + ends * (2 * Kb), k=Kb, dim=1, largest=False)
        new_rows = cand_beam.gather(1, pick).view(-1)
        new_tok = cand_tok.gather(1, pick).view(-1)
        hyp_scores = cand_score.gather(1, pick).view(-1)
        tokens[:, :step] = tokens[:, :step].index_select(0, new_rows)
        tokens[:, step] = new_tok
        if return_scores:
            if prompt_scores is None:
                cum[:, :step - 1] = cum[:, :step - 1].index_select(0, new_rows)
            else:
                cum[:, :step - 1] = prompt_scores.cumsum(dim=1).index_select(0, new_rows)
            cum[:, step - 1] = hyp_scores
        step_input = new_tok.unsqueeze(1)
        enc = enc.index_select(0, new_rows)
        mask = mask.index_select(0, new_rows) if mask is not None else None
        reorder_state(state, new_rows)
    out = []
    for b in range(B):
        nbest = hyps[b]
        if not nbest:        # the length limit came before any EOS: the best live beam, as it stands
            hyp = {'tokens': tokens[b * Kb, 1:].clone(), 'score': -hyp_scores[b * Kb]}
            if return_scores:
                c = cum[b * Kb].clone()
                c[1:] = c[1:] - c[:-1].clone()
                hyp['pos_scores'] = -c[:-1]
            nbest = [hyp]
        for hyp in nbest:
            hyp['normalized_score'] = hyp['score'] / (len(hyp['tokens']) + 1) ** len_penalty
        out.append(sorted(nbest, key=lambda h: float(h['normalized_score'])))
    return out


def search(decoder, encoder_out: Tensor, encoder_mask: Optional[Tensor], *, max_output_len: int = 100, beam_size: int = 5,
           meta: Optional[dict] = None, decoder_input: Optional[Tensor] = None, **opts) -> list:
    """`decoding.search` of the reference (pasero/decoding.py:907-917) for the deterministic case: beam search, greedy being
    beam_size 1. Extra batch entries (`encoder_input`, `indices`, ...) are accepted and ignored, as the reference's do."""
    if opts.get('sampling'):
        raise NotImplementedError('sampling is outside the validation path (Trainer.inference_step decodes with beam search)')
    accepted = {k: opts[k] for k in ('len_penalty', 'blacklist', 'return_scores') if k in opts}
    return beam_search(decoder, encoder_out, encoder_mask, max_output_len, beam_size, meta, decoder_input=decoder_input, **accepted)
