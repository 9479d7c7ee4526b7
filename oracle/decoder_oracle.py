"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not a product path, never a fallback.

Restates, as plain vectorised fp32 tensor arithmetic on the host, the algorithm of the reference's
caption-decoder hot path (``DecoderTransformer.forward`` / ``.predict`` and the train-step loss) for the
three variants:

    G = geo-aware/models.py, K = knowledge-aware/models.py, N = news-knowledge-aware/models.py

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this file.  Nothing under ``image-captioning-with-external-knowledge_b200/`` imports it.

Pinning: the reference holds no tests or golden vectors for this path, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the three unmodified reference
``models.py`` files in the build container, runs them on seeded synthetic inputs with shared weights and commits
the results as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against those vectors.
The arithmetic primitives themselves (multi-head attention, LayerNorm, Linear, Embedding, CrossEntropyLoss) live
in third-party PyTorch (reference pins torch==1.9.0, geo-aware/requirements.txt:7); they are restated here from
their published definitions (post-LN Transformer layers, scaled-dot-product attention with 1/sqrt(head_dim),
biased-variance LayerNorm eps=1e-5) and anchored on the reference's call sites cited per function.

Tensor convention here is batch-major ``(B, S, D)``; the reference is sequence-major ``(S, B, D)``.  Parameters are
addressed by the reference's own ``state_dict`` keys.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


@dataclass
class Spec:
    """Static description of one reference variant (sizes hard-coded in the reference constructors)."""

    variant: str  # "G" | "K" | "N"
    vocab_size: int
    emb_dim: int = 300
    num_heads: int = 10
    num_layers: int = 3
    pad: int = 0  # word_map["<pad>"]  (G/create_input_files.py:311-315)
    start: int = 0
    end: int = 0

    @property
    def has_facts(self) -> bool:
        return self.variant in ("K", "N")

    @property
    def num_feat(self) -> int:
        # leading hand-made feature columns before the type embedding:
        # G/models.py:102 (4), K/models.py:131 (6), N/models.py:126 (5)
        return {"G": 4, "K": 6, "N": 5}[self.variant]

    @property
    def num_predicates(self) -> int:
        # K/models.py:329, N/models.py:312
        return {"G": 0, "K": 3000, "N": 3500}[self.variant]


# --------------------------------------------------------------------------------------------------------------
# dropout hook: the oracle is deterministic; train-mode parity is tested by INJECTING the keep-masks that the
# CUDA kernels generate (a test-side port of the kernels' counter hash supplies them).  site -> mask*scale tensor.
# --------------------------------------------------------------------------------------------------------------
DropFn = Optional[Callable[[str, torch.Size], torch.Tensor]]


def _drop(x: torch.Tensor, site: str, drop: DropFn) -> torch.Tensor:
    if drop is None:
        return x
    m = drop(site, x.shape)
    return x if m is None else x * m


# --------------------------------------------------------------------------------------------------------------
# context preparation
# --------------------------------------------------------------------------------------------------------------
def dist_to_north(az: torch.Tensor) -> torch.Tensor:
    """G/models.py:117-122 (K:146-151): |az| / 180."""
    return az.abs() / 180.0


def dist_to_east(az: torch.Tensor) -> torch.Tensor:
    """G/models.py:106-115 (K:135-144): |90-az|/180 if az >= -90 else (90+|az+180|)/180."""
    return torch.where(az >= -90.0, (90.0 - az).abs(), 90.0 + (az + 180.0).abs()) / 180.0


def fact_counts(facts: torch.Tensor, num_entities: int) -> torch.Tensor:
    """K/models.py:101-121 (N:98-116): number of facts whose subject is entity i; 0 for the last slot <unk_ent>."""
    B = facts.shape[0]
    subj = facts[:, :, 1].long()
    counts = torch.zeros(B, num_entities, dtype=torch.float32)
    ok = (subj >= 0) & (subj < num_entities)
    for b in range(B):
        counts[b].index_add_(0, subj[b][ok[b]], torch.ones(int(ok[b].sum()), dtype=torch.float32))
    counts[:, num_entities - 1] = 0.0
    return counts


def entity_encode(spec: Spec, p: Params, entities: torch.Tensor, facts: Optional[torch.Tensor]) -> torch.Tensor:
    """EntityEncoder.forward — G/models.py:82-104, K/models.py:82-133, N/models.py:79-134.  -> (B,E,D)."""
    B, E, _ = entities.shape
    D = spec.emb_dim
    ent = entities.float()
    out = torch.zeros(B, E, D, dtype=torch.float32)
    type_emb = p["entity_encoder.type_embedding.weight"]
    if spec.variant == "G":
        out[:, :, 0] = ent[:, :, 1]
        out[:, :, 1] = dist_to_north(ent[:, :, 2])
        out[:, :, 2] = dist_to_east(ent[:, :, 2])
        out[:, :, 3] = ent[:, :, 3]
        out = torch.cat([out[:, :, :4], type_emb[ent[:, :, 4].long()]], dim=2)
    elif spec.variant == "K":
        cnt = fact_counts(facts, E)
        out[:, :, 0] = ent[:, :, 1]
        out[:, :, 1] = dist_to_north(ent[:, :, 2])
        out[:, :, 2] = dist_to_east(ent[:, :, 2])
        out[:, :, 3] = ent[:, :, 3]
        out[:, :, 4] = cnt
        out[:, :, 5] = (cnt > 0).float()
        out = torch.cat([out[:, :, :6], type_emb[ent[:, :, 4].long()]], dim=2)
    else:  # N
        cnt = fact_counts(facts, E)
        out[:, :, 0] = ent[:, :, 1]
        out[:, :, 1] = ent[:, :, 2]
        out[:, :, 2] = ent[:, :, 3]
        out[:, :, 3] = cnt
        out[:, :, 4] = (cnt > 0).float()
        out = torch.cat([out[:, :, :5], type_emb[ent[:, :, 4].long()]], dim=2)
        # N/models.py:128-133: multiply by the mean of the five name-word embeddings
        name = p["word_embedding.weight"][ent[:, :, 5:].long()]  # (B,E,5,D)
        out = out * name.mean(dim=-2)
    return out


def fact_encode(p: Params, facts: torch.Tensor, ent_enc: torch.Tensor) -> torch.Tensor:
    """FactEncoder.forward — K/models.py:170-188 (N:153-171): subject encoding + predicate embedding."""
    subj = facts[:, :, 1].long()
    pred = facts[:, :, 2].long()
    gathered = torch.gather(ent_enc, 1, subj.unsqueeze(-1).expand(-1, -1, ent_enc.shape[-1]))
    return gathered + p["predicate_embedding.weight"][pred]


def caption_embed(
    spec: Spec,
    p: Params,
    captions: torch.Tensor,
    masks: torch.Tensor,
    ent_enc: torch.Tensor,
    fact_enc: Optional[torch.Tensor],
) -> torch.Tensor:
    """CaptionEmbedder.forward — G/models.py:143-181, K/models.py:209-259.  captions/masks (B,T) -> (B,T,D)."""
    V = spec.vocab_size
    E = ent_enc.shape[1]
    D = ent_enc.shape[-1]
    tok = captions.long()
    e_idx = tok - V
    e_idx = torch.where((e_idx < 0) | (e_idx >= E), torch.full_like(e_idx, E - 1), e_idx)
    w_idx = torch.where(tok >= V, torch.full_like(tok, spec.pad), tok)
    emb_w = p["word_embedding.weight"][w_idx]
    emb_e = torch.gather(ent_enc, 1, e_idx.unsqueeze(-1).expand(-1, -1, D))
    m = masks.long().unsqueeze(-1)
    out = torch.where(m == 1, emb_e, emb_w)
    if fact_enc is not None:
        Fn = fact_enc.shape[1]
        f_idx = tok - V - E
        f_idx = torch.where((f_idx < 0) | (f_idx >= Fn), torch.full_like(f_idx, Fn - 1), f_idx)
        emb_f = torch.gather(fact_enc, 1, f_idx.unsqueeze(-1).expand(-1, -1, D))
        out = torch.where(m == 2, emb_f, out)
    return out


def positional_table(max_len: int, d: int) -> torch.Tensor:
    """PositionEncoder.__init__ — G/models.py:184-205: sinusoidal table (max_len, d)."""
    pe = torch.zeros(max_len, d)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def context_indicators(
    spec: Spec, captions: torch.Tensor, facts: torch.Tensor, num_entities: int, out_length: int
) -> Tuple[torch.Tensor, torch.Tensor]:
    """
    get_context_indicators — K/models.py:380-418 (N:363-401), vectorised.

    entity_idx_before[b,t',f] = 1 iff an input position t < t' holds an entity token (value in [V, V+E)) whose slot
    equals facts[b,f,1]; predicate_indicator[b,t',p] = 1 iff such an enabled fact has predicate p.
    With out_length == 1 (predict) the OR is over ALL positions, no lag (K/models.py:406-409).
    Returns (B,out_length,F) and (B,out_length,NP) float tensors.
    """
    V = spec.vocab_size
    NP = spec.num_predicates
    B, T = captions.shape
    Fn = facts.shape[1]
    tok = captions.long()
    is_ent = (tok >= V) & (tok < V + num_entities)
    slot = tok - V
    subj = facts[:, :, 1].long()
    pred = facts[:, :, 2].long()
    # hit[b,t,f]: position t mentions the subject of fact f
    hit = is_ent.unsqueeze(-1) & (slot.unsqueeze(-1) == subj.unsqueeze(1))  # (B,T,F)
    if out_length == 1:
        before = hit.any(dim=1, keepdim=True)  # (B,1,F)
    else:
        assert out_length == T
        c = hit.long().cumsum(dim=1)
        excl = torch.cat([torch.zeros(B, 1, Fn, dtype=torch.long), c[:, :-1]], dim=1)
        before = excl > 0  # strictly earlier position
    ent_before = before.float()
    pi = torch.zeros(B, out_length, NP)
    idx = pred.unsqueeze(1).expand(-1, out_length, -1)
    pi.scatter_add_(2, idx, ent_before)
    pred_ind = (pi > 0).float()
    return ent_before, pred_ind


# --------------------------------------------------------------------------------------------------------------
# Transformer blocks (torch.nn.TransformerEncoderLayer / TransformerDecoderLayer, post-LN, ReLU, eps=1e-5;
# constructed at G/models.py:241-244, K/models.py:319-324)
# --------------------------------------------------------------------------------------------------------------
def mha(
    p: Params,
    prefix: str,
    q_in: torch.Tensor,
    kv_in: torch.Tensor,
    H: int,
    causal: bool,
    drop: DropFn = None,
    site: str = "",
) -> torch.Tensor:
    """nn.MultiheadAttention forward (packed in_proj, scale 1/sqrt(head_dim)), batch-major."""
    W = p[prefix + "in_proj_weight"]
    b = p[prefix + "in_proj_bias"]
    D = q_in.shape[-1]
    dh = D // H
    B, Sq, _ = q_in.shape
    Sk = kv_in.shape[1]
    q = q_in @ W[:D].T + b[:D]
    k = kv_in @ W[D : 2 * D].T + b[D : 2 * D]
    v = kv_in @ W[2 * D :].T + b[2 * D :]
    q = q.view(B, Sq, H, dh).transpose(1, 2)
    k = k.view(B, Sk, H, dh).transpose(1, 2)
    v = v.view(B, Sk, H, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if causal:
        # _generate_square_subsequent_mask — G/models.py:256-262: 0 on/below the diagonal, -inf above
        cm = torch.full((Sq, Sk), float("-inf")).triu(1)
        s = s + cm
    a = torch.softmax(s, dim=-1)
    a = _drop(a, site + ".attn", drop)
    o = (a @ v).transpose(1, 2).reshape(B, Sq, D)
    return o @ p[prefix + "out_proj.weight"].T + p[prefix + "out_proj.bias"]


def _ln(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), p[prefix + "weight"], p[prefix + "bias"], 1e-5)


def _ffn(p: Params, prefix: str, x: torch.Tensor, drop: DropFn, site: str) -> torch.Tensor:
    h = torch.relu(x @ p[prefix + "linear1.weight"].T + p[prefix + "linear1.bias"])
    h = _drop(h, site + ".ffn", drop)
    return h @ p[prefix + "linear2.weight"].T + p[prefix + "linear2.bias"]


def encoder_stack(p: Params, name: str, x: torch.Tensor, H: int, L: int, drop: DropFn = None) -> torch.Tensor:
    """nn.TransformerEncoder (no mask, no final norm) — call sites G/models.py:348, K/models.py:495-496."""
    for l in range(L):
        pre = f"{name}.layers.{l}."
        site = f"{name}.{l}"
        x = _ln(p, pre + "norm1.", x + _drop(mha(p, pre + "self_attn.", x, x, H, False, drop, site + ".sa"), site + ".d1", drop))
        x = _ln(p, pre + "norm2.", x + _drop(_ffn(p, pre, x, drop, site), site + ".d2", drop))
    return x


def decoder_stack(
    p: Params, x: torch.Tensor, memory: torch.Tensor, H: int, L: int, drop: DropFn = None
) -> torch.Tensor:
    """nn.TransformerDecoder with causal tgt_mask, no memory mask — call site G/models.py:358."""
    for l in range(L):
        pre = f"transformer_decoder.layers.{l}."
        site = f"transformer_decoder.{l}"
        x = _ln(p, pre + "norm1.", x + _drop(mha(p, pre + "self_attn.", x, x, H, True, drop, site + ".sa"), site + ".d1", drop))
        x = _ln(p, pre + "norm2.", x + _drop(mha(p, pre + "multihead_attn.", x, memory, H, False, drop, site + ".ca"), site + ".d2", drop))
        x = _ln(p, pre + "norm3.", x + _drop(_ffn(p, pre, x, drop, site), site + ".d3", drop))
    return x


# --------------------------------------------------------------------------------------------------------------
# scoring heads
# --------------------------------------------------------------------------------------------------------------
def get_scores(
    spec: Spec,
    p: Params,
    h: torch.Tensor,
    ent_enc: torch.Tensor,
    fact_enc: Optional[torch.Tensor],
    ent_before: Optional[torch.Tensor],
    pred_ind: Optional[torch.Tensor],
) -> torch.Tensor:
    """
    get_scores — G/models.py:291-313, K/models.py:420-455.  h (B,T,D) -> (B,T,V+E[+F]).

    The reference materialises (T,B,E,D) and applies Linear(D->1); sum_d h*ent*w + b is the same contraction.
    The fact mask multiplies the INPUT of fc_fact, so the bias is added after masking (K/models.py:451-452).
    """
    if spec.has_facts:
        gate = pred_ind @ p["fc_predicate.weight"].T + p["fc_predicate.bias"]  # K/models.py:436
        vin = h * gate  # :437
    else:
        vin = h
    s_vocab = vin @ p["fc_vocab.weight"].T + p["fc_vocab.bias"]
    we = p["fc_entity.weight"][0]
    s_ent = torch.einsum("btd,bed->bte", h * we, ent_enc) + p["fc_entity.bias"]
    parts = [s_vocab, s_ent]
    if spec.has_facts:
        wf = p["fc_fact.weight"][0]
        s_fact = torch.einsum("btd,bfd->btf", h * wf, fact_enc) * ent_before + p["fc_fact.bias"]
        parts.append(s_fact)
    return torch.cat(parts, dim=2)


# --------------------------------------------------------------------------------------------------------------
# forward / loss / predict
# --------------------------------------------------------------------------------------------------------------
def forward(
    spec: Spec,
    p: Params,
    captions: torch.Tensor,
    encoder_out: torch.Tensor,
    caption_masks: torch.Tensor,
    caption_lengths: torch.Tensor,
    entities: torch.Tensor,
    facts: Optional[torch.Tensor] = None,
    drop: DropFn = None,
) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """
    DecoderTransformer.forward — G/models.py:315-361, K/models.py:457-514, N/models.py:440-497.

    encoder_out is the reference's (B, D, P) channel-major encoder output.  Returns (scores (B,T,W), captions
    sorted by decreasing length, decode_lengths).
    """
    H, L, D = spec.num_heads, spec.num_layers, spec.emb_dim
    lengths, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True, stable=True)
    encoder_out = encoder_out[sort_ind]
    captions = captions[sort_ind]
    caption_masks = caption_masks[sort_ind]
    entities = entities[sort_ind]
    if facts is not None:
        facts = facts[sort_ind]
    decode_lengths = (lengths - 1).tolist()

    ent_enc = entity_encode(spec, p, entities, facts)
    fact_enc = fact_encode(p, facts, ent_enc) if spec.has_facts else None
    emb = caption_embed(spec, p, captions, caption_masks, ent_enc, fact_enc)
    pixels = encoder_out.permute(0, 2, 1)  # (B,P,D)
    ctx = [pixels, encoder_stack(p, "transformer_encoder_entities", ent_enc, H, L, drop)]
    if spec.has_facts:
        ctx.append(encoder_stack(p, "transformer_encoder_facts", fact_enc, H, L, drop))
    memory = torch.cat(ctx, dim=1)
    T = captions.shape[1]
    x = emb * math.sqrt(D) + p["pos_encoder.pe"][:T, 0, :].unsqueeze(0)
    x = _drop(x, "pos", drop)
    h = decoder_stack(p, x, memory, H, L, drop)
    if spec.has_facts:
        ent_before, pred_ind = context_indicators(spec, captions, facts, entities.shape[1], T)
    else:
        ent_before = pred_ind = None
    scores = get_scores(spec, p, h, ent_enc, fact_enc, ent_before, pred_ind)
    return scores, captions, decode_lengths


def packed_rows(decode_lengths: List[int], T: int) -> torch.Tensor:
    """Boolean (B,T) selector of the rows pack_padded_sequence keeps (G/train.py:277-278)."""
    dl = torch.tensor(decode_lengths).unsqueeze(1)
    return torch.arange(T).unsqueeze(0) < dl


def caption_loss(scores: torch.Tensor, captions_sorted: torch.Tensor, decode_lengths: List[int], pad: int = 0) -> torch.Tensor:
    """
    G/train.py:275-281: targets = captions_sorted[:,1:], both packed to decode_lengths,
    CrossEntropyLoss(ignore_index=<pad>) = mean over the kept non-pad targets.
    """
    B, T, W = scores.shape
    keep = packed_rows(decode_lengths, T - 1)
    tgt = captions_sorted[:, 1:]
    s = scores[:, : T - 1][keep]
    t = tgt[keep]
    return F.cross_entropy(s, t, ignore_index=pad)


def predict(
    spec: Spec,
    p: Params,
    encoder_out: torch.Tensor,
    max_pred_len: int,
    entities: torch.Tensor,
    facts: Optional[torch.Tensor] = None,
    return_margins: bool = False,
):
    """
    DecoderTransformer.predict — G/models.py:363-443, K/models.py:516-609.  Batch 1, greedy + repetition clean-up.

    Follows the reference step by step (full re-decode every step, no cache); returns (max_pred_len, 1) int64.
    If return_margins, also returns the per-step (top1 - top2) logit margins (test diagnostics for near-ties).
    """
    assert encoder_out.shape[0] == 1
    H, L, D = spec.num_heads, spec.num_layers, spec.emb_dim
    V = spec.vocab_size
    ent_enc = entity_encode(spec, p, entities, facts)
    E = ent_enc.shape[1]
    ctx = [encoder_out.permute(0, 2, 1), encoder_stack(p, "transformer_encoder_entities", ent_enc, H, L)]
    fact_enc = None
    if spec.has_facts:
        fact_enc = fact_encode(p, facts, ent_enc)
        ctx.append(encoder_stack(p, "transformer_encoder_facts", fact_enc, H, L))
    memory = torch.cat(ctx, dim=1)
    T = max_pred_len
    captions = torch.full((1, T), spec.start, dtype=torch.long)
    masks = torch.zeros((1, T), dtype=torch.long)
    output = torch.full((T,), spec.pad, dtype=torch.long)
    second: List[int] = []
    margins: List[float] = []
    pe = p["pos_encoder.pe"][:T, 0, :].unsqueeze(0)
    for i in range(T):
        emb = caption_embed(spec, p, captions, masks, ent_enc, fact_enc)
        x = emb * math.sqrt(D) + pe
        h = decoder_stack(p, x, memory, H, L)[:, i : i + 1]
        if spec.has_facts:
            eb, pi = context_indicators(spec, captions, facts, E, 1)
        else:
            eb = pi = None
        sc = get_scores(spec, p, h, ent_enc, fact_enc, eb, pi)[0, 0]
        top = torch.topk(sc, 2)
        margins.append(float(top.values[0] - top.values[1]))
        out = int(top.indices[0])
        output[i] = out
        if out == spec.end:
            break
        second.append(int(top.indices[1]))
        # repetition clean-up, G/models.py:418-435: repeat lengths 1,2,3 (dupl_idx 0,2,4), shortest first
        for dupl_idx in (0, 2, 4):
            if i > dupl_idx:
                n = (dupl_idx + 2) // 2
                one = [int(output[i - k]) for k in range(n)]
                two = [int(output[i - n - k]) for k in range(n)]
                if one == two:
                    for k in range(max(1, dupl_idx)):
                        output[i - k] = second[-(k + 1)]
                    break
        out = int(output[i])
        if i < T - 1:
            captions[0, i + 1] = out
            if spec.has_facts and out >= V + E:
                masks[0, i + 1] = 2
            elif out >= V:
                masks[0, i + 1] = 1
    res = output.view(T, 1)
    return (res, margins) if return_margins else res


def beam_search(
    spec: Spec,
    p: Params,
    encoder_out: torch.Tensor,
    max_pred_len: int,
    entities: torch.Tensor,
    facts: Optional[torch.Tensor] = None,
    beam_size: int = 5,
    return_margin: bool = False,
):
    """
    EXTENSION — PARITY UNPINNED BY THE REFERENCE.  /root/reference has no beam search (its predict() is greedy);
    BASELINE.json asks for beam-5.  This restates the published beam search of the Show-Attend-Tell tutorial that the
    reference's READMEs name as the origin of their infrastructure (geo-aware/README.md:37; tutorial eval.py,
    `evaluate(beam_size)`), run over THIS file's reference-pinned scoring function (the same caption_embed / decoder_stack /
    context_indicators / get_scores calls as `predict` above, i.e. K/models.py:548-575 for a batch of k partial captions):

        k live captions, all starting with <start>;  per step: candidate = score_so_far + log_softmax(scores)[token];
        step 1 takes the k best tokens of the single start row, later steps the k best of the flattened (k, W) table;
        a candidate ending in <end> is moved to the completed list and k shrinks; stop when k == 0 or after
        max_pred_len steps; the answer is the completed caption with the highest summed log-probability (no length
        normalisation, first maximum wins).  No repetition clean-up (that heuristic belongs to the greedy predict()).
        If nothing completed within max_pred_len steps (the tutorial would fail there) the best live caption is returned.

    Batch 1.  Returns (max_pred_len,) int64: tokens without <start>, <end> included, <pad>-filled — predict()'s convention.
    With return_margin also the smallest gap between the k-th selected and the best rejected candidate over all steps
    (test diagnostic for near-ties, where fp32 rounding may legitimately pick another beam).
    """
    assert encoder_out.shape[0] == 1
    H, L, D, V = spec.num_heads, spec.num_layers, spec.emb_dim, spec.vocab_size
    T = max_pred_len
    ent_enc = entity_encode(spec, p, entities, facts)
    E = ent_enc.shape[1]
    ctx = [encoder_out.permute(0, 2, 1), encoder_stack(p, "transformer_encoder_entities", ent_enc, H, L)]
    fact_enc = None
    if spec.has_facts:
        fact_enc = fact_encode(p, facts, ent_enc)
        ctx.append(encoder_stack(p, "transformer_encoder_facts", fact_enc, H, L))
    memory = torch.cat(ctx, dim=1)
    pe = p["pos_encoder.pe"][:T, 0, :].unsqueeze(0)
    k = beam_size
    seqs = torch.full((1, T), spec.start, dtype=torch.long)  # input tokens; row r, position i+1 = output of step i
    masks = torch.zeros((1, T), dtype=torch.long)
    outs: List[List[int]] = [[]]
    cum = torch.zeros(1)
    done: List[Tuple[float, List[int]]] = []
    margin = float("inf")
    for i in range(T):
        n = seqs.shape[0]
        x = caption_embed(spec, p, seqs, masks, ent_enc.expand(n, -1, -1), fact_enc.expand(n, -1, -1) if fact_enc is not None else None)
        x = x * math.sqrt(D) + pe
        h = decoder_stack(p, x, memory.expand(n, -1, -1), H, L)[:, i : i + 1]
        if spec.has_facts:
            eb, pi = context_indicators(spec, seqs, facts.expand(n, -1, -1), E, 1)
        else:
            eb = pi = None
        sc = get_scores(spec, p, h, ent_enc.expand(n, -1, -1), fact_enc.expand(n, -1, -1) if fact_enc is not None else None, eb, pi)[:, 0]
        W = sc.shape[1]
        cand = (cum.unsqueeze(1) + F.log_softmax(sc, dim=1)).reshape(-1)
        top = torch.topk(cand, min(k + 1, cand.numel()))
        if top.values.numel() > k:
            margin = min(margin, float(top.values[k - 1] - top.values[k]))
        vals, idx = top.values[:k], top.indices[:k]
        prev, nxt = (idx // W).tolist(), (idx % W).tolist()
        keep_rows, keep_tok, keep_val, new_outs = [], [], [], []
        for r in range(k):
            seq = outs[prev[r]] + [nxt[r]]
            if nxt[r] == spec.end:
                done.append((float(vals[r]), seq))
            else:
                keep_rows.append(prev[r])
                keep_tok.append(nxt[r])
                keep_val.append(vals[r])
                new_outs.append(seq)
        k = len(keep_rows)
        if k == 0:
            break
        outs = new_outs
        cum = torch.stack(keep_val)
        seqs = seqs[keep_rows].clone()
        masks = masks[keep_rows].clone()
        if i < T - 1:
            for r, tok in enumerate(keep_tok):
                seqs[r, i + 1] = tok
                masks[r, i + 1] = 2 if (spec.has_facts and tok >= V + E) else (1 if tok >= V else 0)
    if done:
        best = max(range(len(done)), key=lambda j: (done[j][0], -j))  # highest score, first one on ties
        seq = done[best][1]
    else:
        seq = outs[0]
    res = torch.full((T,), spec.pad, dtype=torch.long)
    res[: len(seq)] = torch.tensor(seq, dtype=torch.long)
    return (res, margin) if return_margin else res


def prepare_images(raw_f16, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> torch.Tensor:
    """
    Host image path of the reference, restated with the same numpy / torch calls: ``torch.FloatTensor(imgs[i] / 255.)``
    (G/datasets.py:44 — ``imgs`` is the fp16 HDF5 dataset of G/create_input_files.py:99-101, so the division is numpy's fp16
    division) followed by ``transforms.Normalize(mean, std)`` (G/train.py:139-147; torchvision: ``tensor.sub_(mean).div_(std)``
    in fp32).  raw_f16: (N, 3, H, W) numpy float16 array or torch.float16 tensor with values in [0, 255] -> fp32 tensor.
    """
    import numpy as np

    a = raw_f16.numpy() if torch.is_tensor(raw_f16) else np.asarray(raw_f16)
    assert a.dtype == np.float16
    x = torch.FloatTensor(a / 255.0)
    m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    return x.sub_(m).div_(s)
