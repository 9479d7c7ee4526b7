"""
Generates tests/golden/golden_beam_*.npz: beam-search captions computed with the UNMODIFIED reference modules as the scoring
function (imported from /root/reference, present only in the build container).

    python tests/golden/make_golden_beam.py

The reference has NO beam search — its predict() is greedy — so what is pinned here is: (1) the scoring function
(exactly the module calls predict() makes each step, K/models.py:548-575, run for a batch of k partial captions), and (2) this
file's transcription of the Show-Attend-Tell tutorial's beam search (the reference's READMEs name that tutorial as the origin of
their infrastructure, G/README.md:37).  oracle/decoder_oracle.py:beam_search must reproduce these captions
(tests/test_oracle_golden.py); the CUDA path is then checked against the oracle.

Random weights never rank <end> among the best tokens, so besides the plain weights (no caption completes: every beam runs to max_len and the best live
beam is returned) a second setting raises `fc_vocab.bias[<end>]` by END_BIAS (recorded in the file; the tests apply the same
change), which makes some captions complete early while the remaining beams go on with a smaller k.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ickb200  # noqa: E402,F401
from ickb200 import synthetic as syn  # noqa: E402
from make_golden import build_reference_decoder  # noqa: E402

END_BIAS = {"G": (0.0, 1.0), "K": (0.0, 0.5), "N": (0.0, 0.3)}  # no caption completes / some complete early
BEAM = 5


def step_scores(dec, variant, seqs, masks, i, ent_enc, fact_enc, memory, entities, facts):
    """The per-step computation of the reference's predict() for a batch of n partial captions -> (n, W) raw scores."""
    n, T = seqs.shape
    wm = dec.word_map
    if variant == "G":
        emb = dec.caption_embedder(seqs, ent_enc.expand(n, -1, -1), dec.word_embedding, wm["<pad>"], masks.unsqueeze(2))
    else:
        emb = dec.caption_embedder(seqs, ent_enc.expand(n, -1, -1), fact_enc.expand(n, -1, -1), dec.word_embedding, wm["<pad>"],
                                   masks.unsqueeze(2))
    emb = emb.permute(1, 0, 2) * math.sqrt(dec.emb_dim)
    emb = dec.pos_encoder(emb)
    if dec.lookahead_mask is None or dec.lookahead_mask.size(0) != T:
        dec.lookahead_mask = dec._generate_square_subsequent_mask(T)
    h = dec.transformer_decoder(emb, memory.expand(-1, n, -1), dec.lookahead_mask)[i].unsqueeze(0)
    if variant == "G":
        sc = dec.get_scores(h, ent_enc.expand(n, -1, -1))
    else:
        eb, pi = dec.get_context_indicators(seqs, facts.expand(n, -1, -1), entities.shape[1], 1)
        sc = dec.get_scores(h, ent_enc.expand(n, -1, -1), fact_enc.expand(n, -1, -1), eb, pi)
    return sc.squeeze(0)


def beam_search(dec, variant, encoder_out, T, entities, facts, k):
    wm = dec.word_map
    V = len(wm)
    if variant == "G":
        ent_enc = dec.entity_encoder(entities)
    elif variant == "K":
        ent_enc = dec.entity_encoder(entities, facts)
    else:
        ent_enc = dec.entity_encoder(entities, facts, dec.word_embedding)
    E = ent_enc.shape[1]
    ctx = [encoder_out.permute(2, 0, 1), dec.transformer_encoder_entities(ent_enc.permute(1, 0, 2))]
    fact_enc = None
    if variant != "G":
        fact_enc = dec.fact_encoder(facts, ent_enc)
        ctx.append(dec.transformer_encoder_facts(fact_enc.permute(1, 0, 2)))
    memory = torch.cat(ctx)
    seqs = torch.full((1, T), wm["<start>"], dtype=torch.long)
    masks = torch.zeros((1, T), dtype=torch.long)
    outs = [[]]
    cum = torch.zeros(1)
    done = []
    margin = float("inf")
    for i in range(T):
        sc = step_scores(dec, variant, seqs, masks, i, ent_enc, fact_enc, memory, entities, facts)
        W = sc.shape[1]
        cand = (cum.unsqueeze(1) + torch.log_softmax(sc, dim=1)).reshape(-1)
        top = cand.topk(min(k + 1, cand.numel()))
        if top.values.numel() > k:
            margin = min(margin, float(top.values[k - 1] - top.values[k]))
        vals, idx = top.values[:k], top.indices[:k]
        prev, nxt = (idx // W).tolist(), (idx % W).tolist()
        rows, toks, keep, new_outs = [], [], [], []
        for r in range(k):
            seq = outs[prev[r]] + [nxt[r]]
            if nxt[r] == wm["<end>"]:
                done.append((float(vals[r]), seq))
            else:
                rows.append(prev[r]); toks.append(nxt[r]); keep.append(vals[r]); new_outs.append(seq)
        k = len(rows)
        if k == 0:
            break
        outs, cum = new_outs, torch.stack(keep)
        seqs, masks = seqs[rows].clone(), masks[rows].clone()
        if i < T - 1:
            for r, tok in enumerate(toks):
                seqs[r, i + 1] = tok
                masks[r, i + 1] = 2 if (variant != "G" and tok >= V + E) else (1 if tok >= V else 0)
    if done:
        j = max(range(len(done)), key=lambda j: (done[j][0], -j))
        score, seq = done[j]
    else:
        score, seq = float(cum[0]), outs[0]
    res = np.full((T,), wm["<pad>"], dtype=np.int64)
    res[: len(seq)] = seq
    return res, score, margin, len(done)


def scan():
    for variant, cfg in syn.SMALL_CONFIGS.items():
        for bias in (0.0, 0.5, 1.0, 1.5):
            dec = build_reference_decoder(variant, cfg)
            with torch.no_grad():
                dec.fc_vocab.bias[dec.word_map["<end>"]] += bias
            T = 12 if variant == "G" else 14
            cfg6 = cfg.with_batch(6)
            pb = syn.make_batch(cfg6, seed=301)
            with torch.no_grad():
                res = []
                for i in range(cfg6.B):
                    facts = pb["facts"][i : i + 1] if cfg.has_facts else None
                    r, s_, m, nd = beam_search(dec, variant, pb["encoder_out"][i : i + 1], T, pb["entities"][i : i + 1], facts, BEAM)
                    res.append((int((r != 0).sum()), nd, round(m, 4)))
            print(variant, bias, res, flush=True)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for variant, cfg in syn.SMALL_CONFIGS.items():
        T = 12 if variant == "G" else 14
        cfg6 = cfg.with_batch(6)
        pb = syn.make_batch(cfg6, seed=301)
        rec = {"seed": np.asarray(301), "max_len": np.asarray(T), "beam": np.asarray(BEAM), "batch": np.asarray(cfg6.B),
               "end_bias": np.asarray(END_BIAS[variant])}
        for j, bias in enumerate(END_BIAS[variant]):
            dec = build_reference_decoder(variant, cfg)
            with torch.no_grad():
                dec.fc_vocab.bias[dec.word_map["<end>"]] += bias
            toks, scores, margins, ndone = [], [], [], []
            with torch.no_grad():
                for i in range(cfg6.B):
                    facts = pb["facts"][i : i + 1] if cfg.has_facts else None
                    r, s, m, nd = beam_search(dec, variant, pb["encoder_out"][i : i + 1], T, pb["entities"][i : i + 1], facts, BEAM)
                    toks.append(r); scores.append(s); margins.append(m); ndone.append(nd)
            rec.update({f"tokens_{j}": np.stack(toks), f"scores_{j}": np.asarray(scores), f"margins_{j}": np.asarray(margins),
                        f"completed_{j}": np.asarray(ndone)})
            print(variant, "end bias", bias, "completed", ndone, "min margin", min(margins))
            print(rec[f"tokens_{j}"])
        np.savez_compressed(os.path.join(out_dir, f"golden_beam_{variant}.npz"), **rec)


if __name__ == "__main__":
    scan() if "--scan" in sys.argv else main()
