"""Shared test helpers: golden fixtures, oracle specs/params, error metrics."""
import os

import numpy as np
import torch

import ickb200  # noqa: F401  (import shim)
from ickb200 import layout, synthetic as syn
from oracle import decoder_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(variant):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"golden_{variant}.npz")))


def load_base_golden(variant):
    """BASELINE-shape fixtures (E, F, T, V of SURVEY.md §8d at batch 8) written by make_golden.py --baseline."""
    return dict(np.load(os.path.join(GOLDEN_DIR, f"golden_base_{variant}.npz")))


def score_views(scores, V):
    """What the BASELINE-shape fixtures keep of the (B, T, V+E+F) scores (34 MB in fp32 for K at B=8): a strided grid over all
    columns, every pointer column at every third position, and per-row logsumexp / max / argmax over the full row."""
    sc = torch.as_tensor(scores).detach().cpu().float()
    s = sc.double()
    return {"s_grid": sc[:, ::5, ::17].numpy().copy(), "s_ptr": sc[:, ::3, V:].numpy().copy(),
            "s_lse": torch.logsumexp(s, -1).numpy(), "s_rowmax": s.max(-1).values.numpy(), "s_argmax": s.argmax(-1).numpy()}


def spec_for(cfg):
    return orc.Spec(cfg.variant, cfg.V, cfg.D, cfg.H, cfg.L, pad=0, start=cfg.V - 2, end=cfg.V - 1)


def oracle_params(cfg, seed=0, requires_grad=False, profile="test"):
    shapes = layout.param_shapes(cfg.variant, cfg.V, cfg.D, cfg.L, cfg.ff, cfg.ff)
    p = syn.det_weights(shapes, seed=seed, profile=profile)
    p["pos_encoder.pe"] = orc.positional_table(5000, cfg.D).unsqueeze(1)
    if requires_grad:
        for k, v in p.items():
            if k != "pos_encoder.pe":
                v.requires_grad_(True)
    return p


def batch_args(cfg, batch):
    a = [batch["captions"], batch["encoder_out"], batch["caption_masks"], batch["caption_lengths"], batch["entities"]]
    if cfg.has_facts:
        a.append(batch["facts"])
    return a


def nmax_err(a, b):
    """max |a-b| / max |b|  (normalised max error; an absolute floor is needed on near-zero logits)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---- building the product module for tests -------------------------------------------------------------------------------
def module_cls(variant):
    from ickb200 import geo_aware, knowledge_aware, news_knowledge_aware

    return {"G": geo_aware, "K": knowledge_aware, "N": news_knowledge_aware}[variant].DecoderTransformer


def build_module(cfg, device, dtype=torch.float32, seed=0, dropouts=(0.5, 0.5, 0.1), profile="test"):
    wm = syn.make_word_map(cfg.V)
    dec = module_cls(cfg.variant)(wm, cfg.D, cfg.ff, cfg.ff, cfg.H, cfg.L, dropout_dec=dropouts[0], dropout_enc=dropouts[1],
                                  dropout_pos=dropouts[2], compute_dtype=dtype)
    shapes = layout.param_shapes(cfg.variant, cfg.V, cfg.D, cfg.L, cfg.ff, cfg.ff)
    w = syn.det_weights(shapes, seed=seed, profile=profile)
    missing = dec.load_state_dict(w, strict=False)
    assert set(missing.missing_keys) <= {"pos_encoder.pe", "fact_encoder.predicate_embedding.weight"}, missing
    return dec.to(device)


def oracle_drop_fn(seed, ps):
    """DropFn for the oracle that reproduces the kernels' masks: site name -> hash-derived multiplier tensor."""
    from dropout_ref import attn_drop_mul, drop_mul

    def fn(site, shape):
        if site == "pos":
            p = ps["pos"]
        elif site.startswith("transformer_decoder"):
            p = ps["dec"]
        else:
            p = ps["enc"]
        if p <= 0:
            return None
        rows = 1
        for s in shape[:-1]:
            rows *= s
        fn = attn_drop_mul if site.endswith(".attn") else drop_mul  # attention probabilities: bit-parallel keep words
        return fn(p, seed, layout.site_id(site), rows, shape[-1]).view(*shape)

    return fn
