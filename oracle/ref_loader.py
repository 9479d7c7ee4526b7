"""
Loader for the UNMODIFIED reference model files under oracle/_ref/ (see oracle/make_ref.py) — TEST / BENCH INFRASTRUCTURE ONLY.

Only bench.py's reference legs (`--impl reference`, `cpu_baseline`) use this.  The reference fixes its device at import time
(`device = cuda if available else cpu`, geo-aware/models.py:6); the reference arm times the reference's CPU implementation on
the host cores, so the module global is pointed at the CPU after import (functions read it at call time) - nothing else is
touched.
"""
import importlib.util
import json
import os
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
PKG = {"G": "geo_aware", "K": "knowledge_aware", "N": "news_knowledge_aware"}


def available() -> bool:
    return all(os.path.exists(os.path.join(REF, p, "models.py")) for p in PKG.values())


def manifest() -> Optional[dict]:
    path = os.path.join(REF, "MANIFEST.json")
    return json.load(open(path)) if os.path.exists(path) else None


def load_models(variant: str):
    path = os.path.join(REF, PKG[variant], "models.py")
    spec = importlib.util.spec_from_file_location(f"ickref_models_{variant}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.device = torch.device("cpu")
    return mod


def build_decoder(variant: str, word_map, emb_dim=300, ff=512, heads=10, layers=3):
    """The reference's own construction call (geo-aware/train.py:71-78): default dropouts 0.5 / 0.5 / 0.1."""
    import warnings

    mod = load_models(variant)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dec = mod.DecoderTransformer(word_map=word_map, emb_dim=emb_dim, decoder_dim=ff, encoder_dim=ff, num_heads=heads, num_layers=layers)
    return dec
