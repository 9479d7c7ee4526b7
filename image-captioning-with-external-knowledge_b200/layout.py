"""
Parameter inventory of the three ``DecoderTransformer`` variants, keyed exactly like the reference's
``state_dict`` (SURVEY.md §8b "Ownership / state"; constructors G/models.py:217-254, K/models.py:295-339,
N/models.py:278-322), so checkpoints interchange with the reference.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

NUM_FEAT = {"G": 4, "K": 6, "N": 5}        # hand-made feature columns before the type embedding
NUM_TYPES = {"G": 1000, "K": 1000, "N": 20}  # nn.Embedding(1000, D-4/6) / nn.Embedding(20, D-5)
NUM_PRED = {"G": 0, "K": 3000, "N": 3500}


def _attn(prefix: str, D: int, out: Dict[str, Tuple[int, ...]]) -> None:
    out[prefix + "in_proj_weight"] = (3 * D, D)
    out[prefix + "in_proj_bias"] = (3 * D,)
    out[prefix + "out_proj.weight"] = (D, D)
    out[prefix + "out_proj.bias"] = (D,)


def _ffn_norms(prefix: str, D: int, ff: int, n_norm: int, out: Dict[str, Tuple[int, ...]]) -> None:
    out[prefix + "linear1.weight"] = (ff, D)
    out[prefix + "linear1.bias"] = (ff,)
    out[prefix + "linear2.weight"] = (D, ff)
    out[prefix + "linear2.bias"] = (D,)
    for i in range(1, n_norm + 1):
        out[prefix + f"norm{i}.weight"] = (D,)
        out[prefix + f"norm{i}.bias"] = (D,)


def param_shapes(variant: str, V: int, D: int = 300, L: int = 3, ff_dec: int = 512, ff_enc: int = 512
                 ) -> "OrderedDict[str, Tuple[int, ...]]":
    """Trainable parameters in the reference's registration order (``named_parameters()`` order)."""
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for l in range(L):
        pre = f"transformer_decoder.layers.{l}."
        _attn(pre + "self_attn.", D, out)
        _attn(pre + "multihead_attn.", D, out)
        _ffn_norms(pre, D, ff_dec, 3, out)
    stacks = ["transformer_encoder_entities"] + (["transformer_encoder_facts"] if variant != "G" else [])
    for name in stacks:
        for l in range(L):
            pre = f"{name}.layers.{l}."
            _attn(pre + "self_attn.", D, out)
            _ffn_norms(pre, D, ff_enc, 2, out)
    out["word_embedding.weight"] = (V, D)
    out["entity_encoder.type_embedding.weight"] = (NUM_TYPES[variant], D - NUM_FEAT[variant])
    if variant != "G":
        out["predicate_embedding.weight"] = (NUM_PRED[variant], D)
    out["fc_vocab.weight"] = (V, D)
    out["fc_vocab.bias"] = (V,)
    out["fc_entity.weight"] = (1, D)
    out["fc_entity.bias"] = (1,)
    if variant != "G":
        out["fc_fact.weight"] = (1, D)
        out["fc_fact.bias"] = (1,)
        out["fc_predicate.weight"] = (D, NUM_PRED[variant])
        out["fc_predicate.bias"] = (D,)
    return out
