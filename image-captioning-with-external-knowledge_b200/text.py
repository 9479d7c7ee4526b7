"""
Host side of caption generation: token ids -> caption strings for a whole batch of generated sequences
after ONE device->host copy, following the inline conversion of the reference's eval scripts (G/eval.py:86-116,
K/eval.py:103-170): vocabulary ids map through the reversed word map (``<start>``, ``<end>``, ``<pad>`` are dropped),
ids in ``[V, V+E)`` print the name of the pointed-to entity, ids ``>= V+E`` the object of the pointed-to fact.  Names
arrive char-coded as the datasets store them (``[row, length, c0, c1, ...]`` per slot, codec G/utils.py:154-192); a pointer
behind the last slot prints ``<unk_ent>`` / ``<unk_fact>``.  Pure Python/numpy text processing - nothing here touches the GPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


def decode_name(codes: Sequence[int], length: int) -> str:
    """ut.int_to_str (G/utils.py:177-192): the first ``length`` character codes of a slot."""
    return "".join(chr(int(c)) for c in list(codes)[: max(int(length), 0)])


def _names_of(slots) -> List[str]:
    """(S, 2 + L) char-coded slots -> S strings (column 1 = length, columns 2.. = codes)."""
    a = np.asarray(slots.cpu() if torch.is_tensor(slots) else slots)
    return [decode_name(row[2:], row[1]) for row in a]


def tidy(caption: str) -> str:
    """G/eval.py:113-114: a caption with several sentences that does not end in a full stop loses its unfinished last one."""
    if not caption.endswith(".") and caption.count(".") > 1:
        caption = ".".join(caption.split(".")[:-1]) + "."
    return caption


def captions_to_text(tokens, word_map: Dict[str, int], entity_names, fact_names=None) -> List[str]:
    """
    tokens: (N, T) int tensor/array (``predict_batch`` output, one row per image); entity_names: (N, E, 2+L) char-coded entity
    names; fact_names: (N, F, 2+L) or None (geo-aware).  Returns the N caption strings of the reference's eval loop.
    """
    toks = np.asarray(tokens.cpu() if torch.is_tensor(tokens) else tokens)  # the single D2H copy
    V = len(word_map)
    rev = {v: k for k, v in word_map.items()}
    skip = {word_map["<start>"], word_map["<end>"], word_map["<pad>"]}
    out = []
    for i, seq in enumerate(toks):
        ents: Optional[List[str]] = None
        facts: Optional[List[str]] = None
        E = entity_names[i].shape[0]
        words = []
        for t in seq.tolist():
            if t >= V and (fact_names is None or t < V + E):
                if ents is None:
                    ents = _names_of(entity_names[i])  # decoded lazily: most captions point at a handful of slots
                j = t - V
                words.append(ents[j] if j < E else "<unk_ent>")
            elif t >= V + E:
                if facts is None:
                    facts = _names_of(fact_names[i])
                j = t - V - E
                words.append(facts[j] if j < len(facts) else "<unk_fact>")
            elif t not in skip:
                words.append(rev[t])
        out.append(tidy(" ".join(words)))
    return out
