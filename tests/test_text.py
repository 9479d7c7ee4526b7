"""Token ids -> caption strings (ickb200.text) against hand-worked cases of the reference's eval conversion (G/eval.py:86-116,
K/eval.py:103-170)."""
import torch

import ickb200  # noqa: F401
from ickb200 import text


def slots(names, width=12):
    """char-coded name slots as the datasets store them: [row, length, codes..., dummy padding] (G/utils.py:154-175)."""
    rows = []
    for i, n in enumerate(names):
        codes = [ord(c) for c in n][:width]
        rows.append([i, len(codes)] + codes + [0] * (width - len(codes)))
    return torch.tensor(rows)


WM = {"<pad>": 0, "a": 1, "view": 2, "of": 3, ".": 4, "near": 5, "<unk>": 6, "<start>": 7, "<end>": 8}
V = len(WM)


def test_vocabulary_entities_and_specials():
    ents = slots(["Tower Bridge", "Thames", "<unk_ent>"]).unsqueeze(0)
    toks = torch.tensor([[1, 2, 3, V + 0, 5, V + 1, 4, 8, 0, 0]])
    assert text.captions_to_text(toks, WM, ents) == ["a view of Tower Bridge near Thames ."]
    # a pointer behind the last slot (cannot be produced by the decoder, but the reference guards it)
    assert text.captions_to_text(torch.tensor([[1, V + 3, 8]]), WM, ents) == ["a <unk_ent>"]


def test_facts_and_batch():
    ents = torch.stack([slots(["Oslo", "x"]), slots(["Rome", "y"])])
    facts = torch.stack([slots(["1048", "Norway", "z"]), slots(["753 BC", "Italy", "w"])])
    toks = torch.tensor([[7, V + 0, 3, V + 2 + 1, 8, 0], [V + 0, 1, V + 2 + 0, 4, 8, 0]])
    assert text.captions_to_text(toks, WM, ents, facts) == ["Oslo of Norway", "Rome a 753 BC ."]
    assert text.captions_to_text(torch.tensor([[V + 2 + 3, 8]]), WM, ents[:1], facts[:1]) == ["<unk_fact>"]


def test_unfinished_last_sentence_is_dropped():
    ents = slots(["q"]).unsqueeze(0)
    toks = torch.tensor([[1, 4, 2, 4, 3, 5, 8]])  # "a . view . of near" -> two full stops, no final one
    assert text.captions_to_text(toks, WM, ents) == ["a . view ."]
    assert text.tidy("one sentence only") == "one sentence only" and text.tidy("a. b.") == "a. b."


def test_name_codec():
    assert text.decode_name([72, 105, 33, 0, 0], 2) == "Hi" and text.decode_name([72, 105], 5) == "Hi"
