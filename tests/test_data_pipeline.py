"""
CPU: host half of the input pipeline (ickb200.data.CaptionBatchSource) against the reference's own item conversion + default
collate (G/datasets.py:43-54, K/datasets.py:51-62), restated here with the reference's exact expressions on in-memory data
(datasets.py itself needs h5py, which is not installed; the per-item code does not depend on it).
"""
import numpy as np
import torch
from torch.utils.data import default_collate

import ickb200  # noqa: F401
from ickb200 import synthetic as syn
from ickb200.data import CaptionBatchSource
from oracle import decoder_oracle as orc


def reference_item(store, i, normalize):
    """K/datasets.py:51-62 verbatim expressions over python lists / an fp16 array."""
    img = torch.FloatTensor(store["imgs"][i] / 255.0)
    img = normalize(img)
    caption = torch.LongTensor(store["captions"][i])
    caplen = torch.LongTensor([store["caplens"][i]])
    capmask = torch.LongTensor(store["capmasks"][i])
    ent = torch.Tensor([x for x in store["entity_features"][i]])
    names = torch.LongTensor([x for x in store["entity_names"][i]])
    facts = torch.LongTensor([x for x in store["facts"][i]])
    fnames = torch.LongTensor([x for x in store["fact_names"][i]])
    return img, caption, caplen, capmask, ent, names, facts, fnames


def make_store(n=11, seed=0):
    cfg = syn.SMALL_CONFIGS["K"].with_batch(n)
    b = syn.make_batch(cfg, seed=seed)
    rng = np.random.default_rng(seed)
    return cfg, {
        "imgs": (rng.random((n, 3, 16, 24)) * 255).astype(np.float16),
        "captions": b["captions"].tolist(), "caplens": b["caption_lengths"].view(-1).tolist(), "capmasks": b["caption_masks"].tolist(),
        "entity_features": b["entities"].tolist(), "entity_names": rng.integers(0, 60, (n, cfg.E, 4)).tolist(),
        "facts": b["facts"].tolist(), "fact_names": rng.integers(0, 60, (n, cfg.F, 6)).tolist(),
    }


def test_batch_equals_reference_items_collated():
    import torchvision.transforms as T

    cfg, st = make_store()
    norm = T.Compose([T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    src = CaptionBatchSource(st["imgs"], st["captions"], st["caplens"], st["capmasks"], st["entity_features"], st["entity_names"],
                             st["facts"], st["fact_names"], pin_memory=False)
    idx = [7, 2, 9, 0, 2]
    ref = default_collate([reference_item(st, i, norm) for i in idx])
    got = src.batch(idx)
    assert got[0].dtype == torch.float16 and torch.equal(orc.prepare_images(got[0]), ref[0])  # device kernel == oracle (GPU test)
    for g, r in zip(got[1:], ref[1:]):
        assert g.dtype == r.dtype and g.shape == r.shape and torch.equal(g, r)


def test_epoch_covers_every_caption_once():
    cfg, st = make_store(n=10)
    src = CaptionBatchSource(st["imgs"], st["captions"], st["caplens"], st["capmasks"], st["entity_features"], st["entity_names"],
                             pin_memory=False)
    seen = []
    for raw, caps, lens, masks, ents, names in src.batches(4, shuffle=True, seed=3):
        assert raw.shape[0] == caps.shape[0] <= 4 and lens.shape[1] == 1
        seen += [tuple(c.tolist()) for c in caps]
    assert sorted(seen) == sorted(tuple(c) for c in st["captions"])
    assert len(list(src.batches(4, shuffle=False, drop_last=True))) == 2


def test_from_files_reads_the_reference_file_layout(tmp_path):
    """The JSON / pickle files CaptionDataset opens (K/datasets.py:19-49), written with the reference's names; images passed in
    (h5py is not installed here - and without `images=` that must be an ImportError, not a fallback)."""
    import json
    import pickle

    import pytest

    cfg, st = make_store(n=6, seed=2)
    name = "synthetic_5_cap_per_img"
    for kind, key in (("CAPTIONS", "captions"), ("CAPLENS", "caplens"), ("CAPMASKS", "capmasks")):
        with open(tmp_path / f"VAL_{kind}_{name}.json", "w") as f:
            json.dump(st[key], f)
    for kind, key in (("ENT_FEATURES", "entity_features"), ("ENT_NAMES", "entity_names"), ("FACTS", "facts"), ("FACT_NAMES", "fact_names")):
        with open(tmp_path / f"VAL_{kind}_{name}.pkl", "wb") as f:
            pickle.dump(st[key], f)
    src = CaptionBatchSource.from_files(str(tmp_path), name, "VAL", images=st["imgs"], pin_memory=False)
    ref = CaptionBatchSource(st["imgs"], st["captions"], st["caplens"], st["capmasks"], st["entity_features"], st["entity_names"],
                             st["facts"], st["fact_names"], pin_memory=False)
    for a, b in zip(src.batch([4, 1, 5]), ref.batch([4, 1, 5])):
        assert torch.equal(a, b)
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            CaptionBatchSource.from_files(str(tmp_path), name, "VAL", pin_memory=False)
