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


# ======================================================================================================================
# Packing plan: how the reference-layout fp32 master parameters map to the padded, head-aligned operand copies the
# kernels read, and how packed gradient indices map back to the flat master-gradient buffer.
#
#   dense layout : logical width D (300) in a row of DP (320) elements, pad columns zero
#   head layout  : head h in columns [HD*h, HD*h + dh), HD = 32, pad lanes zero  (H*HD = 320 = DP for the reference sizes)
# Padding is done ONCE in the weights (zero rows / columns), so activations never need masking and every GEMM operand
# row is 16-byte aligned for TMA; state_dict() still sees only the reference-shaped master parameters.
# ======================================================================================================================
import zlib  # noqa: E402
from dataclasses import dataclass, field  # noqa: E402
from typing import List, Optional  # noqa: E402

import numpy as np  # noqa: E402

HD = 32  # padded head dim


def site_id(name: str) -> int:
    """Dropout site id shared by the engine and the test-side mask generator."""
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class Linear:
    """One packed projection: W [Np, Kp] (K-major) and its transpose WT [Kp, ldT] in packT, bias [Np] in packF."""
    name: str
    Np: int
    Kp: int
    w_off: int = 0       # element offset of W in packT
    wt_off: int = 0      # element offset of WT in packT
    ldT: int = 0         # leading dimension of WT (Np rounded up to 8)
    b_off: int = 0       # element offset of the bias in packF
    # gradient maps (numpy int32): packed n -> flat offset of the master row (or -1); packed k -> master column (or -1)
    rowoff: np.ndarray = field(default=None, repr=False)
    colmap: np.ndarray = field(default=None, repr=False)
    biasoff: np.ndarray = field(default=None, repr=False)


class PackPlan:
    def __init__(self, variant: str, V: int, D: int = 300, H: int = 10, L: int = 3, ff_dec: int = 512, ff_enc: int = 512):
        assert D % H == 0
        self.variant, self.V, self.D, self.H, self.L = variant, V, D, H, L
        self.ff_dec, self.ff_enc = ff_dec, ff_enc
        self.dh = D // H
        assert self.dh <= HD
        self.DP = H * HD
        assert self.DP >= D and self.DP % 8 == 0
        self.has_facts = variant != "G"
        self.NP = NUM_PRED[variant]
        self.shapes = param_shapes(variant, V, D, L, ff_dec, ff_enc)
        self.offsets = {}
        off = 0
        for k, shp in self.shapes.items():
            self.offsets[k] = off
            off += int(np.prod(shp))
        self.n_params = off
        self.dstA = np.full(off, -1, dtype=np.int32)
        self.dstB = np.full(off, -1, dtype=np.int32)
        self.dstC = np.full(off, -1, dtype=np.int32)
        self._t = 0  # packT cursor (elements)
        self._f = 0  # packF cursor
        self.linears = {}
        self.regions_T = {}
        self.regions_F = {}
        self._build()
        self.packT_size = self._t
        self.packF_size = self._f

    # ---- index helpers ------------------------------------------------------------------------------------------------
    def dense_map(self, n: int) -> np.ndarray:
        """packed dense index (length n>=D) -> original index or -1"""
        m = np.full(n, -1, dtype=np.int64)
        m[: self.D] = np.arange(self.D)
        return m

    def head_map(self) -> np.ndarray:
        """packed head-layout index (length DP) -> original index h*dh + c or -1"""
        m = np.full(self.DP, -1, dtype=np.int64)
        for h in range(self.H):
            m[h * HD : h * HD + self.dh] = h * self.dh + np.arange(self.dh)
        return m

    def _alloc_T(self, n: int) -> int:
        o = self._t
        self._t += _round_up(n, 64)  # keep every region 128-byte aligned in bf16
        return o

    def _alloc_F(self, n: int) -> int:
        o = self._f
        self._f += _round_up(n, 32)
        return o

    def _add_linear(self, name: str, rows: List[tuple], kmap: np.ndarray, biases: Optional[List[tuple]]) -> Linear:
        """
        rows: list of (weight_key, orig_row_index_array) blocks stacked along packed n; orig index -1 = zero row.
        kmap: packed k -> orig column or -1.  biases: matching list of (bias_key, orig_index_array) or None.
        """
        Np = sum(len(r[1]) for r in rows)
        Kp = len(kmap)
        lin = Linear(name, Np, Kp)
        lin.ldT = _round_up(Np, 8)
        lin.w_off = self._alloc_T(Np * Kp)
        lin.wt_off = self._alloc_T(Kp * lin.ldT)
        lin.b_off = self._alloc_F(Np)
        lin.rowoff = np.full(Np, -1, dtype=np.int32)
        lin.biasoff = np.full(Np, -1, dtype=np.int32)
        lin.colmap = kmap.astype(np.int32)
        kvalid = np.nonzero(kmap >= 0)[0]
        n0 = 0
        for bi, (wkey, rmap) in enumerate(rows):
            Ko = self.shapes[wkey][1]
            woff = self.offsets[wkey]
            for i, ro in enumerate(rmap):
                if ro < 0:
                    continue
                n = n0 + i
                lin.rowoff[n] = woff + int(ro) * Ko
                src = woff + int(ro) * Ko + kmap[kvalid]
                self.dstA[src] = lin.w_off + n * Kp + kvalid
                self.dstB[src] = lin.wt_off + kvalid * lin.ldT + n
                if biases is not None:
                    bkey, bmap = biases[bi]
                    lin.biasoff[n] = self.offsets[bkey] + int(bmap[i])
                    self.dstC[self.offsets[bkey] + int(bmap[i])] = lin.b_off + n
            n0 += len(rmap)
        self.linears[name] = lin
        return lin

    def _build(self) -> None:
        D, DP, H = self.D, self.DP, self.H
        dense = self.dense_map(DP)
        head = self.head_map()

        def attn_self(pre: str):
            blocks_w, blocks_b = [], []
            for part in range(3):
                r = np.where(head >= 0, head + part * D, -1)
                blocks_w.append((pre + "in_proj_weight", r))
                blocks_b.append((pre + "in_proj_bias", r))
            self._add_linear(pre + "qkv", blocks_w, dense, blocks_b)
            self._add_linear(pre + "out", [(pre + "out_proj.weight", dense)], head, [(pre + "out_proj.bias", dense)])

        def ffn(pre: str, ff: int):
            idf = np.arange(ff)
            self._add_linear(pre + "ffn1", [(pre + "linear1.weight", idf)], dense, [(pre + "linear1.bias", idf)])
            self._add_linear(pre + "ffn2", [(pre + "linear2.weight", dense)], idf, [(pre + "linear2.bias", dense)])

        kv_w, kv_b = [], []
        for l in range(self.L):
            pre = f"transformer_decoder.layers.{l}."
            attn_self(pre + "self_attn.")
            ca = pre + "multihead_attn."
            self._add_linear(ca + "q", [(ca + "in_proj_weight", head)], dense, [(ca + "in_proj_bias", head)])
            for part in (1, 2):
                r = np.where(head >= 0, head + part * D, -1)
                kv_w.append((ca + "in_proj_weight", r))
                kv_b.append((ca + "in_proj_bias", r))
            self._add_linear(ca + "out", [(ca + "out_proj.weight", dense)], head, [(ca + "out_proj.bias", dense)])
            ffn(pre, self.ff_dec)
        # the memory K/V projections of all decoder layers as ONE GEMM over the shared memory tensor
        self._add_linear("transformer_decoder.kv_all", kv_w, dense, kv_b)
        stacks = ["transformer_encoder_entities"] + (["transformer_encoder_facts"] if self.has_facts else [])
        for name in stacks:
            for l in range(self.L):
                pre = f"{name}.layers.{l}."
                attn_self(pre + "self_attn.")
                ffn(pre, self.ff_enc)
        idv = np.arange(self.V)
        self._add_linear("fc_vocab", [("fc_vocab.weight", idv)], dense, [("fc_vocab.bias", idv)])
        # word embedding rows in dense layout (gathered by the caption embedder / news entity encoder)
        o = self._alloc_T(self.V * DP)
        self.regions_T["word_embedding"] = (o, self.V, DP)
        base = self.offsets["word_embedding.weight"]
        idx = np.arange(self.V)[:, None] * D + np.arange(D)[None, :]
        self.dstA[base + idx.ravel()] = (o + np.arange(self.V)[:, None] * DP + np.arange(D)[None, :]).ravel()
        if self.has_facts:
            # fc_predicate.weight (D, NP) transposed to (NP, DP) fp32 for row gathers in the gate kernel
            o = self._alloc_F(self.NP * DP)
            self.regions_F["fc_predicate_T"] = (o, self.NP, DP)
            base = self.offsets["fc_predicate.weight"]
            c = np.arange(D)[:, None]
            p = np.arange(self.NP)[None, :]
            self.dstC[(base + c * self.NP + p).ravel()] = (o + p * DP + c).ravel()

    # ---- views -----------------------------------------------------------------------------------------------------------
    def linear_views(self, lin: Linear, packT, packF):
        W = packT[lin.w_off : lin.w_off + lin.Np * lin.Kp].view(lin.Np, lin.Kp)
        WT = packT[lin.wt_off : lin.wt_off + lin.Kp * lin.ldT].view(lin.Kp, lin.ldT)[:, : lin.Np]
        b = packF[lin.b_off : lin.b_off + lin.Np]
        return W, WT, b
