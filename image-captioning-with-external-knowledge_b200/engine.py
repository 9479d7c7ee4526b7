"""
Forward / hand-written backward orchestration of ``DecoderTransformer`` over the kernel set (kernels.CudaKernels).

Mirrors the reference's call stack: entity encoder -> fact encoder -> caption embedder -> entity /
fact Transformer encoders -> memory = [pixels; entities; facts] -> 3 post-LN decoder layers -> context indicators ->
vocabulary + pointer scores (G/models.py:315-361, K/models.py:457-514, N/models.py:440-497), in batch-major rows.
The backward pass is written out by hand (no autograd inside): every weight gradient is accumulated by the kernels
straight into one flat fp32 buffer laid out like the reference's parameters.

torch is used here for device memory and streams only; every arithmetic step is a kernel from include/ickb200.h.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace as NS
from typing import Dict, Optional

import torch

from .layout import HD, Linear, PackPlan, site_id

VARIANT_CODE = {"G": 0, "K": 1, "N": 2}


class LinearViews:
    def __init__(self, lin: Linear, W, WT, b, rowoff, colmap, biasoff):
        self.lin, self.W, self.WT, self.b = lin, W, WT, b
        self.rowoff, self.colmap, self.biasoff = rowoff, colmap, biasoff


class DecoderEngine:
    def __init__(self, plan: PackPlan, kernels, device, dtype: torch.dtype, pad: int, start: int, end: int,
                 p_dec: float = 0.5, p_enc: float = 0.5, p_pos: float = 0.1, max_len: int = 5000):
        self.plan, self.K, self.device, self.dtype = plan, kernels, device, dtype
        self.pad, self.start, self.end = pad, start, end
        self.p_dec, self.p_enc, self.p_pos = p_dec, p_enc, p_pos
        self.D, self.DP, self.H, self.L, self.dh = plan.D, plan.DP, plan.H, plan.L, plan.dh
        self.V, self.NP = plan.V, plan.NP
        self.variant = plan.variant
        self.has_facts = plan.has_facts
        self.P: Optional[torch.Tensor] = None
        self.packT = torch.zeros(plan.packT_size, dtype=dtype, device=device)
        self.packF = torch.zeros(plan.packF_size, dtype=torch.float32, device=device)
        self.dstA = torch.from_numpy(plan.dstA).to(device)
        self.dstB = torch.from_numpy(plan.dstB).to(device)
        self.dstC = torch.from_numpy(plan.dstC).to(device)
        self.lin: Dict[str, LinearViews] = {}
        for name, l in plan.linears.items():
            W, WT, b = plan.linear_views(l, self.packT, self.packF)
            self.lin[name] = LinearViews(l, W, WT, b, torch.from_numpy(l.rowoff).to(device), torch.from_numpy(l.colmap).to(device),
                                         torch.from_numpy(l.biasoff).to(device))
        o, r, c = plan.regions_T["word_embedding"]
        self.wemb = self.packT[o : o + r * c].view(r, c)
        if self.has_facts:
            o, r, c = plan.regions_F["fc_predicate_T"]
            self.WpT = self.packF[o : o + r * c].view(r, c)
        # sinusoidal table, PositionEncoder.__init__ (G/models.py:184-205); host-side constant
        pe = torch.zeros(max_len, self.D)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, self.D, 2).float() * (-math.log(10000.0) / self.D))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.pe = pe.to(device)

    # ---- parameters ----------------------------------------------------------------------------------------------------
    def attach(self, flat_params: torch.Tensor) -> None:
        assert flat_params.dtype == torch.float32 and flat_params.numel() == self.plan.n_params
        self.P = flat_params

    def param(self, name: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        buf = self.P if buf is None else buf
        off = self.plan.offsets[name]
        shp = self.plan.shapes[name]
        n = 1
        for s in shp:
            n *= s
        return buf[off : off + n].view(*shp)

    def off(self, name: str) -> int:
        return self.plan.offsets[name]

    def repack(self) -> None:
        """master fp32 parameters -> padded / transposed operand copies (one launch)."""
        self.K.adam_step(self.P, None, None, None, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, None, 1.0, self.dstA, self.dstB, self.dstC,
                         self.packT, self.packF, update=False)

    # ---- helpers ----------------------------------------------------------------------------------------------------------
    def _new(self, rows, cols, dtype=None):
        return torch.empty(rows, cols, dtype=dtype or self.dtype, device=self.device)

    def _newf(self, n):
        return torch.empty(n, dtype=torch.float32, device=self.device)

    def _drop(self, p, seed, name):
        return (p, seed, site_id(name)) if (p > 0.0 and seed is not None) else None

    # ---- Transformer encoder layer (self-attention over context slots) -----------------------------------------------------
    def _enc_layer_fwd(self, stack, l, x, B, S, p, seed, out=None, rowmap=(0, 0, 0)):
        K, DP, D, H, dh = self.K, self.DP, self.D, self.H, self.dh
        pre = f"{stack}.layers.{l}."
        site = f"{stack}.{l}"
        R = B * S
        qkv_l, out_l = self.lin[pre + "self_attn.qkv"], self.lin[pre + "self_attn.out"]
        f1, f2 = self.lin[pre + "ffn1"], self.lin[pre + "ffn2"]
        sv = NS(x=x)
        sv.qkv = self._new(R, 3 * DP)
        K.gemm(x, qkv_l.W, sv.qkv, bias=qkv_l.b)
        sv.o = self._new(R, DP)
        sv.lse = self._newf(B * H * S)
        K.mha_fwd(sv.qkv[:, :DP], sv.qkv[:, DP : 2 * DP], sv.qkv[:, 2 * DP :], sv.o, sv.lse, B, H, S, S, dh, causal=False,
                  drop=self._drop(p, seed, site + ".sa.attn"))
        sv.s1 = self._new(R, DP)
        sv.y1 = self._new(R, DP)
        sv.mean1, sv.rstd1 = self._newf(R), self._newf(R)
        # out-projection + residual + dropout + LayerNorm: one launch (kernels.gemm_add_ln)
        K.gemm_add_ln(sv.o, out_l.W, out_l.b, x, sv.s1, self.param(pre + "norm1.weight"), self.param(pre + "norm1.bias"), sv.y1, sv.mean1,
                      sv.rstd1, D, drop=self._drop(p, seed, site + ".d1"))
        sv.h1 = self._new(R, f1.lin.Np)
        K.gemm(sv.y1, f1.W, sv.h1, bias=f1.b, epi=1, drop=self._drop(p, seed, site + ".ffn"))
        sv.s2 = self._new(R, DP)
        y2 = out if out is not None else self._new(R, DP)
        sv.mean2, sv.rstd2 = self._newf(R), self._newf(R)
        if rowmap[0] == 0:
            K.gemm_add_ln(sv.h1, f2.W, f2.b, sv.y1, sv.s2, self.param(pre + "norm2.weight"), self.param(pre + "norm2.bias"), y2, sv.mean2,
                          sv.rstd2, D, drop=self._drop(p, seed, site + ".d2"))
        else:  # row-mapped output (the last layer writes into the memory buffer): separate LayerNorm kernel
            K.gemm(sv.h1, f2.W, sv.s2, bias=f2.b)
            K.add_ln_fwd(sv.y1, sv.s2, self.param(pre + "norm2.weight"), self.param(pre + "norm2.bias"), y2, sv.mean2, sv.rstd2, D,
                         rowmap=rowmap, drop=self._drop(p, seed, site + ".d2"))
        return y2, sv

    # The weight gradients of a layer do not feed its input gradient, so the backward helpers only RECORD them
    # (wg.append((dY, X, rowoff, colmap, biasoff))) and the layer issues them together at its end (kernels.wgrad_group: one
    # grouped tensor-core launch instead of 4-6 small ones).  Every recorded dY therefore needs its own buffer.
    @staticmethod
    def _wg(wg, dY, X, lin):
        wg.append((dY, X, lin.rowoff, lin.colmap, lin.biasoff))

    def _ffn_bwd(self, pre, site, f1, f2, dB, h1, y_in, dA, wg, p, seed):
        """dB = grad of the FFN output; accumulates the FFN input gradient into dA and records the weight grads in wg."""
        K = self.K
        R = dB.shape[0]
        self._wg(wg, dB, h1, f2)
        dh1 = self._new(R, f1.lin.Np)
        K.gemm(dB, f2.WT, dh1, aux=h1, epi=2, drop=(p if seed is not None else 0.0, seed or 0, 0))
        self._wg(wg, dh1, y_in, f1)
        K.gemm(dh1, f1.WT, dA, accumulate=True)

    def _self_attn_bwd(self, pre, site, qkv_l, out_l, dB, sv_qkv, sv_o, sv_lse, x_in, dC, B, S, causal, wg, p, seed):
        """dB = grad of the out-projection output; accumulates the block-input gradient into dC."""
        K, DP, H, dh = self.K, self.DP, self.H, self.dh
        R = B * S
        self._wg(wg, dB, sv_o, out_l)
        dO = self._new(R, DP)
        dqkv = self._new(R, 3 * DP)
        dsum = self._newf(B * H * S)
        ready = K.gemm_rowdot(dB, out_l.WT, dO, sv_o, dsum, S, H)  # dO and rowsum(dO * O) from one launch
        K.mha_bwd(sv_qkv[:, :DP], sv_qkv[:, DP : 2 * DP], sv_qkv[:, 2 * DP :], sv_o, dO, sv_lse, dsum, dqkv[:, :DP],
                  dqkv[:, DP : 2 * DP], dqkv[:, 2 * DP :], B, H, S, S, dh, causal=causal, drop=self._drop(p, seed, site + ".sa.attn"),
                  dsum_ready=ready)
        self._wg(wg, dqkv, x_in, qkv_l)
        K.gemm(dqkv, qkv_l.WT, dC, accumulate=True)

    def _enc_layer_bwd(self, stack, l, sv, dy, dy_rowmap, B, S, gflat, p, seed):
        K, DP, D = self.K, self.DP, self.D
        pre = f"{stack}.layers.{l}."
        site = f"{stack}.{l}"
        R = B * S
        qkv_l, out_l = self.lin[pre + "self_attn.qkv"], self.lin[pre + "self_attn.out"]
        f1, f2 = self.lin[pre + "ffn1"], self.lin[pre + "ffn2"]
        wg = []
        dA, dB2 = self._new(R, DP), self._new(R, DP)
        K.add_ln_bwd(dy, sv.s2, sv.mean2, sv.rstd2, self.param(pre + "norm2.weight"), dA, dB2, self.param(pre + "norm2.weight", gflat),
                     self.param(pre + "norm2.bias", gflat), D, rowmap=dy_rowmap, drop=self._drop(p, seed, site + ".d2"))
        self._ffn_bwd(pre, site, f1, f2, dB2, sv.h1, sv.y1, dA, wg, p, seed)
        dC, dB1 = self._new(R, DP), self._new(R, DP)
        K.add_ln_bwd(dA, sv.s1, sv.mean1, sv.rstd1, self.param(pre + "norm1.weight"), dC, dB1, self.param(pre + "norm1.weight", gflat),
                     self.param(pre + "norm1.bias", gflat), D, drop=self._drop(p, seed, site + ".d1"))
        self._self_attn_bwd(pre + "self_attn.", site, qkv_l, out_l, dB1, sv.qkv, sv.o, sv.lse, sv.x, dC, B, S, False, wg, p, seed)
        K.wgrad_group(wg, gflat)
        return dC

    # ---- entity + fact encoder stacks in lockstep ---------------------------------------------------------------------------------
    # The two nn.TransformerEncoder stacks (entities K/models.py:321-322,495; facts :323-324,496) have the same layer shapes and
    # different parameters.  Their rows live in ONE buffer [entity rows | pad to 128 | fact rows]; every projection of layer l is
    # a single dual-weight GEMM launch (kernels.gemm_dual), so the fact-sized GEMMs (51 row tiles, pure launch latency on
    # their own) ride along with the entity-sized ones; attention and LayerNorm run per stack on row slices, and the eight
    # weight gradients of a layer pair form one grouped wgrad launch.
    ENC_STACKS = ("transformer_encoder_entities", "transformer_encoder_facts")

    def _dual_geometry(self, B, E, F):
        Re, Rf = B * E, B * F
        off = (Re + 127) // 128 * 128
        return Re, Rf, off, off + Rf

    def _dual_slices(self, B, E, F, P):
        Re, Rf, off, Rc = self._dual_geometry(B, E, F)
        # (stack, tokens per image, rows, first row in the concatenated buffer, first memory slot)
        return ((self.ENC_STACKS[0], E, Re, 0, P), (self.ENC_STACKS[1], F, Rf, off, P + E)), off, Re, Rc

    def _enc_dual_fwd(self, xcat, mem, B, E, F, P, M, p, seed):
        K, DP, D, H, dh, L = self.K, self.DP, self.D, self.H, self.dh, self.L
        stacks, off, Re, Rc = self._dual_slices(B, E, F, P)
        saves = []
        x = xcat
        for l in range(L):
            last = l == L - 1
            pre = [f"{st[0]}.layers.{l}." for st in stacks]
            site = [f"{st[0]}.{l}" for st in stacks]
            lin = lambda n: (self.lin[pre[0] + n], self.lin[pre[1] + n])  # noqa: E731
            qkv_l, out_l, f1, f2 = lin("self_attn.qkv"), lin("self_attn.out"), lin("ffn1"), lin("ffn2")
            sv = NS(x=x)
            sv.qkv = self._new(Rc, 3 * DP)
            K.gemm_dual(x, qkv_l[0].W, qkv_l[1].W, sv.qkv, off, Re, qkv_l[0].b, qkv_l[1].b)
            sv.o = self._new(Rc, DP)
            sv.lse = []
            for i, (_, S, R, r0, _) in enumerate(stacks):
                lse = self._newf(B * H * S)
                q = sv.qkv[r0 : r0 + R]
                K.mha_fwd(q[:, :DP], q[:, DP : 2 * DP], q[:, 2 * DP :], sv.o[r0 : r0 + R], lse, B, H, S, S, dh, causal=False,
                          drop=self._drop(p, seed, site[i] + ".sa.attn"))
                sv.lse.append(lse)
            sv.s1 = self._new(Rc, DP)
            sv.y1 = self._new(Rc, DP)
            sv.mean1, sv.rstd1, sv.mean2, sv.rstd2 = (self._newf(Rc) for _ in range(4))
            Rf = stacks[1][2]
            par = lambda n: (self.param(pre[0] + n), self.param(pre[1] + n))  # noqa: E731
            # out-projection + residual + dropout + LayerNorm of both stacks: one launch
            K.gemm_add_ln_dual(sv.o, out_l[0].W, out_l[1].W, out_l[0].b, out_l[1].b, x, sv.s1, sv.y1, sv.mean1, sv.rstd1, D, off, Re,
                               par("norm1.weight"), par("norm1.bias"),
                               drops=(self._drop(p, seed, site[0] + ".d1"), self._drop(p, seed, site[1] + ".d1")))
            sv.h1 = self._new(Rc, f1[0].lin.Np)
            K.gemm_dual(sv.y1, f1[0].W, f1[1].W, sv.h1, off, Re, f1[0].b, f1[1].b, epi=1, drop0=self._drop(p, seed, site[0] + ".ffn"),
                        drop1=self._drop(p, seed, site[1] + ".ffn"))
            sv.s2 = self._new(Rc, DP)
            y2 = None if last else self._new(Rc, DP)
            drops2 = (self._drop(p, seed, site[0] + ".d2"), self._drop(p, seed, site[1] + ".d2"))
            if not last:
                K.gemm_add_ln_dual(sv.h1, f2[0].W, f2[1].W, f2[0].b, f2[1].b, sv.y1, sv.s2, y2, sv.mean2, sv.rstd2, D, off, Re,
                                   par("norm2.weight"), par("norm2.bias"), drops=drops2)
            else:
                # the last layer writes straight into the decoder's memory rows (the reference concatenates, K/models.py:497-499):
                # a row-mapped output, which the separate LayerNorm kernel provides
                K.gemm_dual(sv.h1, f2[0].W, f2[1].W, sv.s2, off, Re, f2[0].b, f2[1].b)
                maps = tuple((S, M, m0) for (_, S, _, _, m0) in stacks)
                K.add_ln_fwd_dual(sv.y1, sv.s2, mem, sv.mean2, sv.rstd2, D, Re, Rf, off, par("norm2.weight"), par("norm2.bias"),
                                  rowmaps=maps, drops=drops2)
            x = y2
            saves.append(sv)
        return saves

    def _enc_dual_bwd(self, saves, dmem, B, E, F, P, M, gflat, p, seed, done=lambda region: None):
        """dmem: gradient of the memory buffer.  Returns the gradient of the concatenated stack input (entity rows | pad | fact rows)."""
        K, DP, D, H, dh, L = self.K, self.DP, self.D, self.H, self.dh, self.L
        stacks, off, Re, Rc = self._dual_slices(B, E, F, P)
        ep = (p if seed is not None else 0.0, seed or 0, 0)
        d = None
        for l in reversed(range(L)):
            last = l == L - 1
            sv = saves[l]
            pre = [f"{st[0]}.layers.{l}." for st in stacks]
            site = [f"{st[0]}.{l}" for st in stacks]
            lin = lambda n: (self.lin[pre[0] + n], self.lin[pre[1] + n])  # noqa: E731
            qkv_l, out_l, f1, f2 = lin("self_attn.qkv"), lin("self_attn.out"), lin("ffn1"), lin("ffn2")
            wg = []
            dA, dB2 = self._new(Rc, DP), self._new(Rc, DP)
            Rf = stacks[1][2]
            par = lambda n: (self.param(pre[0] + n), self.param(pre[1] + n))  # noqa: E731
            gpar = lambda n: (self.param(pre[0] + n, gflat), self.param(pre[1] + n, gflat))  # noqa: E731
            maps = tuple((S, M, m0) if last else (0, 0, 0) for (_, S, _, _, m0) in stacks)
            K.add_ln_bwd_dual(dmem if last else d, sv.s2, sv.mean2, sv.rstd2, dA, dB2, D, Re, Rf, off, par("norm2.weight"), gpar("norm2.weight"),
                              gpar("norm2.bias"), rowmaps=maps,
                              drops=(self._drop(p, seed, site[0] + ".d2"), self._drop(p, seed, site[1] + ".d2")))
            for i, (_, S, R, r0, m0) in enumerate(stacks):
                sl = slice(r0, r0 + R)
                self._wg(wg, dB2[sl], sv.h1[sl], f2[i])
            dh1 = self._new(Rc, f1[0].lin.Np)
            K.gemm_dual(dB2, f2[0].WT, f2[1].WT, dh1, off, Re, aux=sv.h1, epi=2, drop0=ep, drop1=ep)
            K.gemm_dual(dh1, f1[0].WT, f1[1].WT, dA, off, Re, accumulate=True)
            dC, dB1 = self._new(Rc, DP), self._new(Rc, DP)
            K.add_ln_bwd_dual(dA, sv.s1, sv.mean1, sv.rstd1, dC, dB1, D, Re, Rf, off, par("norm1.weight"), gpar("norm1.weight"), gpar("norm1.bias"),
                              drops=(self._drop(p, seed, site[0] + ".d1"), self._drop(p, seed, site[1] + ".d1")))
            for i, (_, S, R, r0, _) in enumerate(stacks):
                sl = slice(r0, r0 + R)
                self._wg(wg, dh1[sl], sv.y1[sl], f1[i])
                self._wg(wg, dB1[sl], sv.o[sl], out_l[i])
            dO = self._new(Rc, DP)
            dsums = [self._newf(B * H * S) for (_, S, _, _, _) in stacks]
            # dO of both stacks and their rowsum(dO * O) terms from one launch
            ready = K.gemm_rowdot(dB1, out_l[0].WT, dO, sv.o, dsums[0], stacks[0][1], H, W1=out_l[1].WT, m_split=off, rows0=Re,
                                  dsum1=dsums[1], S1=stacks[1][1])
            dqkv = self._new(Rc, 3 * DP)
            for i, (_, S, R, r0, _) in enumerate(stacks):
                sl = slice(r0, r0 + R)
                q, dq = sv.qkv[sl], dqkv[sl]
                K.mha_bwd(q[:, :DP], q[:, DP : 2 * DP], q[:, 2 * DP :], sv.o[sl], dO[sl], sv.lse[i], dsums[i], dq[:, :DP], dq[:, DP : 2 * DP],
                          dq[:, 2 * DP :], B, H, S, S, dh, causal=False, drop=self._drop(p, seed, site[i] + ".sa.attn"), dsum_ready=ready)
                self._wg(wg, dq, sv.x[sl], qkv_l[i])
            K.gemm_dual(dqkv, qkv_l[0].WT, qkv_l[1].WT, dC, off, Re, accumulate=True)
            K.wgrad_group(wg, gflat)
            done(f"enc{l}")
            d = dC
        return d

    # ---- Transformer decoder layer -------------------------------------------------------------------------------------------
    def _dec_layer_fwd(self, l, x, kv, B, T, M, p, seed):
        K, DP, D, H, dh = self.K, self.DP, self.D, self.H, self.dh
        pre = f"transformer_decoder.layers.{l}."
        site = f"transformer_decoder.{l}"
        R = B * T
        qkv_l, out_l = self.lin[pre + "self_attn.qkv"], self.lin[pre + "self_attn.out"]
        q_l, out2_l = self.lin[pre + "multihead_attn.q"], self.lin[pre + "multihead_attn.out"]
        f1, f2 = self.lin[pre + "ffn1"], self.lin[pre + "ffn2"]
        sv = NS(x=x)
        # causal self-attention
        sv.qkv = self._new(R, 3 * DP)
        K.gemm(x, qkv_l.W, sv.qkv, bias=qkv_l.b)
        sv.o = self._new(R, DP)
        sv.lse = self._newf(B * H * T)
        K.mha_fwd(sv.qkv[:, :DP], sv.qkv[:, DP : 2 * DP], sv.qkv[:, 2 * DP :], sv.o, sv.lse, B, H, T, T, dh, causal=True,
                  drop=self._drop(p, seed, site + ".sa.attn"))
        sv.s1 = self._new(R, DP)
        sv.y1 = self._new(R, DP)
        sv.mean1, sv.rstd1 = self._newf(R), self._newf(R)
        K.gemm_add_ln(sv.o, out_l.W, out_l.b, x, sv.s1, self.param(pre + "norm1.weight"), self.param(pre + "norm1.bias"), sv.y1, sv.mean1,
                      sv.rstd1, D, drop=self._drop(p, seed, site + ".d1"))
        # cross-attention over the memory (its K/V projections were computed for all layers at once)
        sv.q = self._new(R, DP)
        K.gemm(sv.y1, q_l.W, sv.q, bias=q_l.b)
        sv.k = kv[:, l * 2 * DP : l * 2 * DP + DP]
        sv.v = kv[:, l * 2 * DP + DP : (l + 1) * 2 * DP]
        sv.o2 = self._new(R, DP)
        sv.lse2 = self._newf(B * H * T)
        K.mha_fwd(sv.q, sv.k, sv.v, sv.o2, sv.lse2, B, H, T, M, dh, causal=False, drop=self._drop(p, seed, site + ".ca.attn"))
        sv.s2 = self._new(R, DP)
        sv.y2 = self._new(R, DP)
        sv.mean2, sv.rstd2 = self._newf(R), self._newf(R)
        K.gemm_add_ln(sv.o2, out2_l.W, out2_l.b, sv.y1, sv.s2, self.param(pre + "norm2.weight"), self.param(pre + "norm2.bias"), sv.y2,
                      sv.mean2, sv.rstd2, D, drop=self._drop(p, seed, site + ".d2"))
        # feed-forward
        sv.h1 = self._new(R, f1.lin.Np)
        K.gemm(sv.y2, f1.W, sv.h1, bias=f1.b, epi=1, drop=self._drop(p, seed, site + ".ffn"))
        sv.s3 = self._new(R, DP)
        y3 = self._new(R, DP)
        sv.mean3, sv.rstd3 = self._newf(R), self._newf(R)
        K.gemm_add_ln(sv.h1, f2.W, f2.b, sv.y2, sv.s3, self.param(pre + "norm3.weight"), self.param(pre + "norm3.bias"), y3, sv.mean3,
                      sv.rstd3, D, drop=self._drop(p, seed, site + ".d3"))
        return y3, sv

    def _dec_layer_bwd(self, l, sv, dy, dkv, B, T, M, gflat, p, seed):
        K, DP, D, H, dh = self.K, self.DP, self.D, self.H, self.dh
        pre = f"transformer_decoder.layers.{l}."
        site = f"transformer_decoder.{l}"
        R = B * T
        qkv_l, out_l = self.lin[pre + "self_attn.qkv"], self.lin[pre + "self_attn.out"]
        q_l, out2_l = self.lin[pre + "multihead_attn.q"], self.lin[pre + "multihead_attn.out"]
        f1, f2 = self.lin[pre + "ffn1"], self.lin[pre + "ffn2"]
        g = lambda n: self.param(pre + n, gflat)  # noqa: E731
        w = lambda n: self.param(pre + n)  # noqa: E731
        wg = []
        dA, dB3 = self._new(R, DP), self._new(R, DP)
        K.add_ln_bwd(dy, sv.s3, sv.mean3, sv.rstd3, w("norm3.weight"), dA, dB3, g("norm3.weight"), g("norm3.bias"), D,
                     drop=self._drop(p, seed, site + ".d3"))
        self._ffn_bwd(pre, site, f1, f2, dB3, sv.h1, sv.y2, dA, wg, p, seed)
        dC, dB2 = self._new(R, DP), self._new(R, DP)
        K.add_ln_bwd(dA, sv.s2, sv.mean2, sv.rstd2, w("norm2.weight"), dC, dB2, g("norm2.weight"), g("norm2.bias"), D,
                     drop=self._drop(p, seed, site + ".d2"))
        # cross-attention backward
        self._wg(wg, dB2, sv.o2, out2_l)
        dO2 = self._new(R, DP)
        dq = self._new(R, DP)
        dsum = self._newf(B * H * T)
        ready = K.gemm_rowdot(dB2, out2_l.WT, dO2, sv.o2, dsum, T, H)
        K.mha_bwd(sv.q, sv.k, sv.v, sv.o2, dO2, sv.lse2, dsum, dq, dkv[:, l * 2 * DP : l * 2 * DP + DP],
                  dkv[:, l * 2 * DP + DP : (l + 1) * 2 * DP], B, H, T, M, dh, causal=False, drop=self._drop(p, seed, site + ".ca.attn"),
                  dsum_ready=ready)
        self._wg(wg, dq, sv.y1, q_l)
        K.gemm(dq, q_l.WT, dC, accumulate=True)
        dE, dB1 = self._new(R, DP), self._new(R, DP)
        K.add_ln_bwd(dC, sv.s1, sv.mean1, sv.rstd1, w("norm1.weight"), dE, dB1, g("norm1.weight"), g("norm1.bias"), D,
                     drop=self._drop(p, seed, site + ".d1"))
        self._self_attn_bwd(pre + "self_attn.", site, qkv_l, out_l, dB1, sv.qkv, sv.o, sv.lse, sv.x, dE, B, T, True, wg, p, seed)
        K.wgrad_group(wg, gflat)
        return dE

    # ---- context encoders -------------------------------------------------------------------------------------------------------
    def _encode_context(self, inp, B, E, F):
        K, D, DP = self.K, self.D, self.DP
        if self.has_facts:
            # entity rows, padding up to a multiple of 128 rows, fact rows: the layout the lockstep encoder stacks work on
            Re, Rf, off, Rc = self._dual_geometry(B, E, F)
            self._xcat = self._new(Rc, DP)
            ent_enc = self._xcat[:Re]
        else:
            ent_enc = self._new(B * E, DP)
        K.entity_encode_fwd(inp.entities, inp.facts, self.param("entity_encoder.type_embedding.weight"),
                            self.wemb if self.variant == "N" else None, ent_enc, VARIANT_CODE[self.variant], B, E, F, D,
                            self.plan.shapes["entity_encoder.type_embedding.weight"][0], self.V)
        fact_enc = None
        if self.has_facts:
            fact_enc = self._xcat[off : off + Rf]
            K.fact_encode_fwd(inp.facts, ent_enc, self.param("predicate_embedding.weight"), fact_enc, B, E, F, D, self.NP)
        return ent_enc, fact_enc

    def _build_memory(self, inp, ent_enc, fact_enc, B, E, F, P, M, p_enc, seed):
        K, D, DP, L = self.K, self.D, self.DP, self.L
        mem = self._new(B * M, DP)
        K.pixels_fwd(inp.encoder_out, mem, B, D, P, M)
        if self.has_facts:
            xcat, self._xcat = self._xcat, None
            return mem, {"dual": self._enc_dual_fwd(xcat, mem, B, E, F, P, M, p_enc, seed)}
        saves = {}
        stacks = [("transformer_encoder_entities", ent_enc, E, P)]
        if self.has_facts:
            stacks.append(("transformer_encoder_facts", fact_enc, F, P + E))
        for stack, x, S, off in stacks:
            svs = []
            for l in range(L):
                last = l == L - 1
                x, sv = self._enc_layer_fwd(stack, l, x, B, S, p_enc, seed, out=mem if last else None,
                                            rowmap=(S, M, off) if last else (0, 0, 0))
                svs.append(sv)
            saves[stack] = svs
        return mem, saves

    # ---- forward -------------------------------------------------------------------------------------------------------------------
    def forward(self, inp, train: bool = False, seed: Optional[int] = None):
        """
        inp: namespace of DEVICE tensors already sorted by decreasing length (the module does the sort):
          captions (B,T) i64, caption_masks (B,T) i64, encoder_out (B,D,P) f32, entities (B,E,C) f32, facts (B,F,3) i64|None.
        Returns (scores (B,T,W) fp32, ctx for backward).
        """
        K, D, DP, L = self.K, self.D, self.DP, self.L
        B, T = inp.captions.shape
        E = inp.entities.shape[1]
        F = inp.facts.shape[1] if self.has_facts else 0
        P = inp.encoder_out.shape[2]
        M = P + E + F
        W = self.V + E + F
        p_dec = self.p_dec if train else 0.0
        p_enc = self.p_enc if train else 0.0
        p_pos = self.p_pos if train else 0.0
        if not train:
            seed = None
        ctx = NS(inp=inp, B=B, T=T, E=E, F=F, P=P, M=M, W=W, seed=seed, p_dec=p_dec, p_enc=p_enc, p_pos=p_pos)
        ctx.ent_enc, ctx.fact_enc = self._encode_context(inp, B, E, F)
        x0 = self._new(B * T, DP)
        K.caption_embed_fwd(inp.captions, inp.caption_masks, self.wemb, ctx.ent_enc, ctx.fact_enc, self.pe, x0, B, T, 0, T, self.V, E, F,
                            D, self.pad, math.sqrt(D), drop=self._drop(p_pos, seed, "pos"))
        ctx.mem, ctx.enc_saves = self._build_memory(inp, ctx.ent_enc, ctx.fact_enc, B, E, F, P, M, p_enc, seed)
        kv_l = self.lin["transformer_decoder.kv_all"]
        ctx.kv = self._new(B * M, kv_l.lin.Np)
        K.gemm(ctx.mem, kv_l.W, ctx.kv, bias=kv_l.b)
        x = x0
        ctx.dec_saves = []
        for l in range(L):
            x, sv = self._dec_layer_fwd(l, x, ctx.kv, B, T, M, p_dec, seed)
            ctx.dec_saves.append(sv)
        ctx.h = x
        scores = torch.empty(B, T, W, dtype=torch.float32, device=self.device)
        s2 = scores.view(B * T, W)
        self._heads_fwd(ctx, inp.captions, x, s2, B, T, 0, T, E, F, lag=0)
        return scores, ctx

    def _heads_fwd(self, ctx, captions, h, s2, B, Tn, t0, Tcap, E, F, lag, group=1):
        """get_scores (K/models.py:420-455): gate + vocabulary GEMM + entity / fact pointer scores into one buffer.
        group > 1 (beam search): that many consecutive rows of captions / h / s2 share one image context."""
        K, D, DP, V = self.K, self.D, self.DP, self.V
        fv = self.lin["fc_vocab"]
        if self.has_facts:
            ctx.first_t = torch.empty(B * F, dtype=torch.int32, device=self.device)
            ctx.tmin = torch.empty(B * F, dtype=torch.int32, device=self.device)
            K.fact_first_mention(captions, ctx.inp.facts, ctx.first_t, ctx.tmin, B, Tcap, F, V, E, group=group, NP=self.NP)
            ctx.gate = self._new(B * Tn, DP)
            ctx.hg = self._new(B * Tn, DP)
            K.pred_gate_fwd(ctx.tmin, ctx.inp.facts, self.WpT, self.param("fc_predicate.bias"), h, ctx.gate, ctx.hg, B, Tn, t0, F, D,
                            self.NP, lag, group=group)
            vin = ctx.hg
        else:
            vin = h
        K.gemm(vin, fv.W, s2[:, :V], bias=fv.b)
        K.pointer_fwd(h, ctx.ent_enc, self.param("fc_entity.weight"), self.param("fc_entity.bias"), None, s2, B, Tn, t0, E, D, V, lag,
                      group=group)
        if self.has_facts:
            K.pointer_fwd(h, ctx.fact_enc, self.param("fc_fact.weight"), self.param("fc_fact.bias"), ctx.first_t, s2, B, Tn, t0, F, D,
                          V + E, lag, group=group)

    # ---- loss ------------------------------------------------------------------------------------------------------------------------
    def loss(self, scores, captions_sorted, decode_len_dev, want_grad: bool = True, loss_acc=None):
        """pack_padded_sequence + CrossEntropyLoss(ignore_index=pad) (G/train.py:275-281) fused with its gradient.
        Returns (loss_acc = [sum of row losses, kept rows] fp32 on device, dscores (B*T, ldW) activation dtype, UNSCALED)."""
        B, T, W = scores.shape
        ldW = (W + 7) // 8 * 8
        if loss_acc is None:
            loss_acc = torch.zeros(2, dtype=torch.float32, device=self.device)
        ds = self._new(B * T, ldW) if want_grad else None
        self.K.ce(scores.view(B * T, W), captions_sorted, decode_len_dev, loss_acc, ds, B, T, W, self.pad)
        return loss_acc, ds

    # ---- backward --------------------------------------------------------------------------------------------------------------------
    def grad_regions(self):
        """Contiguous slices (lo, hi) of the flat gradient buffer in the order the backward pass COMPLETES them - what a data-parallel
        trainer can all-reduce while the rest of the backward still runs (see backward(on_done=...)):
          "heads"     fc_vocab .. fc_predicate (the tail of the buffer)          done after the score heads
          "decoder"   the three decoder layers                                   done after the decoder stack
          "word"      word_embedding                                             done after the caption embedder (G, K; N: at the end)
          "enc{l}"    layer l of the entity / fact encoder stacks (two slices)   done after that lock-step layer
          "rest"      the remaining embeddings (N: word_embedding too)           done at the end"""
        off, n = self.plan.offsets, self.plan.n_params
        names = list(self.plan.shapes.keys())

        def span(pred):
            ks = [k for k in names if pred(k)]
            lo = min(off[k] for k in ks)
            hi = max(off[k] + int(math.prod(self.plan.shapes[k])) for k in ks)
            return (lo, hi)

        reg = {"heads": [span(lambda k: k.startswith("fc_"))], "decoder": [span(lambda k: k.startswith("transformer_decoder."))]}
        word = span(lambda k: k == "word_embedding.weight")
        rest = span(lambda k: k in ("entity_encoder.type_embedding.weight", "predicate_embedding.weight"))
        if self.variant == "N":
            rest = (min(word[0], rest[0]), max(word[1], rest[1]))
        else:
            reg["word"] = [word]
        reg["rest"] = [rest]
        for l in range(self.L):
            sl = [span(lambda k, l=l: k.startswith(f"transformer_encoder_entities.layers.{l}."))]
            if self.has_facts:
                sl.append(span(lambda k, l=l: k.startswith(f"transformer_encoder_facts.layers.{l}.")))
            reg[f"enc{l}"] = sl
        covered = sorted(s for v in reg.values() for s in v)
        assert covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:])), covered
        return reg

    def backward(self, ctx, dscores, gflat, need_encoder_grad: bool = False, on_done=None):
        """
        dscores: (B*T, ld>=W) in the activation dtype (pad columns ignored).  Accumulates every parameter gradient into
        gflat (flat fp32, reference parameter order) and returns d encoder_out (B,D,P) fp32 or None.
        on_done(region): called on the launching thread as soon as every kernel that writes the gradient region (grad_regions())
        has been LAUNCHED - the hook where a data-parallel trainer starts that region's all-reduce on a side stream.
        """
        done = on_done if on_done is not None else (lambda region: None)
        K, D, DP, L, V = self.K, self.D, self.DP, self.L, self.V
        B, T, E, F, P, M = ctx.B, ctx.T, ctx.E, ctx.F, ctx.P, ctx.M
        inp, seed = ctx.inp, ctx.seed
        R = B * T
        dEnt = torch.zeros(B * E, DP, dtype=torch.float32, device=self.device)
        dFact = torch.zeros(B * F, DP, dtype=torch.float32, device=self.device) if self.has_facts else None
        fv = self.lin["fc_vocab"]
        dSv = dscores[:, :V]
        dh = self._new(R, DP)
        if self.has_facts:
            dhg = self._new(R, DP)
            K.gemm(dSv, fv.WT, dhg)
            K.wgrad(dSv, ctx.hg, gflat, fv.rowoff, fv.colmap, fv.biasoff)
            dG = self._new(R, DP)
            K.gate_mul_bwd(dhg, ctx.h, ctx.gate, dG, dh)
            K.pred_gate_bwd(dG, ctx.tmin, inp.facts, gflat, self.off("fc_predicate.weight"), B, T, F, D, self.NP, 0)
            K.colsum(dG, self.param("fc_predicate.bias", gflat), D)
        else:
            K.gemm(dSv, fv.WT, dh)
            K.wgrad(dSv, ctx.h, gflat, fv.rowoff, fv.colmap, fv.biasoff)
        K.pointer_bwd(dscores, ctx.h, ctx.ent_enc, self.param("fc_entity.weight"), None, dEnt, dh, gflat, self.off("fc_entity.weight"),
                      self.off("fc_entity.bias"), B, T, E, D, V, 0)
        if self.has_facts:
            K.pointer_bwd(dscores, ctx.h, ctx.fact_enc, self.param("fc_fact.weight"), ctx.first_t, dFact, dh, gflat,
                          self.off("fc_fact.weight"), self.off("fc_fact.bias"), B, T, F, D, V + E, 0)
        done("heads")
        # decoder layers
        kv_l = self.lin["transformer_decoder.kv_all"]
        dkv = self._new(B * M, kv_l.lin.Np)
        dx = dh
        for l in reversed(range(L)):
            dx = self._dec_layer_bwd(l, ctx.dec_saves[l], dx, dkv, B, T, M, gflat, ctx.p_dec, seed)
        # memory K/V projections (all layers at once): parameters of the decoder layers (multihead_attn.in_proj rows 300..899)
        K.wgrad(dkv, ctx.mem, gflat, kv_l.rowoff, kv_l.colmap, kv_l.biasoff)
        done("decoder")
        K.caption_embed_bwd(dx, inp.captions, inp.caption_masks, dEnt, dFact, gflat, self.off("word_embedding.weight"), B, T, V, E, F, D,
                            self.pad, math.sqrt(D), drop=self._drop(ctx.p_pos, seed, "pos"))
        if self.variant != "N":
            done("word")  # (N: the entity encoder adds name-word gradients at the very end)
        dmem = self._new(B * M, DP)
        K.gemm(dkv, kv_l.WT, dmem)
        d_enc = None
        if need_encoder_grad:
            d_enc = torch.empty(B, D, P, dtype=torch.float32, device=self.device)
            K.pixels_bwd(dmem, d_enc, B, D, P, M)
        # context encoders; the fact encodings' input gradient also flows into the entity encodings (FactEncoder gathers them)
        if self.has_facts:
            Re, Rf, off, Rc = self._dual_geometry(B, E, F)
            dcat = self._enc_dual_bwd(ctx.enc_saves["dual"], dmem, B, E, F, P, M, gflat, ctx.p_enc, seed, done)
            K.accum_f32(dcat[off : off + Rf], dFact)
            K.fact_encode_bwd(dFact, inp.facts, dEnt, gflat, self.off("predicate_embedding.weight"), B, E, F, D, self.NP)
            d = dcat[:Re]
        else:
            d = dmem
            rowmap = (E, M, P)
            for l in reversed(range(L)):
                d = self._enc_layer_bwd("transformer_encoder_entities", l, ctx.enc_saves["transformer_encoder_entities"][l], d, rowmap, B,
                                        E, gflat, ctx.p_enc, seed)
                rowmap = (0, 0, 0)
                done(f"enc{l}")
        K.accum_f32(d, dEnt)
        K.entity_encode_bwd(dEnt, inp.entities, inp.facts, self.param("entity_encoder.type_embedding.weight"),
                            self.wemb if self.variant == "N" else None, gflat, self.off("entity_encoder.type_embedding.weight"),
                            self.off("word_embedding.weight"), 0 if self.dtype == torch.float32 else 1, VARIANT_CODE[self.variant], B, E, F,
                            D, self.plan.shapes["entity_encoder.type_embedding.weight"][0], self.V)
        done("rest")
        return d_enc

    def _memory_kv(self, mem, rows):
        """Cross-attention K / V of the memory for the decode loops -> [(K_l, V_l, row stride)] per decoder layer.  One buffer of
        K|V rows PER LAYER (the training path projects all layers in one [rows, L*2*DP] GEMM): the cached cross-attention of a
        layer then streams sequential memory instead of 1280-byte pieces 3840 bytes apart (ICK_DECODE_KV_SPLIT=0: old layout)."""
        K, DP, L = self.K, self.DP, self.L
        kv_l = self.lin["transformer_decoder.kv_all"]
        if os.environ.get("ICK_DECODE_KV_SPLIT", "1") == "0":
            kvw = kv_l.lin.Np
            kv = self._new(rows, kvw)
            K.gemm(mem, kv_l.W, kv, bias=kv_l.b)
            return [(kv[:, l * 2 * DP : l * 2 * DP + DP], kv[:, l * 2 * DP + DP : (l + 1) * 2 * DP], kvw) for l in range(L)]
        out = []
        for l in range(L):
            kvl = self._new(rows, 2 * DP)
            K.gemm(mem, kv_l.W[l * 2 * DP : (l + 1) * 2 * DP], kvl, bias=kv_l.b[l * 2 * DP : (l + 1) * 2 * DP])
            out.append((kvl[:, :DP], kvl[:, DP:], 2 * DP))
        return out

    def _dec_step_layer(self, l, x, row, next_row, b, attn_self, attn_cross, rows, mean, rstd, fused):
        """One decoder layer of a single-position decode step.  `row` holds this position's q|k|v (already projected when
        `fused`), `next_row` is where the NEXT layer's q|k|v goes (None for the last layer).  attn_self(q_row, out) and
        attn_cross(q, out) launch the cached attentions.  fused (bf16, opt-in: ICK_DECODE_CHAIN=1): the row-wise work between two
        attentions is one ick_decode_chain launch (4 launches per layer instead of 11); else the GEMM / LayerNorm kernels one by one."""
        K, D, DP = self.K, self.D, self.DP
        pre = f"transformer_decoder.layers.{l}."
        qkv_l, out_l = self.lin[pre + "self_attn.qkv"], self.lin[pre + "self_attn.out"]
        q_l, out2_l = self.lin[pre + "multihead_attn.q"], self.lin[pre + "multihead_attn.out"]
        f1, f2 = self.lin[pre + "ffn1"], self.lin[pre + "ffn2"]
        n1 = (self.param(pre + "norm1.weight"), self.param(pre + "norm1.bias"))
        n2 = (self.param(pre + "norm2.weight"), self.param(pre + "norm2.bias"))
        n3 = (self.param(pre + "norm3.weight"), self.param(pre + "norm3.bias"))
        if not fused:
            K.gemm(x, qkv_l.W, row, bias=qkv_l.b)
            attn_self(row, b.o)
            K.gemm_add_ln(b.o, out_l.W, out_l.b, x, b.s, n1[0], n1[1], b.y1, mean, rstd, D)
            K.gemm(b.y1, q_l.W, b.q, bias=q_l.b)
            attn_cross(b.q, b.o)
            K.gemm_add_ln(b.o, out2_l.W, out2_l.b, b.y1, b.s, n2[0], n2[1], b.y2, mean, rstd, D)
            K.gemm(b.y2, f1.W, b.h1, bias=f1.b, epi=1)
            K.gemm_add_ln(b.h1, f2.W, f2.b, b.y2, b.s, n3[0], n3[1], b.y3, mean, rstd, D)
            return b.y3
        attn_self(row, b.o)
        K.decode_chain(b.o, x, out_l.W, out_l.b, n1[0], n1[1], b.y1, D, proj=(q_l.W, q_l.b, b.q))
        attn_cross(b.q, b.o)
        nxt = None
        if next_row is not None:
            nq = self.lin[f"transformer_decoder.layers.{l + 1}.self_attn.qkv"]
            nxt = (nq.W, nq.b, next_row)
        K.decode_chain(b.o, b.y1, out2_l.W, out2_l.b, n2[0], n2[1], b.y3, D, ffn=(f1.W, f1.b, f2.W, f2.b, n3[0], n3[1]), proj=nxt)
        return b.y3

    def _decode_fused(self) -> bool:
        # OFF by default: measured 30.2 ms against 26.7 ms per 625-image greedy decode on a B200 (profiles/README.md) - the 16-row
        # mma.sync chains are latency-bound, the tcgen05 GEMMs with programmatic dependent launch are faster.  "1": bf16 decode loops
        # use it; "force": also for fp32 (host-simulated tests of the orchestration).
        mode = os.environ.get("ICK_DECODE_CHAIN", "0")
        return mode == "force" or (self.dtype == torch.bfloat16 and mode == "1")

    # ---- greedy decode (predict) -----------------------------------------------------------------------------------------------------
    def greedy_decode(self, inp, Tmax: int, return_margins: bool = False):
        """
        DecoderTransformer.predict (G/models.py:363-443, K/models.py:516-609) for a BATCH of images, device-resident:
        KV-cached incremental decoding (the reference re-decodes all positions every step), and the argmax / top-2 /
        <end> / repetition clean-up / next-token bookkeeping runs in one kernel per step, so there is no host
        round-trip inside the loop.  inp: encoder_out (B,D,P) f32, entities (B,E,C) f32, facts (B,F,3) i64 | None.
        Returns output (B, Tmax) int64 (pad-filled after <end>).
        """
        K, D, DP, L, H, dh, V = self.K, self.D, self.DP, self.L, self.H, self.dh, self.V
        B = inp.encoder_out.shape[0]
        E = inp.entities.shape[1]
        F = inp.facts.shape[1] if self.has_facts else 0
        P = inp.encoder_out.shape[2]
        M = P + E + F
        W = V + E + F
        dev = self.device
        ctx = NS(inp=inp)
        ctx.ent_enc, ctx.fact_enc = self._encode_context(inp, B, E, F)
        mem, _ = self._build_memory(inp, ctx.ent_enc, ctx.fact_enc, B, E, F, P, M, 0.0, None)
        kvs = self._memory_kv(mem, B * M)
        captions = torch.full((B, Tmax), self.start, dtype=torch.int64, device=dev)
        masks = torch.zeros((B, Tmax), dtype=torch.int64, device=dev)
        output = torch.full((B, Tmax), self.pad, dtype=torch.int64, device=dev)
        second = torch.zeros((B, Tmax), dtype=torch.int32, device=dev)
        done = torch.zeros(B, dtype=torch.int32, device=dev)
        margins = torch.zeros((B, Tmax), dtype=torch.float32, device=dev) if return_margins else None
        cache = [self._new(B * Tmax, 3 * DP) for _ in range(L)]
        scores = torch.empty(B, W, dtype=torch.float32, device=dev)
        mean, rstd = self._newf(B), self._newf(B)
        x0 = self._new(B, DP)
        bufs = [NS(o=self._new(B, DP), s=self._new(B, DP), y1=self._new(B, DP), q=self._new(B, DP), y2=self._new(B, DP),
                   h1=self._new(B, self.lin[f"transformer_decoder.layers.{l}.ffn1"].lin.Np), y3=self._new(B, DP)) for l in range(L)]
        fused = self._decode_fused()
        for i in range(Tmax):
            K.caption_embed_fwd(captions, masks, self.wemb, ctx.ent_enc, ctx.fact_enc, self.pe, x0, B, Tmax, i, 1, V, E, F, D, self.pad,
                                math.sqrt(D))
            x = x0
            rows_i = [cache[l].view(B, Tmax, 3 * DP)[:, i, :] for l in range(L)]  # this step's q|k|v rows inside the caches
            if fused:
                qkv0 = self.lin["transformer_decoder.layers.0.self_attn.qkv"]
                K.gemm(x, qkv0.W, rows_i[0], bias=qkv0.b)
            for l in range(L):
                Kl, Vl, kvw = kvs[l]
                x = self._dec_step_layer(
                    l, x, rows_i[l], rows_i[l + 1] if l + 1 < L else None, bufs[l],
                    lambda row, out, l=l: K.mha_decode(row[:, :DP], cache[l][:, DP : 2 * DP], cache[l][:, 2 * DP :], out, B, H, dh,
                                                       Tmax * 3 * DP, Tmax * 3 * DP, i + 1),
                    lambda q, out, Kl=Kl, Vl=Vl, kvw=kvw: K.mha_decode(q, Kl, Vl, out, B, H, dh, M * kvw, M * kvw, M),
                    B, mean, rstd, fused)
            self._heads_fwd(ctx, captions, x, scores, B, 1, i, Tmax, E, F, lag=1)
            K.greedy_select(scores, W, output, second, captions, masks, done, margins, B, i, Tmax, V, E, self.has_facts, self.end)
        return (output, margins) if return_margins else output

    # ---- beam-search decode (extension: the reference only has the greedy predict) ---------------------------------------------------
    def beam_decode(self, inp, Tmax: int, beam: int = 5):
        """
        Beam search over the same KV-cached single-position decoder pass as greedy_decode, for a BATCH of images, device-resident.
        EXTENSION: /root/reference has no beam search (BASELINE.json asks for beam-5).  The algorithm is the
        Show-Attend-Tell tutorial's (oracle/decoder_oracle.py:beam_search restates it): summed log-probabilities, k shrinks as
        captions complete, best completed caption wins, no length normalisation, no repetition clean-up.

        Rows are (image, beam slot).  The beams of an image share its entity / fact encodings and its memory K/V (read once per
        image per step by the cross-attention); the self-attention cache is position-major and never re-ordered - each beam
        carries a (Tmax,) table of the slots its ancestors occupied, and `beam_select` (top-k over the live rows, <end>
        handling, history / ancestor-table hand-over, best-caption bookkeeping) is one kernel per step, so the loop has no
        host round trip.  Returns (result (B, Tmax) int64 - no <start>, <end> included, <pad>-filled -, score (B,) fp32).
        """
        K, D, DP, L, H, dh, V = self.K, self.D, self.DP, self.L, self.H, self.dh, self.V
        NI = inp.encoder_out.shape[0]
        G = int(beam)
        assert 1 <= G <= 8, "beam width must be in [1, 8]"
        R = NI * G
        E = inp.entities.shape[1]
        F = inp.facts.shape[1] if self.has_facts else 0
        P = inp.encoder_out.shape[2]
        M = P + E + F
        W = V + E + F
        dev = self.device
        ctx = NS(inp=inp)
        ctx.ent_enc, ctx.fact_enc = self._encode_context(inp, NI, E, F)
        mem, _ = self._build_memory(inp, ctx.ent_enc, ctx.fact_enc, NI, E, F, P, M, 0.0, None)
        kvs = self._memory_kv(mem, NI * M)
        tok = [torch.full((R, Tmax), self.start, dtype=torch.int64, device=dev) for _ in range(2)]
        msk = [torch.zeros((R, Tmax), dtype=torch.int64, device=dev) for _ in range(2)]
        own = (torch.arange(R, dtype=torch.int32, device=dev) % G).unsqueeze(1).expand(R, Tmax)
        # anc[r, j]: slot of row r's ancestor at position j (own slot by default); two distinct buffers (clone: .contiguous() of
        # an expanded (R, 1) view is the view itself)
        anc = [own.clone(memory_format=torch.contiguous_format) for _ in range(2)]
        cum = torch.zeros(R, dtype=torch.float32, device=dev)
        ksel = torch.full((NI,), G, dtype=torch.int32, device=dev)
        best = torch.full((NI,), float("-inf"), dtype=torch.float32, device=dev)
        result = torch.full((NI, Tmax), self.pad, dtype=torch.int64, device=dev)
        cache = [self._new(Tmax * R, 3 * DP) for _ in range(L)]  # row j*R + r: q|k|v of row r at position j
        pos = R * 3 * DP
        scores = torch.empty(R, W, dtype=torch.float32, device=dev)
        mean, rstd = self._newf(R), self._newf(R)
        x0 = self._new(R, DP)
        bufs = [NS(o=self._new(R, DP), s=self._new(R, DP), y1=self._new(R, DP), q=self._new(R, DP), y2=self._new(R, DP),
                   h1=self._new(R, self.lin[f"transformer_decoder.layers.{l}.ffn1"].lin.Np), y3=self._new(R, DP)) for l in range(L)]
        # cross-attention of the G beams of an image: "tma" (default) = ick_mha_decode_beam, which runs shared contiguous K|V rows on the
        # TMA-streamed tensor-core kernel (csrc/attention_decode_tma.cu); "flash" = the G beams as G query positions of one
        # flash-attention item per (image, head) (the round-1 path: per-head TMA boxes, 4 TB/s)
        xattn_flash = os.environ.get("ICK_BEAM_XATTN", "tma") == "flash" and self.dtype != torch.float32
        lse = self._newf(NI * H * G)
        cand = self._newf(R * G * 2)  # per-row candidate lists of beam_select
        fused = self._decode_fused()
        for i in range(Tmax):
            cur, nxt = i & 1, (i + 1) & 1
            K.caption_embed_fwd(tok[cur], msk[cur], self.wemb, ctx.ent_enc, ctx.fact_enc, self.pe, x0, R, Tmax, i, 1, V, E, F, D,
                                self.pad, math.sqrt(D), group=G)
            x = x0
            rows_i = [cache[l][i * R : (i + 1) * R] for l in range(L)]
            if fused:
                qkv0 = self.lin["transformer_decoder.layers.0.self_attn.qkv"]
                K.gemm(x, qkv0.W, rows_i[0], bias=qkv0.b)

            def cross(q, out, l):
                Kl, Vl, kvw = kvs[l]
                if xattn_flash:  # the G beams of an image are G query positions of one flash-attention item over its memory
                    K.mha_fwd(q, Kl, Vl, out, lse, NI, H, G, M, dh)
                else:
                    K.mha_decode_beam(q, Kl, Vl, out, R, G, H, dh, M, kimg_stride=M * kvw, vimg_stride=M * kvw)

            for l in range(L):
                x = self._dec_step_layer(
                    l, x, rows_i[l], rows_i[l + 1] if l + 1 < L else None, bufs[l],
                    lambda row, out, l=l: K.mha_decode_beam(row[:, :DP], cache[l][:, DP : 2 * DP], cache[l][:, 2 * DP :], out, R, G, H, dh,
                                                            i + 1, anc=anc[cur], kpos_stride=pos, vpos_stride=pos),
                    lambda q, out, l=l: cross(q, out, l), R, mean, rstd, fused)
            self._heads_fwd(ctx, tok[cur], x, scores, R, 1, i, Tmax, E, F, lag=1, group=G)
            K.beam_select(scores, W, cum, ksel, tok[cur], msk[cur], tok[nxt], msk[nxt], anc[cur], anc[nxt], best, result, NI, G, i, Tmax,
                          V, E, self.has_facts, self.end, self.pad, workspace=cand)
        return result, best
