"""
HOST SIMULATION OF THE KERNEL SET — TEST INFRASTRUCTURE ONLY (never imported by the product).

``HostKernels`` implements the method-for-method contract of ``ickb200.kernels.CudaKernels`` (i.e. of
include/ickb200.h) with plain torch CPU ops, so that the host-side orchestration — engine.forward / the hand-written
engine.backward chain, the packing plan and its gradient index maps, the module's sort / autograd / flat-parameter
plumbing, the trainer — can be checked against the oracle on a box without a GPU (`pytest -m "not gpu"`).
It is deliberately slow and simple; the GPU tests check the real kernels against the same oracle.
"""
import math

import numpy as np
import torch

from dropout_ref import attn_drop_mul, drop_mul

FIRST_NONE = 1 << 29
HD = 32


_SEED_SRC = [None]


def _mul(drop, rows, cols, attn=False):
    if drop is None:
        return None
    p, seed, site = drop
    if p <= 0:
        return None
    if _SEED_SRC[0] is not None:
        seed = (seed + int(_SEED_SRC[0][0])) & 0xFFFFFFFF
    return (attn_drop_mul if attn else drop_mul)(p, seed, site, rows, cols)


def _rows(rowmap, R):
    s_in, s_out, off = rowmap
    r = torch.arange(R)
    if s_in == 0:
        return r
    return (r // s_in) * s_out + off + r % s_in


class HostKernels:
    name = "hostsim"

    def __init__(self):
        self.calls = 0

    # ---- dense ---------------------------------------------------------------------------------------------------
    def gemm(self, A, W, C, bias=None, aux=None, epi=0, accumulate=False, drop=None, force_simt=False):
        self.calls += 1
        M, K = A.shape
        N = W.shape[0]
        acc = A.float() @ W.float().t()
        if bias is not None:
            acc = acc + bias.float()[:N]
        if accumulate:
            acc = acc + C.float()
        if epi == 1:
            acc = torch.relu(acc)
            m = _mul(drop, M, N)
            if m is not None:
                acc = acc * m.view(M, N)
        elif epi == 2:
            inv = 1.0 / (1.0 - drop[0]) if drop is not None and drop[0] > 0 else 1.0
            acc = torch.where(aux.float() != 0, acc * inv, torch.zeros_like(acc))
        C.copy_(acc.to(C.dtype))

    def wgrad(self, dY, X, gflat, rowoff, colmap=None, biasoff=None, force_simt=False):
        self.calls += 1
        M, N = dY.shape
        K = X.shape[1]
        G = dY.float().t() @ X.float()  # (N, K)
        ro = rowoff[:N].long()
        cm = colmap[:K].long() if colmap is not None else torch.arange(K)
        nv = (ro >= 0).nonzero().flatten()
        kv = (cm >= 0).nonzero().flatten()
        idx = (ro[nv][:, None] + cm[kv][None, :]).reshape(-1)
        gflat.index_add_(0, idx, G[nv][:, kv].reshape(-1))
        if biasoff is not None:
            bo = biasoff[:N].long()
            bv = (bo >= 0).nonzero().flatten()
            gflat.index_add_(0, bo[bv], dY.float().sum(0)[bv])

    def gemm_dual(self, A, W0, W1, C, m_split, rows0, bias0=None, bias1=None, aux=None, epi=0, accumulate=False, drop0=None, drop1=None):
        sl = lambda t, a, b: None if t is None else t[a:b]  # noqa: E731
        M = A.shape[0]
        self.gemm(A[:rows0], W0, C[:rows0], bias0, sl(aux, 0, rows0), epi, accumulate, drop0)
        self.gemm(A[m_split:], W1, C[m_split:], bias1, sl(aux, m_split, M), epi, accumulate, drop1)

    def wgrad_group(self, probs, gflat):
        for dY, X, rowoff, colmap, biasoff in probs:
            self.wgrad(dY, X, gflat, rowoff, colmap, biasoff)

    def gemm_rowdot(self, A, W, C, O, dsum, S, H, W1=None, m_split=0, rows0=None, dsum1=None, S1=0) -> bool:
        """ick_gemm_tn_tc_rowdot: C = A W^T and dsum[(b*H + h)*S + i] = sum_d C[b*S+i, 32h+d] * O[b*S+i, 32h+d] (C as stored)."""
        M = A.shape[0]
        if W1 is None:
            self.gemm(A, W, C)
            groups = ((0, M, S, dsum),)
        else:
            self.gemm_dual(A, W, W1, C, m_split, rows0)
            groups = ((0, rows0, S, dsum), (m_split, M - m_split, S1, dsum1))
        for r0, n, Sg, out in groups:
            prod = (C[r0:r0 + n, :H * HD].float() * O[r0:r0 + n, :H * HD].float()).view(n // Sg, Sg, H, HD).sum(-1)  # (b, i, h)
            out.copy_(prod.permute(0, 2, 1).reshape(-1))
        return True

    def gemm_add_ln(self, A, W, bias, x, s, gamma, beta, y, mean, rstd, d, eps=1e-5, drop=None):
        """ick_gemm_add_ln_tc as its two-kernel equivalent"""
        self.gemm(A, W, s, bias=bias)
        self.add_ln_fwd(x, s, gamma, beta, y, mean, rstd, d, eps, drop=drop)

    def gemm_add_ln_dual(self, A, W0, W1, bias0, bias1, x, s, y, mean, rstd, d, m_split, rows0, gammas, betas, drops=(None, None), eps=1e-5):
        self.gemm_dual(A, W0, W1, s, m_split, rows0, bias0, bias1)
        self.add_ln_fwd_dual(x, s, y, mean, rstd, d, rows0, A.shape[0] - m_split, m_split, gammas, betas, drops=drops, eps=eps)

    # ---- attention -------------------------------------------------------------------------------------------------
    @staticmethod
    def _heads(X, B, S, H):
        return X.float().reshape(B, S, H, HD).permute(0, 2, 1, 3)  # (B,H,S,32)

    def _probs(self, Q, K, B, H, Sq, Sk, dh, causal):
        q = self._heads(Q, B, Sq, H)[..., :dh]
        k = self._heads(K, B, Sk, H)[..., :dh]
        s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
        if causal:
            s = s + torch.full((Sq, Sk), float("-inf")).triu(1)
        return q, k, s

    def mha_fwd(self, Q, K, V, O, lse, B, H, Sq, Sk, dh, causal=False, drop=None):
        self.calls += 1
        q, k, s = self._probs(Q, K, B, H, Sq, Sk, dh, causal)
        v = self._heads(V, B, Sk, H)[..., :dh]
        p = torch.softmax(s, dim=-1)
        lse.copy_((torch.logsumexp(s, dim=-1) * 1.4426950408889634).reshape(-1))
        m = _mul(drop, p.numel() // p.shape[-1], p.shape[-1], attn=True)
        pd = p * m.view_as(p) if m is not None else p
        o = pd @ v  # (B,H,Sq,dh)
        out = torch.zeros(B, H, Sq, HD)
        out[..., :dh] = o
        O.copy_(out.permute(0, 2, 1, 3).reshape(B * Sq, H * HD).to(O.dtype))

    def mha_bwd(self, Q, K, V, O, dO, lse, dsum, dQ, dK, dV, B, H, Sq, Sk, dh, causal=False, drop=None, dsum_ready=False):
        self.calls += 2
        given = dsum.clone() if dsum_ready else None
        q, k, s = self._probs(Q, K, B, H, Sq, Sk, dh, causal)
        v = self._heads(V, B, Sk, H)[..., :dh]
        go = self._heads(dO, B, Sq, H)[..., :dh]
        p = torch.softmax(s, dim=-1)
        m = _mul(drop, p.numel() // p.shape[-1], p.shape[-1], attn=True)
        mm = m.view_as(p) if m is not None else torch.ones_like(p)
        dv = (p * mm).transpose(-1, -2) @ go
        dp = (go @ v.transpose(-1, -2)) * mm
        D = (dp * p).sum(-1, keepdim=True)
        if given is not None:  # the row term handed over by gemm_rowdot must be the one this backward would compute itself
            tol = 1e-4 if O.dtype == torch.float32 else 3e-2
            assert float((given - D.reshape(-1)).abs().max()) <= tol * max(1.0, float(D.abs().max())), "dsum_ready: stale / wrong row term"
        dsum.copy_(D.reshape(-1))
        ds = p * (dp - D)
        dq = ds @ k / math.sqrt(dh)
        dk = ds.transpose(-1, -2) @ q / math.sqrt(dh)

        def put(dst, x, S):
            out = torch.zeros(B, H, S, HD)
            out[..., :dh] = x
            dst.copy_(out.permute(0, 2, 1, 3).reshape(B * S, H * HD).to(dst.dtype))

        put(dQ, dq, Sq)
        put(dK, dk, Sk)
        put(dV, dv, Sk)

    def mha_decode(self, Q, K, V, O, B, H, dh, kbatch_stride, vbatch_stride, klen):
        self.calls += 1
        ldk, ldv = K.stride(0), V.stride(0)
        rk, rv = kbatch_stride // ldk, vbatch_stride // ldv
        for b in range(B):
            q = Q[b].float().view(H, HD)[:, :dh]
            k = K[b * rk : b * rk + klen].float().reshape(klen, H, HD)[..., :dh]
            v = V[b * rv : b * rv + klen].float().reshape(klen, H, HD)[..., :dh]
            s = torch.einsum("hd,khd->hk", q, k) / math.sqrt(dh)
            p = torch.softmax(s, dim=-1)
            o = torch.zeros(H, HD)
            o[:, :dh] = torch.einsum("hk,khd->hd", p, v)
            O[b] = o.reshape(-1).to(O.dtype)

    def mha_decode_beam(self, Q, K, V, O, rows, group, H, dh, klen, kimg_stride=0, vimg_stride=0, anc=None, kpos_stride=0, vpos_stride=0):
        self.calls += 1
        ldk, ldv = K.stride(0), V.stride(0)
        for r in range(rows):
            img = r // group
            if anc is None:
                kr = [img * (kimg_stride // ldk) + j for j in range(klen)]
                vr = [img * (vimg_stride // ldv) + j for j in range(klen)]
            else:
                kr = [j * (kpos_stride // ldk) + img * group + int(anc[r, j]) for j in range(klen)]
                vr = [j * (vpos_stride // ldv) + img * group + int(anc[r, j]) for j in range(klen)]
            q = Q[r].float().view(H, HD)[:, :dh]
            k = K[kr].float().reshape(klen, H, HD)[..., :dh]
            v = V[vr].float().reshape(klen, H, HD)[..., :dh]
            p = torch.softmax(torch.einsum("hd,khd->hk", q, k) / math.sqrt(dh), dim=-1)
            o = torch.zeros(H, HD)
            o[:, :dh] = torch.einsum("hk,khd->hd", p, v)
            O[r] = o.reshape(-1).to(O.dtype)

    # ---- residual + dropout + layer norm ---------------------------------------------------------------------------
    def add_ln_fwd(self, x, sub, gamma, beta, y, mean, rstd, d, eps=1e-5, rowmap=(0, 0, 0), drop=None):
        self.calls += 1
        R = sub.shape[0]
        s = sub[:, :d].float()
        m = _mul(drop, R, d)
        if m is not None:
            s = s * m.view(R, d)
        if x is not None:
            s = s + x[:, :d].float()
        sub[:, :d] = s.to(sub.dtype)
        mu = s.mean(-1, keepdim=True)
        var = ((s - mu) ** 2).mean(-1, keepdim=True)
        rs = torch.rsqrt(var + eps)
        out = (s - mu) * rs * gamma.float() + beta.float()
        rows = _rows(rowmap, R)
        y[rows] = torch.cat([out, torch.zeros(R, y.shape[1] - d)], dim=1).to(y.dtype)
        mean.copy_(mu.flatten())
        rstd.copy_(rs.flatten())

    def add_ln_fwd_dual(self, x, sub, y, mean, rstd, d, rows0, rows1, row1, gammas, betas, rowmaps=((0, 0, 0), (0, 0, 0)), drops=(None, None),
                        eps=1e-5):
        for i, (r0, n) in enumerate(((0, rows0), (row1, rows1))):
            mapped = rowmaps[i][0] != 0
            self.add_ln_fwd(x[r0:r0 + n], sub[r0:r0 + n], gammas[i], betas[i], y if mapped else y[r0:r0 + n], mean[r0:r0 + n], rstd[r0:r0 + n],
                            d, eps, rowmaps[i], drops[i])

    def add_ln_bwd_dual(self, dy, s, mean, rstd, dres, dsub, d, rows0, rows1, row1, gammas, dgammas, dbetas, rowmaps=((0, 0, 0), (0, 0, 0)),
                        drops=(None, None), acc_res=False):
        for i, (r0, n) in enumerate(((0, rows0), (row1, rows1))):
            mapped = rowmaps[i][0] != 0
            self.add_ln_bwd(dy if mapped else dy[r0:r0 + n], s[r0:r0 + n], mean[r0:r0 + n], rstd[r0:r0 + n], gammas[i], dres[r0:r0 + n],
                            dsub[r0:r0 + n], dgammas[i], dbetas[i], d, rowmaps[i], acc_res, drops[i])

    def add_ln_bwd(self, dy, s, mean, rstd, gamma, dres, dsub, dgamma, dbeta, d, rowmap=(0, 0, 0), acc_res=False, drop=None):
        self.calls += 1
        R = s.shape[0]
        rows = _rows(rowmap, R)
        g_in = dy[rows][:, :d].float()
        xh = (s[:, :d].float() - mean.view(R, 1)) * rstd.view(R, 1)
        g = g_in * gamma.float()
        mg = g.mean(-1, keepdim=True)
        mgx = (g * xh).mean(-1, keepdim=True)
        dx = rstd.view(R, 1) * (g - mg - xh * mgx)
        if dgamma is not None:
            dgamma.add_((g_in * xh).sum(0).view_as(dgamma))
        if dbeta is not None:
            dbeta.add_(g_in.sum(0).view_as(dbeta))
        if dsub is not None:
            m = _mul(drop, R, d)
            v = dx * m.view(R, d) if m is not None else dx
            dsub.zero_()
            dsub[:, :d] = v.to(dsub.dtype)
        if dres is not None:
            if acc_res:
                dres[:, :d] = (dres[:, :d].float() + dx).to(dres.dtype)
            else:
                dres.zero_()
                dres[:, :d] = dx.to(dres.dtype)

    # ---- context preparation ------------------------------------------------------------------------------------------
    @staticmethod
    def _counts(facts, B, E):
        subj = facts[:, :, 1]
        cnt = torch.zeros(B, E)
        for b in range(B):
            for sidx in subj[b].tolist():
                if 0 <= sidx < E - 1:
                    cnt[b, sidx] += 1
        return cnt

    def _ent_base(self, entities, facts, type_emb, variant, B, E, D, ntypes):
        ent = entities.float().view(B, E, -1)
        ty = ent[:, :, 4].long().clamp(0, ntypes - 1)
        az = ent[:, :, 2]
        east = torch.where(az >= -90.0, (90.0 - az).abs(), 90.0 + (az + 180.0).abs()) / 180.0
        if variant == 0:
            feats = [ent[:, :, 1], az.abs() / 180.0, east, ent[:, :, 3]]
        elif variant == 1:
            cnt = self._counts(facts, B, E)
            feats = [ent[:, :, 1], az.abs() / 180.0, east, ent[:, :, 3], cnt, (cnt > 0).float()]
        else:
            cnt = self._counts(facts, B, E)
            feats = [ent[:, :, 1], ent[:, :, 2], ent[:, :, 3], cnt, (cnt > 0).float()]
        return torch.cat([torch.stack(feats, dim=2), type_emb.float()[ty]], dim=2), ty, len(feats)

    def entity_encode_fwd(self, entities, facts, type_emb, word_emb, out, variant, B, E, F, D, ntypes, V):
        self.calls += 1
        base, _, _ = self._ent_base(entities, facts, type_emb, variant, B, E, D, ntypes)
        if variant == 2:
            nm = entities.float().view(B, E, -1)[:, :, 5:10].long().clamp(0, V - 1)
            base = base * word_emb.float()[nm][..., :D].mean(dim=-2)
        out.zero_()
        out[:, :D] = base.view(B * E, D).to(out.dtype)

    def entity_encode_bwd(self, dEnt, entities, facts, type_emb, word_emb, gflat, type_off, word_off, dt, variant, B, E, F, D, ntypes, V):
        self.calls += 1
        base, ty, nf = self._ent_base(entities, facts, type_emb, variant, B, E, D, ntypes)
        g = dEnt[:, :D].float().view(B, E, D)
        w = D - nf
        if variant == 2:
            nm = entities.float().view(B, E, -1)[:, :, 5:10].long().clamp(0, V - 1)
            avg = word_emb.float()[nm][..., :D].mean(dim=-2)
            gt = (g * avg)[:, :, nf:]
            dn = (g * base / 5.0).reshape(B * E, D)
            for k in range(5):
                idx = (word_off + nm[:, :, k].reshape(-1, 1) * D + torch.arange(D)[None, :]).reshape(-1)
                gflat.index_add_(0, idx, dn.reshape(-1))
        else:
            gt = g[:, :, nf:]
        idx = (type_off + ty.reshape(-1, 1) * w + torch.arange(w)[None, :]).reshape(-1)
        gflat.index_add_(0, idx, gt.reshape(-1))

    def fact_encode_fwd(self, facts, ent_enc, pred_emb, out, B, E, F, D, NP):
        self.calls += 1
        subj = facts[:, :, 1].long().clamp(0, E - 1)
        pred = facts[:, :, 2].long().clamp(0, NP - 1)
        e = ent_enc.float().view(B, E, -1)[:, :, :D]
        g = torch.gather(e, 1, subj.unsqueeze(-1).expand(-1, -1, D)) + pred_emb.float()[pred]
        out.zero_()
        out[:, :D] = g.view(B * F, D).to(out.dtype)

    def fact_encode_bwd(self, dFact, facts, dEnt, gflat, pred_off, B, E, F, D, NP):
        self.calls += 1
        subj = facts[:, :, 1].long().clamp(0, E - 1)
        pred = facts[:, :, 2].long().clamp(0, NP - 1)
        g = dFact[:, :D].float()
        rows = (torch.arange(B)[:, None] * E + subj).reshape(-1)
        dEnt[:, :D].index_add_(0, rows, g)
        idx = (pred_off + pred.reshape(-1, 1) * D + torch.arange(D)[None, :]).reshape(-1)
        gflat.index_add_(0, idx, g.reshape(-1))

    @staticmethod
    def _select(tok, mask, V, E, F, pad):
        kind = torch.zeros_like(tok)
        idx = torch.where((tok >= V) | (tok < 0), torch.full_like(tok, pad), tok)
        e = tok - V
        e = torch.where((e < 0) | (e >= E), torch.full_like(e, E - 1), e)
        kind = torch.where(mask == 1, torch.ones_like(kind), kind)
        idx = torch.where(mask == 1, e, idx)
        if F > 0:
            f = tok - V - E
            f = torch.where((f < 0) | (f >= F), torch.full_like(f, F - 1), f)
            kind = torch.where(mask == 2, torch.full_like(kind, 2), kind)
            idx = torch.where(mask == 2, f, idx)
        return kind, idx

    def caption_embed_fwd(self, captions, masks, word_emb, ent_enc, fact_enc, pe, out, B, Tstride, t0, Tn, V, E, F, D, pad, scale, drop=None,
                          group=1):
        self.calls += 1
        if group > 1:  # beams of an image share its context
            ent_enc = ent_enc.view(B // group, E, -1).repeat_interleave(group, 0).reshape(B * E, -1)
            if F > 0:
                fact_enc = fact_enc.view(B // group, F, -1).repeat_interleave(group, 0).reshape(B * F, -1)
        tok = captions.view(B, Tstride)[:, t0 : t0 + Tn]
        mk = masks.view(B, Tstride)[:, t0 : t0 + Tn]
        kind, idx = self._select(tok, mk, V, E, F, pad)
        emb = word_emb.float()[torch.where(kind == 0, idx, torch.zeros_like(idx))][..., :D]
        e = ent_enc.float().view(B, E, -1)[..., :D]
        ge = torch.gather(e, 1, torch.where(kind == 1, idx, torch.zeros_like(idx)).unsqueeze(-1).expand(-1, -1, D))
        emb = torch.where((kind == 1).unsqueeze(-1), ge, emb)
        if F > 0:
            f = fact_enc.float().view(B, F, -1)[..., :D]
            gf = torch.gather(f, 1, torch.where(kind == 2, idx, torch.zeros_like(idx)).unsqueeze(-1).expand(-1, -1, D))
            emb = torch.where((kind == 2).unsqueeze(-1), gf, emb)
        x = emb * scale + pe[t0 : t0 + Tn].float().unsqueeze(0)
        m = _mul(drop, B * Tn, D)
        if m is not None:
            x = x * m.view(B, Tn, D)
        out.zero_()
        out[:, :D] = x.reshape(B * Tn, D).to(out.dtype)

    def caption_embed_bwd(self, dX, captions, masks, dEnt, dFact, gflat, word_off, B, T, V, E, F, D, pad, scale, drop=None):
        self.calls += 1
        kind, idx = self._select(captions.view(B, T), masks.view(B, T), V, E, F, pad)
        g = dX[:, :D].float() * scale
        m = _mul(drop, B * T, D)
        if m is not None:
            g = g * m.view(B * T, D)
        kind, idx = kind.reshape(-1), idx.reshape(-1)
        brow = torch.arange(B).repeat_interleave(T)
        w = kind == 0
        gi = (word_off + idx[w].reshape(-1, 1) * D + torch.arange(D)[None, :]).reshape(-1)
        gflat.index_add_(0, gi, g[w].reshape(-1))
        e = kind == 1
        dEnt[:, :D].index_add_(0, brow[e] * E + idx[e], g[e])
        if F > 0:
            f = kind == 2
            dFact[:, :D].index_add_(0, brow[f] * F + idx[f], g[f])

    def pixels_fwd(self, encoder_out, memory, B, D, P, M):
        self.calls += 1
        mem = memory.view(B, M, -1)
        mem[:, :P, :] = 0
        mem[:, :P, :D] = encoder_out.float().permute(0, 2, 1).to(memory.dtype)

    def pixels_bwd(self, dmemory, d_encoder_out, B, D, P, M):
        self.calls += 1
        d_encoder_out.copy_(dmemory.view(B, M, -1)[:, :P, :D].float().permute(0, 2, 1))

    def pool_rows_fwd(self, x, rows, B, C, Hin, Win, Hout, Wout):
        self.calls += 1
        pooled = torch.nn.functional.adaptive_avg_pool2d(x.float().view(B, C, Hin, Win), (Hout, Wout))
        rows[:, :C] = pooled.permute(0, 2, 3, 1).reshape(B * Hout * Wout, C).to(rows.dtype)

    # ---- indicators / gate ------------------------------------------------------------------------------------------------
    def fact_first_mention(self, captions, facts, first_t, tmin, B, T, F, V, E, group=1, NP=0):
        self.calls += 1
        if group > 1:
            facts = facts.repeat_interleave(group, 0)
        caps = captions.view(B, T)
        ft = first_t.view(B, F)
        tm = tmin.view(B, F)
        for b in range(B):
            toks = caps[b].tolist()
            subj = facts[b, :, 1].tolist()
            pred = facts[b, :, 2].tolist()
            firsts = []
            for f in range(F):
                first = FIRST_NONE
                for t, tok in enumerate(toks):
                    if V <= tok < V + E and tok - V == subj[f]:
                        first = t
                        break
                firsts.append(first)
                ft[b, f] = first
            for f in range(F):
                same = [g for g in range(F) if pred[g] == pred[f]]
                tm[b, f] = min(firsts[g] for g in same) if same[0] == f else FIRST_NONE

    def pred_gate_fwd(self, tmin, facts, WpT, bias, h, gate, hg, B, Tn, t0, F, D, NP, lag, group=1):
        self.calls += 1
        if group > 1:
            facts = facts.repeat_interleave(group, 0)
        tm = tmin.view(B, F)
        gate.zero_()
        for b in range(B):
            pred = facts[b, :, 2].long().clamp(0, NP - 1)
            for tt in range(Tn):
                act = tm[b] < (t0 + tt + lag)
                g = bias.float() + WpT.float()[pred[act]][:, :D].sum(0)
                gate[b * Tn + tt, :D] = g.to(gate.dtype)
        if hg is not None:
            hg.copy_((h.float() * gate.float()).to(hg.dtype))

    def gate_mul_bwd(self, dHG, h, gate, dG, dH):
        self.calls += 1
        dG.copy_((dHG.float() * h.float()).to(dG.dtype))
        dH.copy_((dHG.float() * gate.float()).to(dH.dtype))

    def pred_gate_bwd(self, dG, tmin, facts, gflat, wp_off, B, T, F, D, NP, lag):
        self.calls += 1
        tm = tmin.view(B, F)
        g = dG[:, :D].float().view(B, T, D)
        for b in range(B):
            for f in range(F):
                t_m = int(tm[b, f])
                if t_m >= FIRST_NONE:
                    continue
                p = int(facts[b, f, 2].clamp(0, NP - 1))
                s = g[b, max(0, t_m + 1 - lag) :].sum(0)
                gflat.index_add_(0, wp_off + torch.arange(D) * NP + p, s)

    # ---- pointer heads -----------------------------------------------------------------------------------------------------
    def _mask(self, first_t, B, Tn, t0, S, lag):
        if first_t is None:
            return torch.ones(B, Tn, S)
        t = (t0 + torch.arange(Tn) + lag).view(1, Tn, 1)
        return (first_t.view(B, 1, S) < t).float()

    def pointer_fwd(self, h, ctx, w, bias, first_t, scores, B, Tn, t0, S, D, col0, lag, group=1):
        self.calls += 1
        if group > 1:
            assert Tn == 1
            ctx = ctx.view(B // group, S, -1).repeat_interleave(group, 0).reshape(B * S, -1)
        hh = h[:, :D].float().view(B, Tn, D) * w.float().view(1, 1, D)
        c = ctx[:, :D].float().view(B, S, D)
        s = torch.einsum("btd,bsd->bts", hh, c) * self._mask(first_t, B, Tn, t0, S, lag) + bias.float().view(1, 1, 1)
        scores.view(B, Tn, -1)[:, :, col0 : col0 + S] = s

    def pointer_bwd(self, dS, h, ctx, w, first_t, dCtx, dH, gflat, w_off, bias_off, B, T, S, D, col0, lag):
        self.calls += 2
        g = dS.float().view(B, T, -1)[:, :, col0 : col0 + S]
        gflat[bias_off] += g.sum()
        gm = g * self._mask(first_t, B, T, 0, S, lag)
        hh = h[:, :D].float().view(B, T, D)
        c = ctx[:, :D].float().view(B, S, D)
        G = torch.einsum("bts,bsd->btd", gm, c)
        dH[:, :D] = (dH[:, :D].float() + (G * w.float().view(1, 1, D)).reshape(B * T, D)).to(dH.dtype)
        gflat[w_off : w_off + D] += (hh * G).sum((0, 1))
        dCtx[:, :D] += (torch.einsum("bts,btd->bsd", gm, hh) * w.float().view(1, 1, D)).reshape(B * S, D)

    # ---- loss / optimizer / misc ---------------------------------------------------------------------------------------------
    def ce(self, scores, captions_sorted, decode_len, loss_acc, dscores, B, T, W, pad):
        self.calls += 1
        s = scores.view(B, T, -1)[:, :, :W].float()
        caps = captions_sorted.view(B, T)
        tgt = torch.cat([caps[:, 1:], torch.full((B, 1), pad, dtype=caps.dtype)], dim=1)
        t = torch.arange(T).view(1, T)
        valid = (t < decode_len.view(B, 1)) & (t + 1 < T) & (tgt != pad)
        lse = torch.logsumexp(s, dim=-1)
        picked = torch.gather(s, 2, tgt.clamp(0, W - 1).unsqueeze(-1)).squeeze(-1)
        loss_acc[0] += ((lse - picked) * valid).sum()
        loss_acc[1] += valid.sum()
        if dscores is not None:
            g = torch.softmax(s, dim=-1)
            g.scatter_add_(2, tgt.clamp(0, W - 1).unsqueeze(-1), -torch.ones(B, T, 1))
            g = g * valid.unsqueeze(-1)
            dscores.zero_()
            dscores.view(B, T, -1)[:, :, :W] = g.to(dscores.dtype)

    def set_seed_source(self, seed_dev):
        _SEED_SRC[0] = seed_dev

    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, bc1, bc2, clip, count, grad_scale, dstA, dstB, dstC, packT, packF, update=True,
                  step_dev=None, lr_dev=None):
        self.calls += 1
        if update and step_dev is not None:
            t = float(step_dev[0])
            bc1, bc2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
        if update and lr_dev is not None:
            lr = float(lr_dev[0])
        if update:
            gs = grad_scale / max(float(count[0]), 1.0) if count is not None else grad_scale
            gv = g * gs
            if clip > 0:
                gv = gv.clamp(-clip, clip)
            m.mul_(beta1).add_(gv, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(gv, gv, value=1 - beta2)
            p.sub_((lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps))
        for dst, pack in ((dstA, packT), (dstB, packT), (dstC, packF)):
            if dst is not None:
                ok = dst >= 0
                pack[dst[ok].long()] = p[ok].to(pack.dtype)

    def cast2d(self, src, dst, cols):
        self.calls += 1
        dst.zero_()
        dst[:, :cols] = src[:, :cols].to(dst.dtype)

    def accum_f32(self, src, dst):
        self.calls += 1
        dst.add_(src.float())

    def colsum(self, x, out, cols):
        self.calls += 1
        out.add_(x[:, :cols].float().sum(0).view_as(out))

    def greedy_select(self, scores, W, output, second, captions, masks, done, margins, B, step, Tmax, V, E, has_facts, end_tok):
        self.calls += 1
        for b in range(B):
            if int(done[b]):
                continue
            top = torch.topk(scores[b, :W], 2)
            i1, i2 = int(top.indices[0]), int(top.indices[1])
            output[b, step] = i1
            if margins is not None:
                margins[b, step] = top.values[0] - top.values[1]
            if i1 == end_tok:
                done[b] = 1
                continue
            second[b, step] = i2
            for dupl in (0, 2, 4):
                if step > dupl:
                    n = (dupl + 2) // 2
                    if all(int(output[b, step - k]) == int(output[b, step - n - k]) for k in range(n)):
                        for k in range(max(1, dupl)):
                            output[b, step - k] = int(second[b, step - k])
                        break
            if step < Tmax - 1:
                o = int(output[b, step])
                captions[b, step + 1] = o
                masks[b, step + 1] = 2 if (has_facts and o >= V + E) else (1 if o >= V else 0)

    def beam_select(self, scores, W, cum, ksel, tok_in, mask_in, tok_out, mask_out, anc_in, anc_out, best, result, images, group, step,
                    Tmax, V, E, has_facts, end_tok, pad_tok, workspace=None):
        self.calls += 1
        G = group
        for img in range(images):
            k = int(ksel[img])
            if k <= 0:
                continue
            nrows = 1 if step == 0 else k
            rows = slice(img * G, img * G + nrows)
            cand = (cum[rows].unsqueeze(1) + torch.log_softmax(scores[rows, :W].float(), dim=1)).reshape(-1)
            top = torch.topk(cand, k)
            nalive, bv, bestseq = 0, float(best[img]), None
            first_alive = None
            for r in range(k):
                val, j, c = float(top.values[r]), int(top.indices[r]) // W, int(top.indices[r]) % W
                src = img * G + j
                if c == end_tok:
                    if val > bv:
                        bv, bestseq = val, tok_in[src, 1 : step + 1].tolist() + [c]
                    continue
                dst = img * G + nalive
                tok_out[dst, : step + 1] = tok_in[src, : step + 1]
                mask_out[dst, : step + 1] = mask_in[src, : step + 1]
                anc_out[dst, : step + 1] = anc_in[src, : step + 1]
                cum[dst] = val
                if step + 1 < Tmax:
                    tok_out[dst, step + 1] = c
                    mask_out[dst, step + 1] = 2 if (has_facts and c >= V + E) else (1 if c >= V else 0)
                    anc_out[dst, step + 1] = nalive
                if first_alive is None:
                    first_alive = (val, tok_in[src, 1 : step + 1].tolist() + [c])
                nalive += 1
            if step == Tmax - 1 and bv == float("-inf") and first_alive is not None:
                bv, bestseq = first_alive
            best[img] = bv
            ksel[img] = nalive
            if bestseq is not None:
                result[img] = pad_tok
                result[img, : len(bestseq)] = torch.tensor(bestseq, dtype=result.dtype)

    def image_prep(self, raw, out, mean, std, channels_last=False):
        self.calls += 1
        x = (raw.float() / 255.0).half().float()
        m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
        s = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
        out.copy_(((x - m) / s).to(out.dtype))

    def decode_chain(self, attn_out, x_res, Wo, bo, gamma1, beta1, y, D, ffn=None, proj=None, eps=1e-5):
        self.calls += 1

        def ln(v, g, b):
            v = v[:, :D]
            m = v.mean(1, keepdim=True)
            var = ((v - m) ** 2).mean(1, keepdim=True)
            return (v - m) / torch.sqrt(var + eps) * g.float() + b.float()

        def pad(v, width):
            out = torch.zeros(v.shape[0], width)
            out[:, : v.shape[1]] = v
            return out

        DP = Wo.shape[1]
        s1 = x_res.float() + attn_out.float() @ Wo.float().T + bo.float()
        yy = pad(ln(s1, gamma1, beta1), DP).to(y.dtype).float()
        if ffn is not None:
            W1, b1, W2, b2, g2, be2 = ffn
            h = torch.relu(yy @ W1.float().T + b1.float()).to(y.dtype).float()
            s2 = yy + h @ W2.float().T + b2.float()
            yy = pad(ln(s2, g2, be2), DP).to(y.dtype).float()
        y.copy_(yy.to(y.dtype))
        if proj is not None:
            Wn, bn, out = proj
            out.copy_((yy @ Wn.float().T + bn.float()).to(out.dtype))
