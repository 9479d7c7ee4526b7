"""
Tensor-level wrappers over the C ABI (include/ickb200.h).  Every method takes CUDA tensors that the caller allocated
and launches hand-written sm_100a kernels on torch's current stream; nothing here computes with torch ops.

2-D operands are row-major views: ``x.shape = (rows, cols)``, ``x.stride(0) = ld`` (may exceed cols), ``x.stride(1) = 1``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

F32, BF16 = 0, 1
Drop = Optional[Tuple[float, int, int]]  # (p, seed, site)


def dt_of(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), (t.shape, t.stride())
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _drop(d: Drop) -> Tuple[float, int, int]:
    if d is None:
        return 0.0, 0, 0
    p, seed, site = d
    return float(p), int(seed) & 0xFFFFFFFF, int(site) & 0xFFFFFFFF


class CudaKernels:
    """The product kernel set.  Constructing it loads csrc/libickb200.so or raises."""

    name = "cuda"

    def __init__(self, use_tensor_cores: bool = True):
        import os

        self.lib = _lib.get()
        # ICKB200_NO_TC=1 routes bf16 GEMMs through the CUDA-core kernels (bring-up / bisecting aid)
        self.use_tc = use_tensor_cores and os.environ.get("ICKB200_NO_TC") != "1"
        # bench.py sets prof = [] to bracket every launch with CUDA events on the launching stream and to record the
        # ALGORITHMIC work of the launch (bytes, flops) for the roofline line; None = no instrumentation.
        self.prof = None
        self._ws = None  # fp32 workspace for the split-partials of the tensor-core wgrad
        self._ds_ws = None  # bf16 dS^T workspace of the attention backward

    # ------------------------------------------------------------------------------------------------------------
    def _s(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def _call(self, name, *args, n=1, work=None):
        if self.prof is None:
            self.lib.call(name, *args, self._s())
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.lib.call(name, *args, self._s())
            e1.record()
            self.prof.append((name, e0, e1, work() if work is not None else (0, 0)))

    # ---- dense ---------------------------------------------------------------------------------------------------
    def gemm(self, A, W, C, bias=None, aux=None, epi=0, accumulate=False, drop: Drop = None, force_simt=False):
        """C[M,N] (+)= A[M,K] @ W[N,K]^T + bias, fused epilogue (0 none, 1 relu+dropout, 2 relu/dropout backward)."""
        M, K = A.shape
        N = W.shape[0]
        assert W.shape[1] == K and C.shape[0] == M and C.shape[1] == N, (A.shape, W.shape, C.shape)
        p, seed, site = _drop(drop)
        tc = (self.use_tc and not force_simt and A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
              and A.data_ptr() % 16 == 0 and W.data_ptr() % 16 == 0 and (aux is None or C.dtype == torch.bfloat16))
        work = lambda: ((M * K) * A.element_size() + (N * K) * W.element_size() + M * N * C.element_size() * (2 if accumulate else 1),  # noqa: E731
                        2 * M * N * K)
        if tc:
            self._call("ick_gemm_tn_tc", _p(A), _p(W), _p(C), dt_of(C), _p(bias), _p(aux), M, N, K, _ld(A), _ld(W), _ld(C),
                       _ld(aux) if aux is not None else 0, epi, int(accumulate), p, seed, site, work=work)
        else:
            self._call("ick_gemm_tn_simt", _p(A), dt_of(A), _p(W), dt_of(W), _p(C), dt_of(C), _p(bias), _p(aux), M, N, K,
                       _ld(A), _ld(W), _ld(C), _ld(aux) if aux is not None else 0, epi, int(accumulate), p, seed, site, work=work)

    def gemm_dual(self, A, W0, W1, C, m_split, rows0, bias0=None, bias1=None, aux=None, epi=0, accumulate=False, drop0: Drop = None,
                  drop1: Drop = None):
        """
        Two row groups with their own weights in one launch: rows [0, rows0) use (W0, bias0, drop0), rows [m_split, M) use
        (W1, bias1, drop1; dropout rows counted from m_split); rows [rows0, m_split) are padding (m_split % 128 == 0).
        """
        M, K = A.shape
        N = W0.shape[0]
        assert W0.shape == W1.shape and W0.stride(0) == W1.stride(0) and C.shape[0] == M and C.shape[1] == N and rows0 <= m_split <= M
        p, seed, site0 = _drop(drop0)
        p1, seed1, site1 = _drop(drop1)
        assert (p, seed) == (p1, seed1) or drop0 is None or drop1 is None
        tc = (self.use_tc and A.dtype == torch.bfloat16 and W0.dtype == torch.bfloat16 and A.data_ptr() % 16 == 0
              and W0.data_ptr() % 16 == 0 and W1.data_ptr() % 16 == 0 and (aux is None or C.dtype == torch.bfloat16)
              and m_split % 128 == 0 and 0 < m_split < M)
        if not tc:
            sl = lambda t, a, b: None if t is None else t[a:b]  # noqa: E731
            self.gemm(A[:rows0], W0, C[:rows0], bias0, sl(aux, 0, rows0), epi, accumulate, drop0)
            self.gemm(A[m_split:], W1, C[m_split:], bias1, sl(aux, m_split, M), epi, accumulate, drop1)
            return
        rows = rows0 + (M - m_split)
        work = lambda: ((rows * K) * 2 + 2 * (N * K) * 2 + rows * N * C.element_size() * (2 if accumulate else 1), 2 * rows * N * K)  # noqa: E731
        self._call("ick_gemm_tn_tc_dual", _p(A), _p(W0), _p(W1), _p(C), dt_of(C), _p(bias0), _p(bias1), _p(aux), M, m_split, N, K, _ld(A),
                   _ld(W0), _ld(C), _ld(aux) if aux is not None else 0, epi, int(accumulate), max(p, p1), seed or seed1, site0, site1, work=work)

    def gemm_rowdot(self, A, W, C, O, dsum, S, H, W1=None, m_split=0, rows0=None, dsum1=None, S1=0) -> bool:
        """
        C = A @ W^T (the out-projection input gradient dO of an attention block) with the attention backward's row term
        dsum[(b*H + h)*S + i] = sum_d dO * O written by the GEMM epilogue (include/ickb200.h: ick_gemm_tn_tc_rowdot).  W1 given:
        two row groups as in gemm_dual (rows [m_split, M) use W1, dsum1, S1).  Returns True when dsum was written (tensor-core
        bf16 path) - the caller then passes dsum_ready=True to mha_bwd - and False after a plain GEMM (dsum untouched).
        """
        M, K = A.shape
        N = W.shape[0]
        rows0 = M if rows0 is None else rows0
        tc = (self.use_tc and A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16 and C.dtype == torch.bfloat16 and O.dtype == torch.bfloat16
              and A.data_ptr() % 16 == 0 and W.data_ptr() % 16 == 0 and N >= H * 32 and _ld(O) % 8 == 0 and O.data_ptr() % 16 == 0
              and (W1 is None or (W1.data_ptr() % 16 == 0 and m_split % 128 == 0 and 0 < m_split < M)))
        if not tc:
            if W1 is None:
                self.gemm(A, W, C)
            else:
                self.gemm_dual(A, W, W1, C, m_split, rows0)
            return False
        rows = rows0 + ((M - m_split) if W1 is not None else 0)
        work = lambda: ((rows * K) * 2 + (1 if W1 is None else 2) * (N * K) * 2 + 2 * rows * N * 2 + rows * H // 8 * 4, 2 * rows * N * K)  # noqa: E731
        self._call("ick_gemm_tn_tc_rowdot", _p(A), _p(W), _p(W1), _p(C), _p(O), _p(dsum), _p(dsum1), M, m_split, rows0, N, K, _ld(A), _ld(W),
                   _ld(C), _ld(O), S, S1, H, work=work)
        return True

    def gemm_add_ln(self, A, W, bias, x, s, gamma, beta, y, mean, rstd, d, eps=1e-5, drop: Drop = None):
        """s = x + dropout(A @ W^T + bias); y = LN(s) * gamma + beta; mean / rstd per row (the post-LN sublayer tail with its last
        Linear folded in).  bf16 rows of 320 elements run as ONE tcgen05 launch (ick_gemm_add_ln_tc); anything else as
        gemm + add_ln_fwd."""
        M, K = A.shape
        if self._ln_fusable(A, W, x, s, y, d):
            p, seed, site = _drop(drop)
            self._call("ick_gemm_add_ln_tc", _p(A), _p(W), None, _p(bias), None, _p(x), _p(s), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta),
                       None, None, M, 0, K, d, _ld(A), _ld(W), _ld(x) if x is not None else 0, _ld(s), _ld(y), eps, p, seed, site, 0,
                       work=lambda: ((M * K) * 2 + 320 * K * 2 + M * d * 2 * (3 if x is not None else 2), 2 * M * 320 * K))
            return
        self.gemm(A, W, s, bias=bias)
        self.add_ln_fwd(x, s, gamma, beta, y, mean, rstd, d, eps, drop=drop)

    def gemm_add_ln_dual(self, A, W0, W1, bias0, bias1, x, s, y, mean, rstd, d, m_split, rows0, gammas, betas, drops=(None, None), eps=1e-5):
        """gemm_add_ln for two row groups with their own weights / LayerNorm parameters / dropout sites (rows [0, rows0) and
        [m_split, M); the lockstep entity / fact encoder stacks)."""
        M, K = A.shape
        (p, seed, site0), (p1, seed1, site1) = _drop(drops[0]), _drop(drops[1])
        if (self._ln_fusable(A, W0, x, s, y, d) and W1.data_ptr() % 16 == 0 and W0.stride(0) == W1.stride(0) and m_split % 128 == 0
                and 0 < m_split < M and (p, seed) == (p1, seed1)):
            rows = rows0 + (M - m_split)
            self._call("ick_gemm_add_ln_tc", _p(A), _p(W0), _p(W1), _p(bias0), _p(bias1), _p(x), _p(s), _p(y), _p(mean), _p(rstd),
                       _p(gammas[0]), _p(betas[0]), _p(gammas[1]), _p(betas[1]), M, m_split, K, d, _ld(A), _ld(W0), _ld(x), _ld(s), _ld(y),
                       eps, p, seed, site0, site1,
                       work=lambda: ((rows * K) * 2 + 2 * 320 * K * 2 + rows * d * 2 * 3, 2 * rows * 320 * K))
            return
        self.gemm_dual(A, W0, W1, s, m_split, rows0, bias0, bias1)
        self.add_ln_fwd_dual(x, s, y, mean, rstd, d, rows0, M - m_split, m_split, gammas, betas, drops=drops, eps=eps)

    def _ln_fusable(self, A, W, x, s, y, d) -> bool:
        # OPT-IN (ICK_FUSE_LN=1).  Measured on a B200 (profiles/README.md, r02e): the K train step takes 5.72 ms with the fused
        # kernel against 5.62 ms with GEMM + LayerNorm launches, and the 625-image greedy decode drops from 23.5k to 20.6k
        # captions/s: a 128 x 320 tile per CTA leaves 102 of 148 SMs busy on the decoder rows (5 at decode sizes), the single
        # 320-column TMEM accumulator cannot overlap the epilogue with the next tile's main loop, and the two-sweep row-per-lane
        # epilogue is latency-bound where the stand-alone LayerNorm kernel has a warp per row.
        import os

        return (self.use_tc and os.environ.get("ICK_FUSE_LN", "0") == "1" and A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
                and s.dtype == torch.bfloat16 and y.dtype == torch.bfloat16 and (x is None or x.dtype == torch.bfloat16)
                and W.shape[0] == 320 and d <= 320 and d % 2 == 0 and s.shape[1] == 320 and y.shape[1] == 320
                and all(t.data_ptr() % 16 == 0 and _ld(t) % 8 == 0 for t in (A, W, s, y)))

    def wgrad(self, dY, X, gflat, rowoff, colmap=None, biasoff=None, force_simt=False):
        """gflat[rowoff[n] + colmap[k]] += sum_m dY[m,n] X[m,k];  gflat[biasoff[n]] += sum_m dY[m,n]."""
        M, N = dY.shape
        K = X.shape[1]
        assert X.shape[0] == M and rowoff.numel() >= N and (colmap is None or colmap.numel() >= K)
        tc = (self.use_tc and not force_simt and dY.dtype == torch.bfloat16 and X.dtype == torch.bfloat16
              and dY.data_ptr() % 16 == 0 and X.data_ptr() % 16 == 0)
        work = lambda: ((M * N) * dY.element_size() + (M * K) * X.element_size() + N * K * 4, 2 * M * N * K)  # noqa: E731
        if tc:
            if self._ws is None or self._ws.device != dY.device:
                self._ws = torch.empty(96 << 20, dtype=torch.uint8, device=dY.device)
            self._call("ick_wgrad_tc", _p(dY), _p(X), _p(gflat), _p(rowoff), _p(colmap), _p(biasoff), M, N, K, _ld(dY), _ld(X),
                       _p(self._ws), self._ws.numel(),
                       # launches: wgrad + reduce; the bias gradient is a third kernel only when it cannot ride along (K > 480)
                       n=3 if (biasoff is not None and (K + 63) // 64 * 64 + 32 > 512) else 2, work=work)
        else:
            self._call("ick_wgrad_simt", _p(dY), dt_of(dY), _p(X), dt_of(X), _p(gflat), _p(rowoff), _p(colmap), _p(biasoff), M, N,
                       K, _ld(dY), _ld(X), work=work)

    def wgrad_group(self, probs, gflat):
        """
        Weight gradients of several linear layers at once (one Transformer layer's worth): probs = [(dY, X, rowoff, colmap,
        biasoff), ...], same meaning as wgrad().  bf16 operands go through ONE grouped tensor-core launch (+ one reduce) per 8
        problems; anything else falls back to one wgrad() per problem.
        """
        import ctypes

        probs = list(probs)
        ok = self.use_tc and all(dY.dtype == torch.bfloat16 and X.dtype == torch.bfloat16 and dY.data_ptr() % 16 == 0
                                 and X.data_ptr() % 16 == 0 for dY, X, *_ in probs)
        if not ok or len(probs) < 2:
            for dY, X, rowoff, colmap, biasoff in probs:
                self.wgrad(dY, X, gflat, rowoff, colmap, biasoff)
            return
        dev = probs[0][0].device
        if self._ws is None or self._ws.device != dev:
            self._ws = torch.empty(96 << 20, dtype=torch.uint8, device=dev)
        for i in range(0, len(probs), 8):
            chunk = probs[i:i + 8]
            n = len(chunk)
            vp, ip = ctypes.c_void_p * n, ctypes.c_int * n
            a_dy = vp(*[_p(c[0]) for c in chunk])
            a_x = vp(*[_p(c[1]) for c in chunk])
            a_ro = vp(*[_p(c[2]) for c in chunk])
            a_cm = vp(*[_p(c[3]) for c in chunk])
            a_bo = vp(*[_p(c[4]) for c in chunk])
            a_m = ip(*[c[0].shape[0] for c in chunk])
            a_n = ip(*[c[0].shape[1] for c in chunk])
            a_k = ip(*[c[1].shape[1] for c in chunk])
            a_ldy = ip(*[_ld(c[0]) for c in chunk])
            a_ldx = ip(*[_ld(c[1]) for c in chunk])
            work = lambda ch=chunk: (sum(c[0].numel() * 2 + c[1].numel() * 2 + c[0].shape[1] * c[1].shape[1] * 4 for c in ch),  # noqa: E731
                                     sum(2 * c[0].shape[0] * c[0].shape[1] * c[1].shape[1] for c in ch))
            self._call("ick_wgrad_group_tc", n, ctypes.addressof(a_dy), ctypes.addressof(a_x), ctypes.addressof(a_ro),
                       ctypes.addressof(a_cm), ctypes.addressof(a_bo), ctypes.addressof(a_m), ctypes.addressof(a_n),
                       ctypes.addressof(a_k), ctypes.addressof(a_ldy), ctypes.addressof(a_ldx), _p(gflat), _p(self._ws),
                       self._ws.numel(), n=2, work=work)

    # ---- attention -------------------------------------------------------------------------------------------------
    def mha_fwd(self, Q, K, V, O, lse, B, H, Sq, Sk, dh, causal=False, drop: Drop = None):
        p, seed, site = _drop(drop)
        c = 0.5 if causal else 1.0
        self._call("ick_mha_fwd", _p(Q), _p(K), _p(V), _p(O), _p(lse), dt_of(Q), B, H, Sq, Sk, dh, _ld(Q), _ld(K), _ld(V), _ld(O),
                   int(causal), p, seed, site,
                   work=lambda: (B * H * dh * (2 * Sq + 2 * Sk) * Q.element_size() + B * H * Sq * 4, int(4 * B * H * Sq * Sk * dh * c)))

    def mha_bwd(self, Q, K, V, O, dO, lse, dsum, dQ, dK, dV, B, H, Sq, Sk, dh, causal=False, drop: Drop = None, dsum_ready=False):
        p, seed, site = _drop(drop)
        ws = None
        if Q.dtype == torch.bfloat16:
            # dS^T workspace of the dQ-from-dS scheme (1 KiB per 16-key x 32-query block); grown on demand, shared by all calls
            need = B * H * ((Sk + 15) // 16) * 2 * ((Sq + 63) // 64) * 1024
            if self._ds_ws is None or self._ds_ws.numel() < need or self._ds_ws.device != Q.device:
                self._ds_ws = torch.empty(need, dtype=torch.uint8, device=Q.device)
            ws = self._ds_ws
        self._call("ick_mha_bwd", _p(Q), _p(K), _p(V), _p(O), _p(dO), _p(lse), _p(dsum), _p(dQ), _p(dK), _p(dV), dt_of(Q), B, H, Sq,
                   Sk, dh, _ld(Q), _ld(K), _ld(V), _ld(O), _ld(dO), _ld(dQ), _ld(dK), _ld(dV), int(causal), p, seed, site, _p(ws),
                   ws.numel() if ws is not None else 0, int(dsum_ready),
                   work=lambda: (B * H * dh * (4 * Sq + 4 * Sk) * Q.element_size() + 2 * B * H * Sq * 4,
                                 int(10 * B * H * Sq * Sk * dh * (0.5 if causal else 1.0))))

    def mha_decode(self, Q, K, V, O, B, H, dh, kbatch_stride, vbatch_stride, klen):
        self._call("ick_mha_decode", _p(Q), _p(K), _p(V), _p(O), dt_of(Q), B, H, dh, _ld(Q), _ld(K), _ld(V), _ld(O),
                   int(kbatch_stride), int(vbatch_stride), klen)

    def mha_decode_beam(self, Q, K, V, O, rows, group, H, dh, klen, kimg_stride=0, vimg_stride=0, anc=None, kpos_stride=0, vpos_stride=0):
        """Beam-search decode attention (include/ickb200.h): anc=None -> the `group` beams of an image share its keys (image
        strides), else position-major cache with the (rows, Tmax) int32 ancestor-slot table `anc` (position strides)."""
        self._call("ick_mha_decode_beam", _p(Q), _p(K), _p(V), _p(O), dt_of(Q), rows, group, H, dh, _ld(Q), _ld(K), _ld(V), _ld(O),
                   int(kimg_stride), int(vimg_stride), klen, _p(anc), anc.stride(0) if anc is not None else 0, int(kpos_stride),
                   int(vpos_stride))

    # ---- residual + dropout + layer norm ---------------------------------------------------------------------------
    def add_ln_fwd(self, x, sub, gamma, beta, y, mean, rstd, d, eps=1e-5, rowmap=(0, 0, 0), drop: Drop = None):
        rows = sub.shape[0]
        p, seed, site = _drop(drop)
        self._call("ick_add_ln_fwd", _p(x), _p(sub), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), dt_of(sub), rows, d,
                   _ld(x) if x is not None else 0, _ld(sub), _ld(y), eps, rowmap[0], rowmap[1], rowmap[2], p, seed, site,
                   work=lambda: (rows * d * sub.element_size() * (4 if x is not None else 3), 0))

    def add_ln_bwd(self, dy, s, mean, rstd, gamma, dres, dsub, dgamma, dbeta, d, rowmap=(0, 0, 0), acc_res=False, drop: Drop = None):
        rows = s.shape[0]
        p, seed, site = _drop(drop)
        self._call("ick_add_ln_bwd", _p(dy), _p(s), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dsub), _p(dgamma), _p(dbeta),
                   dt_of(s), rows, d, _ld(dy), _ld(s), _ld(dres) if dres is not None else 0, _ld(dsub) if dsub is not None else 0,
                   rowmap[0], rowmap[1], rowmap[2], int(acc_res), p, seed, site,
                   work=lambda: (rows * d * s.element_size() * (4 + (1 if acc_res else 0)), 0))

    # Two row groups (rows [0, rows0) and [row1, row1 + rows1) of the same buffers) with their own LayerNorm parameters, dropout
    # sites and row maps in one launch; y / dy are the concatenated buffer (identity map) or the row-mapped target of both groups.
    def add_ln_fwd_dual(self, x, sub, y, mean, rstd, d, rows0, rows1, row1, gammas, betas, rowmaps=((0, 0, 0), (0, 0, 0)), drops=(None, None),
                        eps=1e-5):
        (p, seed, site0), (p1, seed1, site1) = _drop(drops[0]), _drop(drops[1])
        if not (d == 300 and _ld(sub) == 320 and _ld(x) == 320 and _ld(y) == 320 and (p, seed) == (p1, seed1)):
            for i, (r0, n) in enumerate(((0, rows0), (row1, rows1))):
                mapped = rowmaps[i][0] != 0
                self.add_ln_fwd(x[r0:r0 + n], sub[r0:r0 + n], gammas[i], betas[i], y if mapped else y[r0:r0 + n], mean[r0:r0 + n],
                                rstd[r0:r0 + n], d, eps, rowmaps[i], drops[i])
            return
        rows = rows0 + rows1
        self._call("ick_add_ln_fwd_dual", _p(x), _p(sub), _p(y), _p(mean), _p(rstd), dt_of(sub), d, 320, eps, rows0, rows1, row1,
                   _p(gammas[0]), _p(betas[0]), _p(gammas[1]), _p(betas[1]), *rowmaps[0], *rowmaps[1], p, seed, site0, site1,
                   work=lambda: (rows * d * sub.element_size() * 4, 0))

    def add_ln_bwd_dual(self, dy, s, mean, rstd, dres, dsub, d, rows0, rows1, row1, gammas, dgammas, dbetas, rowmaps=((0, 0, 0), (0, 0, 0)),
                        drops=(None, None), acc_res=False):
        (p, seed, site0), (p1, seed1, site1) = _drop(drops[0]), _drop(drops[1])
        if not (d == 300 and _ld(s) == 320 and _ld(dy) == 320 and _ld(dres) == 320 and _ld(dsub) == 320 and (p, seed) == (p1, seed1)):
            for i, (r0, n) in enumerate(((0, rows0), (row1, rows1))):
                mapped = rowmaps[i][0] != 0
                self.add_ln_bwd(dy if mapped else dy[r0:r0 + n], s[r0:r0 + n], mean[r0:r0 + n], rstd[r0:r0 + n], gammas[i], dres[r0:r0 + n],
                                dsub[r0:r0 + n], dgammas[i], dbetas[i], d, rowmaps[i], acc_res, drops[i])
            return
        rows = rows0 + rows1
        self._call("ick_add_ln_bwd_dual", _p(dy), _p(s), _p(mean), _p(rstd), _p(dres), _p(dsub), dt_of(s), d, 320, rows0, rows1, row1,
                   _p(gammas[0]), _p(gammas[1]), _p(dgammas[0]), _p(dbetas[0]), _p(dgammas[1]), _p(dbetas[1]), *rowmaps[0], *rowmaps[1],
                   int(acc_res), p, seed, site0, site1, work=lambda: (rows * d * s.element_size() * (4 + (1 if acc_res else 0)), 0))

    # ---- context preparation ------------------------------------------------------------------------------------------
    def entity_encode_fwd(self, entities, facts, type_emb, word_emb, out, variant, B, E, F, D, ntypes, V):
        self._call("ick_entity_encode_fwd", _p(entities), _p(facts), _p(type_emb), _p(word_emb), _p(out), dt_of(out), variant, B, E,
                   entities.shape[-1], F, D, _ld(out), _ld(word_emb) if word_emb is not None else 0, ntypes, V)

    def entity_encode_bwd(self, dEnt, entities, facts, type_emb, word_emb, gflat, type_off, word_off, dt, variant, B, E, F, D, ntypes, V):
        self._call("ick_entity_encode_bwd", _p(dEnt), _p(entities), _p(facts), _p(type_emb), _p(word_emb), _p(gflat), type_off,
                   word_off, dt, variant, B, E, entities.shape[-1], F, D, _ld(dEnt),
                   _ld(word_emb) if word_emb is not None else 0, ntypes, V)

    def fact_encode_fwd(self, facts, ent_enc, pred_emb, out, B, E, F, D, NP):
        self._call("ick_fact_encode_fwd", _p(facts), _p(ent_enc), _p(pred_emb), _p(out), dt_of(out), B, E, F, D, _ld(out), NP)

    def fact_encode_bwd(self, dFact, facts, dEnt, gflat, pred_off, B, E, F, D, NP):
        self._call("ick_fact_encode_bwd", _p(dFact), _p(facts), _p(dEnt), _p(gflat), pred_off, B, E, F, D, _ld(dFact), NP)

    def caption_embed_fwd(self, captions, masks, word_emb, ent_enc, fact_enc, pe, out, B, Tstride, t0, Tn, V, E, F, D, pad, scale,
                          drop: Drop = None, group=1):
        p, seed, site = _drop(drop)
        self._call("ick_caption_embed_fwd", _p(captions), _p(masks), _p(word_emb), _p(ent_enc), _p(fact_enc), _p(pe), _p(out),
                   dt_of(out), B, Tstride, t0, Tn, V, E, F, D, _ld(out), _ld(word_emb), pad, scale, group, p, seed, site)

    def caption_embed_bwd(self, dX, captions, masks, dEnt, dFact, gflat, word_off, B, T, V, E, F, D, pad, scale, drop: Drop = None):
        p, seed, site = _drop(drop)
        self._call("ick_caption_embed_bwd", _p(dX), _p(captions), _p(masks), _p(dEnt), _p(dFact), _p(gflat), word_off, dt_of(dX), B,
                   T, V, E, F, D, _ld(dX), pad, scale, p, seed, site)

    def pixels_fwd(self, encoder_out, memory, B, D, P, M):
        self._call("ick_pixels_fwd", _p(encoder_out), _p(memory), dt_of(memory), B, D, P, M, _ld(memory))

    def pixels_bwd(self, dmemory, d_encoder_out, B, D, P, M):
        self._call("ick_pixels_bwd", _p(dmemory), _p(d_encoder_out), dt_of(dmemory), B, D, P, M, _ld(dmemory))

    def pool_rows_fwd(self, x, rows, B, C, Hin, Win, Hout, Wout):
        self._call("ick_pool_rows_fwd", _p(x), _p(rows), dt_of(rows), B, C, Hin, Win, Hout, Wout, _ld(rows),
                   work=lambda: (x.numel() * 4 + B * Hout * Wout * C * rows.element_size(), 0))

    def pool_rows_bwd(self, drows, dx, B, C, Hin, Win, Hout, Wout):
        self._call("ick_pool_rows_bwd", _p(drows), _p(dx), dt_of(drows), B, C, Hin, Win, Hout, Wout, _ld(drows),
                   work=lambda: (dx.numel() * 4 + B * Hout * Wout * C * drows.element_size(), 0))

    def image_prep(self, raw, out, mean, std, channels_last=False):
        """raw (N, 3, H, W) fp16 in [0, 255] (the HDF5 storage format) -> out = ((raw / 255 in fp16) - mean[c]) / std[c], fp32 or bf16,
        NCHW or channels-last memory order (include/ickb200.h: ick_image_prep)."""
        import ctypes

        N, C, H, W = raw.shape
        assert raw.dtype == torch.float16 and raw.is_contiguous() and out.numel() == raw.numel()
        m, sd = (ctypes.c_float * C)(*mean), (ctypes.c_float * C)(*std)
        self._call("ick_image_prep", _p(raw), _p(out), dt_of(out), N, C, H * W, ctypes.addressof(m), ctypes.addressof(sd),
                   int(channels_last), work=lambda: (raw.numel() * (2 + out.element_size()), 0))

    # ---- indicators / gate ------------------------------------------------------------------------------------------------
    def fact_first_mention(self, captions, facts, first_t, tmin, B, T, F, V, E, group=1, NP=0):
        self._call("ick_fact_first_mention", _p(captions), _p(facts), _p(first_t), _p(tmin), B, T, F, V, E, group, NP)

    def pred_gate_fwd(self, tmin, facts, WpT, bias, h, gate, hg, B, Tn, t0, F, D, NP, lag, group=1):
        self._call("ick_pred_gate_fwd", _p(tmin), _p(facts), _p(WpT), _p(bias), _p(h), _p(gate), _p(hg), dt_of(gate), B, Tn, t0, F, D,
                   _ld(gate), _ld(WpT), NP, lag, group)

    def gate_mul_bwd(self, dHG, h, gate, dG, dH):
        assert dHG.is_contiguous() and h.is_contiguous() and gate.is_contiguous() and dG.is_contiguous() and dH.is_contiguous()
        self._call("ick_gate_mul_bwd", _p(dHG), _p(h), _p(gate), _p(dG), _p(dH), dt_of(h), h.numel())

    def pred_gate_bwd(self, dG, tmin, facts, gflat, wp_off, B, T, F, D, NP, lag):
        self._call("ick_pred_gate_bwd", _p(dG), _p(tmin), _p(facts), _p(gflat), wp_off, dt_of(dG), B, T, F, D, _ld(dG), NP, lag)

    # ---- pointer heads -----------------------------------------------------------------------------------------------------
    def pointer_fwd(self, h, ctx, w, bias, first_t, scores, B, Tn, t0, S, D, col0, lag, group=1):
        self._call("ick_pointer_fwd", _p(h), _p(ctx), _p(w), _p(bias), _p(first_t), _p(scores), dt_of(h), B, Tn, t0, S, D, _ld(h),
                   _ld(scores), col0, lag, group)

    def pointer_bwd(self, dS, h, ctx, w, first_t, dCtx, dH, gflat, w_off, bias_off, B, T, S, D, col0, lag):
        self._call("ick_pointer_bwd", _p(dS), _p(h), _p(ctx), _p(w), _p(first_t), _p(dCtx), _p(dH), _p(gflat), w_off, bias_off,
                   dt_of(h), B, T, S, D, _ld(h), _ld(dS), col0, lag, n=2)

    # ---- loss / optimizer / misc ---------------------------------------------------------------------------------------------
    def ce(self, scores, captions_sorted, decode_len, loss_acc, dscores, B, T, W, pad):
        self._call("ick_ce_fwd_bwd", _p(scores), _p(captions_sorted), _p(decode_len), _p(loss_acc), _p(dscores),
                   dt_of(dscores) if dscores is not None else F32, B, T, W, _ld(scores), _ld(dscores) if dscores is not None else W, pad,
                   work=lambda: (B * T * W * (4 + (dscores.element_size() if dscores is not None else 0)), 0))

    def set_seed_source(self, seed_dev):
        """int32/uint32 device tensor whose value is added to every dropout seed at kernel run time (None = off)."""
        rc = self.lib.fn["ick_set_seed_source"](_p(seed_dev))
        assert rc == 0

    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, bc1, bc2, clip, count, grad_scale, dstA, dstB, dstC, packT, packF,
                  update=True, step_dev=None, lr_dev=None):
        self._call("ick_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, bc1, bc2, clip, _p(count),
                   grad_scale, _p(dstA), _p(dstB), _p(dstC), _p(packT), dt_of(packT), _p(packF), int(update), _p(step_dev), _p(lr_dev),
                   work=lambda: (p.numel() * (4 * (7 if update else 1) + 12 + 2 * packT.element_size()), 0))

    def cast2d(self, src, dst, cols):
        rows = src.shape[0]
        self._call("ick_cast2d", _p(src), dt_of(src), _p(dst), dt_of(dst), rows, cols, _ld(src), _ld(dst))

    def accum_f32(self, src, dst):
        assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
        self._call("ick_accum_f32", _p(src), dt_of(src), _p(dst), src.numel())

    def colsum(self, x, out, cols):
        self._call("ick_colsum", _p(x), dt_of(x), _p(out), x.shape[0], cols, _ld(x))

    def greedy_select(self, scores, W, output, second, captions, masks, done, margins, B, step, Tmax, V, E, has_facts, end_tok):
        self._call("ick_greedy_select", _p(scores), W, _ld(scores), _p(output), _p(second), _p(captions), _p(masks), _p(done),
                   _p(margins), B, step, Tmax, V, E, int(has_facts), end_tok)

    def beam_select(self, scores, W, cum, ksel, tok_in, mask_in, tok_out, mask_out, anc_in, anc_out, best, result, images, group, step,
                    Tmax, V, E, has_facts, end_tok, pad_tok, workspace=None):
        if workspace is None:
            workspace = torch.empty(images * group * group * 2, dtype=torch.float32, device=scores.device)
        self._call("ick_beam_select", _p(scores), W, _ld(scores), _p(cum), _p(ksel), _p(tok_in), _p(mask_in), _p(tok_out), _p(mask_out),
                   _p(anc_in), _p(anc_out), _p(best), _p(result), images, group, step, Tmax, V, E, int(has_facts), end_tok, pad_tok,
                   _p(workspace), workspace.numel() * workspace.element_size(), n=2)

    def decode_chain(self, attn_out, x_res, Wo, bo, gamma1, beta1, y, D, ffn=None, proj=None, eps=1e-5):
        """One launch for the row-wise tail of a decoder layer in the decode loops (include/ickb200.h: ick_decode_chain):
        y = LN(x_res + attn_out Wo^T + bo); ffn = (W1, b1, W2, b2, gamma2, beta2) adds the feed-forward sublayer;
        proj = (Wn, bn, out) also writes out = y Wn^T + bn.  bf16 only."""
        assert attn_out.dtype == torch.bfloat16 and Wo.dtype == torch.bfloat16, "ick_decode_chain is a bf16 kernel"
        rows, DP = attn_out.shape[0], Wo.shape[1]
        W1 = b1 = W2 = b2 = g2 = be2 = None
        FFP = DP
        if ffn is not None:
            W1, b1, W2, b2, g2, be2 = ffn
            FFP = W1.shape[0]
        Wn = bn = out = None
        Nn = 0
        if proj is not None:
            Wn, bn, out = proj
            Nn = Wn.shape[0]
        self._call("ick_decode_chain", _p(attn_out), _ld(attn_out), _p(x_res), _ld(x_res), _p(Wo), _ld(Wo), _p(bo), _p(gamma1), _p(beta1),
                   _p(W1), _ld(W1) if W1 is not None else 0, _p(b1), _p(W2), _ld(W2) if W2 is not None else 0, _p(b2), _p(g2), _p(be2),
                   _p(Wn), _ld(Wn) if Wn is not None else 0, _p(bn), _p(y), _ld(y), _p(out), _ld(out) if out is not None else 0, rows, D, DP,
                   FFP, Nn, eps)
