"""
GPU: every C-ABI kernel (through ickb200.kernels.CudaKernels -> libickb200.so) against the host simulation
(tests/hostsim.py, itself checked against the oracle by test_host_wiring.py) on identical seeded inputs.
Integer / index outputs must match exactly; floating point within the tolerance written next to each check
(fp32: 1e-4 of the output's max magnitude; bf16 storage: 2e-2, north_star).
"""
import math
import os

import numpy as np
import pytest
import torch

from hostsim import FIRST_NONE, HostKernels

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
DTYPES = [torch.float32, torch.bfloat16]
BIG_SEED = 1234003703  # a 31-bit seed as Trainer / the module derive them (initial_seed * 1000003 + step): the toy seeds hide wrap-around bugs


DRYRUN = os.environ.get("ICK_DRYRUN") == "1"  # debugging aid for the test code itself on a GPU-less box
DEV = "cpu" if DRYRUN else "cuda"
if DRYRUN:
    torch.Tensor.cuda = lambda self, *a, **k: self


@pytest.fixture(scope="module")
def K():
    if DRYRUN:
        return HostKernels()
    from ickb200.kernels import CudaKernels

    return CudaKernels()


@pytest.fixture(scope="module")
def Hk():
    return HostKernels()


def g(seed):
    return torch.Generator().manual_seed(seed)


def rnd(shape, dtype, seed, scale=1.0):
    return (torch.randn(*shape, generator=g(seed)) * scale).to(dtype)


def cu(x):
    return None if x is None else x.cuda()


def err(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def headify(x, H, dh):
    """zero the two pad lanes of every 32-wide head block"""
    R = x.shape[0]
    v = x.view(R, -1, 32).clone()
    v[:, :, dh:] = 0
    return v.view(R, -1)


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K_", [(300, 960, 320), (128, 320, 512), (77, 57, 320), (513, 1000, 320), (260, 320, 1000), (1, 960, 320)])
def test_gemm_plain_bias(K, Hk, tc, dtype, M, N, K_):
    if tc and dtype != torch.bfloat16:
        pytest.skip("tensor-core path is bf16")
    A, W = rnd((M, K_), dtype, 1), rnd((N, K_), dtype, 2, 0.1)
    bias = rnd((N,), torch.float32, 3)
    for out_dtype in ([dtype, torch.float32] if dtype == torch.bfloat16 else [dtype]):
        ldc = N + 5
        Cr = torch.zeros(M, ldc, dtype=out_dtype)
        Cg = torch.zeros(M, ldc, dtype=out_dtype).cuda()
        Hk.gemm(A, W, Cr[:, :N], bias=bias)
        K.gemm(cu(A), cu(W), Cg[:, :N], bias=cu(bias), force_simt=not tc)
        assert err(Cg[:, :N], Cr[:, :N]) < TOL[out_dtype]
        assert float(Cg[:, N:].abs().max()) == 0.0  # columns beyond N untouched


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_epilogues(K, Hk, tc, dtype):
    if tc and dtype != torch.bfloat16:
        pytest.skip("tensor-core path is bf16")
    M, N, K_ = 333, 512, 320
    A, W, bias = rnd((M, K_), dtype, 1), rnd((N, K_), dtype, 2, 0.1), rnd((N,), torch.float32, 3)
    drop = (0.3, BIG_SEED, 77)
    # relu + dropout
    Cr, Cg = torch.zeros(M, N, dtype=dtype), torch.zeros(M, N, dtype=dtype).cuda()
    Hk.gemm(A, W, Cr, bias=bias, epi=1, drop=drop)
    K.gemm(cu(A), cu(W), Cg, bias=cu(bias), epi=1, drop=drop, force_simt=not tc)
    assert err(Cg, Cr) < TOL[dtype]
    assert torch.equal(Cg.cpu() == 0, Cr == 0)  # identical dropout mask and ReLU pattern
    # relu/dropout backward against the saved activation
    dY, W2 = rnd((M, 320), dtype, 5), rnd((N, 320), dtype, 6, 0.1)
    Dr, Dg = torch.zeros(M, N, dtype=dtype), torch.zeros(M, N, dtype=dtype).cuda()
    Hk.gemm(dY, W2, Dr, aux=Cr, epi=2, drop=(0.3, 0, 0))
    K.gemm(cu(dY), cu(W2), Dg, aux=cu(Cr), epi=2, drop=(0.3, 0, 0), force_simt=not tc)
    assert err(Dg, Dr) < TOL[dtype]
    # accumulate
    C0 = rnd((M, N), dtype, 7)
    Er, Eg = C0.clone(), C0.clone().cuda()
    Hk.gemm(A, W, Er, accumulate=True)
    K.gemm(cu(A), cu(W), Eg, accumulate=True, force_simt=not tc)
    assert err(Eg, Er) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows0,rows1,N,K_", [(301 * 3, 51 * 3, 960, 320), (128, 7, 320, 512), (1000, 300, 512, 320), (5, 400, 320, 320)])
def test_gemm_dual(K, Hk, dtype, rows0, rows1, N, K_):
    """Two row groups (entity rows | pad to 128 | fact rows) with their own weights, biases and dropout sites in one launch."""
    m_split = (rows0 + 127) // 128 * 128
    M = m_split + rows1
    A = rnd((M, K_), dtype, 1)
    W0, W1 = rnd((N, K_), dtype, 2, 0.1), rnd((N, K_), dtype, 3, 0.1)
    b0, b1 = rnd((N,), torch.float32, 4), rnd((N,), torch.float32, 5)
    d0, d1 = (0.3, 99, 11), (0.3, 99, 12)
    real = torch.cat([torch.arange(rows0), torch.arange(m_split, M)])
    for epi, acc in ((0, False), (1, False), (0, True)):
        C0 = rnd((M, N), dtype, 7)
        Cr, Cg = C0.clone(), C0.clone().cuda()
        kw = dict(bias0=b0, bias1=b1, epi=epi, accumulate=acc, drop0=d0 if epi else None, drop1=d1 if epi else None)
        Hk.gemm_dual(A, W0, W1, Cr, m_split, rows0, **kw)
        K.gemm_dual(cu(A), cu(W0), cu(W1), Cg, m_split, rows0, **{k: (cu(v) if torch.is_tensor(v) else v) for k, v in kw.items()})
        assert err(Cg[real], Cr[real]) < TOL[dtype]
        if epi == 1:
            assert torch.equal(Cg.cpu()[real] == 0, Cr[real] == 0)  # per-group dropout site and row numbering
    # relu/dropout backward against a saved activation
    aux = torch.relu(rnd((M, N), dtype, 8))
    Dr, Dg = torch.zeros(M, N, dtype=dtype), torch.zeros(M, N, dtype=dtype).cuda()
    Hk.gemm_dual(A, W0, W1, Dr, m_split, rows0, aux=aux, epi=2, drop0=(0.3, 0, 0), drop1=(0.3, 0, 0))
    K.gemm_dual(cu(A), cu(W0), cu(W1), Dg, m_split, rows0, aux=cu(aux), epi=2, drop0=(0.3, 0, 0), drop1=(0.3, 0, 0))
    assert err(Dg[real], Dr[real]) < TOL[dtype]


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K_", [(1000, 960, 320), (203, 320, 512), (4100, 136, 320), (64, 1000, 320)])
def test_wgrad(K, Hk, tc, dtype, M, N, K_):
    if tc and dtype != torch.bfloat16:
        pytest.skip("tensor-core path is bf16")
    dY, X = rnd((M, N), dtype, 1), rnd((M, K_), dtype, 2)
    # scatter maps with holes: every 7th packed row and every 11th packed column is padding
    n_real = [n for n in range(N) if n % 7 != 6]
    k_real = [k for k in range(K_) if k % 11 != 10]
    Ko = len(k_real)
    rowoff = torch.full((N,), -1, dtype=torch.int32)
    rowoff[n_real] = (100 + torch.arange(len(n_real)) * Ko).int()
    colmap = torch.full((K_,), -1, dtype=torch.int32)
    colmap[k_real] = torch.arange(Ko).int()
    bias0 = 100 + len(n_real) * Ko
    biasoff = torch.full((N,), -1, dtype=torch.int32)
    biasoff[n_real] = (bias0 + torch.arange(len(n_real))).int()
    size = bias0 + len(n_real) + 50
    Gr = rnd((size,), torch.float32, 9)
    Gg = Gr.clone().cuda()
    Hk.wgrad(dY, X, Gr, rowoff, colmap, biasoff)
    K.wgrad(cu(dY), cu(X), Gg, cu(rowoff), cu(colmap), cu(biasoff), force_simt=not tc)
    assert err(Gg, Gr) < 5e-4 if dtype == torch.float32 else err(Gg, Gr) < 5e-3  # fp32 accumulation of the same products


def _wgrad_problem(M, N, K_, base, seed):
    """dY, X and scatter maps with holes (every 7th packed row / 11th packed column is padding) writing at `base`."""
    dY, X = rnd((M, N), torch.bfloat16, seed), rnd((M, K_), torch.bfloat16, seed + 1)
    n_real = [n for n in range(N) if n % 7 != 6]
    k_real = [k for k in range(K_) if k % 11 != 10]
    Ko = len(k_real)
    rowoff = torch.full((N,), -1, dtype=torch.int32)
    rowoff[n_real] = (base + torch.arange(len(n_real)) * Ko).int()
    colmap = torch.full((K_,), -1, dtype=torch.int32)
    colmap[k_real] = torch.arange(Ko).int()
    bias0 = base + len(n_real) * Ko
    biasoff = torch.full((N,), -1, dtype=torch.int32)
    biasoff[n_real] = (bias0 + torch.arange(len(n_real))).int()
    return (dY, X, rowoff, colmap, biasoff), bias0 + len(n_real)


@pytest.mark.parametrize("shapes", [
    # one encoder layer's worth at a small row count (QKV, out, FFN1, FFN2 with K = 512 -> two k tiles)
    [(1000, 960, 320), (1000, 320, 320), (1000, 512, 320), (1000, 320, 512)],
    # different row counts per problem, ragged N / K, a problem without bias / colmap, more than 8 problems (two launches)
    [(130, 104, 72), (4000, 320, 320), (77, 136, 504), (640, 64, 64), (3000, 960, 320), (513, 320, 512), (64, 8, 8), (900, 248, 328),
     (1200, 320, 320), (333, 96, 200)],
])
def test_wgrad_group(K, Hk, shapes):
    probs, base = [], 64
    for i, (M, N, K_) in enumerate(shapes):
        pr, base = _wgrad_problem(M, N, K_, base, 10 * i + 1)
        if i % 5 == 3:
            pr = pr[:3] + (pr[3], None)  # no bias gradient
        probs.append(pr)
        base += 17
    Gr = rnd((base + 50,), torch.float32, 9)
    Gg = Gr.clone().cuda()
    Hk.wgrad_group(probs, Gr)
    K.wgrad_group([tuple(cu(t) if t is not None else None for t in pr) for pr in probs], Gg)
    assert err(Gg, Gr) < 5e-3  # fp32 accumulation of the same bf16 products


# ---------------------------------------------------------------------------------------------------------------------------
ATT_CASES = [(2, 10, 301, 301, 30, False), (3, 10, 102, 102, 30, True), (2, 10, 37, 548, 30, False), (1, 4, 5, 5, 32, True),
             (2, 10, 130, 130, 30, True),
             # more than 640 streamed positions: the resident-tile chunk loop of the bf16 kernels (K/V in fwd and dQ, Q/dO in dK-dV)
             (1, 2, 70, 700, 30, False), (1, 2, 700, 700, 30, True)]


# bf16 attention implementations behind ick_mha_fwd / ick_mha_bwd (csrc/attention_mma.cu dispatch): the defaults (mma.sync forward,
# tcgen05 backward) and the alternatives kept as switches
ATT_MODES = [("mma", "tc"), ("tc", "tc"), ("mma", "hybrid"), ("mma", "split")]


@pytest.fixture(params=ATT_MODES, ids=lambda m: f"fwd-{m[0]}_bwd-{m[1]}")
def att_mode(request, monkeypatch):
    monkeypatch.setenv("ICK_ATTN_FWD", request.param[0])
    monkeypatch.setenv("ICK_ATTN_BWD", request.param[1])
    return request.param


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("p", [0.0, 0.5])
@pytest.mark.parametrize("B,H,Sq,Sk,dh,causal", ATT_CASES)
def test_mha_fwd_bwd(K, Hk, att_mode, dtype, p, B, H, Sq, Sk, dh, causal):
    if dtype == torch.float32 and att_mode != ATT_MODES[0]:
        pytest.skip("the fp32 CUDA-core kernels have one implementation")
    ld = 3 * H * 32 + 8
    qkv_q = headify(rnd((B * Sq, H * 32), torch.float32, 1), H, dh).to(dtype)
    qkv_k = headify(rnd((B * Sk, H * 32), torch.float32, 2), H, dh).to(dtype)
    qkv_v = headify(rnd((B * Sk, H * 32), torch.float32, 3), H, dh).to(dtype)
    # Q lives in a wider buffer (leading dimension != width), like the packed QKV projection output
    Qbuf = torch.zeros(B * Sq, ld, dtype=dtype)
    Qbuf[:, : H * 32] = qkv_q
    Q = Qbuf[:, : H * 32]
    drop = (p, BIG_SEED + 99, 5) if p > 0 else None
    Or, lr = torch.zeros(B * Sq, H * 32, dtype=dtype), torch.zeros(B * H * Sq)
    Og, lg = torch.zeros(B * Sq, H * 32, dtype=dtype).cuda(), torch.zeros(B * H * Sq).cuda()
    Hk.mha_fwd(Q, qkv_k, qkv_v, Or, lr, B, H, Sq, Sk, dh, causal, drop)
    Qg = cu(Qbuf)[:, : H * 32]
    K.mha_fwd(Qg, cu(qkv_k), cu(qkv_v), Og, lg, B, H, Sq, Sk, dh, causal, drop)
    assert err(Og, Or) < TOL[dtype]
    assert float((lg.cpu() - lr).abs().max()) < (1e-3 if dtype == torch.float32 else 5e-2)
    # backward (uses the reference forward outputs so that errors do not compound)
    dO = headify(rnd((B * Sq, H * 32), torch.float32, 4), H, dh).to(dtype)
    outs_r = [torch.zeros(B * Sq, H * 32, dtype=dtype), torch.zeros(B * Sk, H * 32, dtype=dtype), torch.zeros(B * Sk, H * 32, dtype=dtype)]
    outs_g = [torch.full_like(o, float("nan")).cuda() for o in outs_r]
    dsr, dsg = torch.zeros(B * H * Sq), torch.zeros(B * H * Sq).cuda()
    Hk.mha_bwd(Q, qkv_k, qkv_v, Or, dO, lr, dsr, *outs_r, B, H, Sq, Sk, dh, causal, drop)
    K.mha_bwd(Qg, cu(qkv_k), cu(qkv_v), cu(Or), cu(dO), cu(lr), dsg, *outs_g, B, H, Sq, Sk, dh, causal, drop)
    for a, b, name in zip(outs_g, outs_r, ["dQ", "dK", "dV"]):
        assert not torch.isnan(a).any(), name  # every element (pads included) is written
        assert err(a, b) < TOL[dtype] * 2, name


@pytest.mark.parametrize("qscale", [4.0, 10.0])
@pytest.mark.parametrize("B,H,Sq,Sk,dh,causal", [(3, 10, 301, 301, 30, False), (4, 10, 102, 548, 30, False), (5, 4, 130, 130, 30, True)])
def test_mha_fwd_sharp_scores_rescaled_accumulator(K, Hk, att_mode, qscale, B, H, Sq, Sk, dh, causal):
    """Large score magnitudes: row maxima that grow by more than 2^8 from one key tile to the next force the tcgen05 forward to
    rescale its TMEM accumulator (the lazy running maximum), and tiles far below the row maximum underflow to exact zeros."""
    dtype = torch.bfloat16
    q = (headify(rnd((B * Sq, H * 32), torch.float32, 1), H, dh) * qscale).to(dtype)
    k = headify(rnd((B * Sk, H * 32), torch.float32, 2), H, dh).to(dtype)
    v = headify(rnd((B * Sk, H * 32), torch.float32, 3), H, dh).to(dtype)
    drop = (0.5, 11, 3)
    Or, lr = torch.zeros(B * Sq, H * 32, dtype=dtype), torch.zeros(B * H * Sq)
    Hk.mha_fwd(q, k, v, Or, lr, B, H, Sq, Sk, dh, causal, drop)
    Og, lg = torch.full((B * Sq, H * 32), float("nan"), dtype=dtype).cuda(), torch.zeros(B * H * Sq).cuda()
    K.mha_fwd(cu(q), cu(k), cu(v), Og, lg, B, H, Sq, Sk, dh, causal, drop)
    assert not torch.isnan(Og).any()
    assert err(Og, Or) < TOL[dtype]
    assert float((lg.cpu() - lr).abs().max()) < 5e-2
    # and the backward on top of it (the probabilities are recomputed from the saved LSE)
    dO = headify(rnd((B * Sq, H * 32), torch.float32, 4), H, dh).to(dtype)
    outs_r = [torch.zeros(B * Sq, H * 32, dtype=dtype), torch.zeros(B * Sk, H * 32, dtype=dtype), torch.zeros(B * Sk, H * 32, dtype=dtype)]
    outs_g = [torch.full_like(o, float("nan")).cuda() for o in outs_r]
    dsr, dsg = torch.zeros(B * H * Sq), torch.zeros(B * H * Sq).cuda()
    Hk.mha_bwd(q, k, v, Or, dO, lr, dsr, *outs_r, B, H, Sq, Sk, dh, causal, drop)
    K.mha_bwd(cu(q), cu(k), cu(v), cu(Or), cu(dO), cu(lr), dsg, *outs_g, B, H, Sq, Sk, dh, causal, drop)
    for a, b, name in zip(outs_g, outs_r, ["dQ", "dK", "dV"]):
        assert not torch.isnan(a).any(), name
        assert err(a, b) < TOL[dtype] * 2, name


@pytest.mark.parametrize("p", [0.0, 0.3, 0.5])
@pytest.mark.parametrize("B,H,Sq,Sk,dh,causal", [(100, 10, 70, 70, 30, True), (30, 10, 37, 548, 30, False), (16, 10, 301, 301, 30, False),
                                                 (20, 10, 52, 598, 30, False), (9, 3, 200, 40, 32, False), (160, 10, 51, 51, 30, False)])
def test_mha_bwd_fused_many_items(K, Hk, att_mode, p, B, H, Sq, Sk, dh, causal):
    if att_mode[1] == "split" or (p == 0.3 and att_mode != ATT_MODES[1]):
        pytest.skip("many-item slot re-use concerns the persistent fused kernels; the 14-plane keep words are checked once")
    """The fused bf16 backward (attention_bwd_fused.cu) with several (image, head) items per CTA: operand stages, ring slots and the
    TMEM dQ slots are re-used (1000 items on 148 CTAs: 6-7 tenants per slot), p = 0.3 needs 14 bit planes per keep word."""
    dtype = torch.bfloat16
    q = headify(rnd((B * Sq, H * 32), torch.float32, 1), H, dh).to(dtype)
    k = headify(rnd((B * Sk, H * 32), torch.float32, 2), H, dh).to(dtype)
    v = headify(rnd((B * Sk, H * 32), torch.float32, 3), H, dh).to(dtype)
    drop = (p, 77, 9) if p > 0 else None
    Or, lr = torch.zeros(B * Sq, H * 32, dtype=dtype), torch.zeros(B * H * Sq)
    Hk.mha_fwd(q, k, v, Or, lr, B, H, Sq, Sk, dh, causal, drop)
    Og, lg = torch.zeros(B * Sq, H * 32, dtype=dtype).cuda(), torch.zeros(B * H * Sq).cuda()
    K.mha_fwd(cu(q), cu(k), cu(v), Og, lg, B, H, Sq, Sk, dh, causal, drop)
    assert err(Og, Or) < TOL[dtype]
    dO = headify(rnd((B * Sq, H * 32), torch.float32, 4), H, dh).to(dtype)
    outs_r = [torch.zeros(B * Sq, H * 32, dtype=dtype), torch.zeros(B * Sk, H * 32, dtype=dtype), torch.zeros(B * Sk, H * 32, dtype=dtype)]
    outs_g = [torch.full_like(o, float("nan")).cuda() for o in outs_r]
    dsr, dsg = torch.zeros(B * H * Sq), torch.zeros(B * H * Sq).cuda()
    Hk.mha_bwd(q, k, v, Or, dO, lr, dsr, *outs_r, B, H, Sq, Sk, dh, causal, drop)
    K.mha_bwd(cu(q), cu(k), cu(v), cu(Or), cu(dO), cu(lr), dsg, *outs_g, B, H, Sq, Sk, dh, causal, drop)
    for a, b, name in zip(outs_g, outs_r, ["dQ", "dK", "dV"]):
        assert not torch.isnan(a).any(), name
        # per (image, head) item: a wrong tenant in a re-used slot would corrupt single items, which a global max norm could hide
        ai = a.float().cpu().view(B, -1, H, 32).permute(0, 2, 1, 3).reshape(B * H, -1)
        bi = b.float().view(B, -1, H, 32).permute(0, 2, 1, 3).reshape(B * H, -1)
        rel = (ai - bi).abs().amax(1) / bi.abs().amax(1).clamp_min(1e-6)
        assert float(rel.max()) < TOL[dtype] * 3, (name, int(rel.argmax()), float(rel.max()))


@pytest.mark.parametrize("B,H,Sq,Sk,dh", [(40, 10, 5, 548, 30), (7, 10, 5, 598, 30), (33, 10, 3, 300, 30), (9, 10, 16, 548, 30), (400, 10, 1, 257, 30)])
def test_mha_fwd_few_queries_many_items(K, Hk, monkeypatch, B, H, Sq, Sk, dh):
    """One 16-row slab of queries per (image, head) item - the beam-search decode step's cross-attention (Sq = beam width) - with more
    items than CTAs, so that pipeline stages are re-used; checked per item."""
    monkeypatch.setenv("ICK_ATTN_FWD", "mma")
    dtype = torch.bfloat16
    q = headify(rnd((B * Sq, H * 32), torch.float32, 1, 2.0), H, dh).to(dtype)
    k = headify(rnd((B * Sk, H * 32), torch.float32, 2, 2.0), H, dh).to(dtype)
    v = headify(rnd((B * Sk, H * 32), torch.float32, 3), H, dh).to(dtype)
    Or, Og = torch.zeros(B * Sq, H * 32, dtype=dtype), torch.full((B * Sq, H * 32), float("nan"), dtype=dtype).cuda()
    lr, lg = torch.zeros(B * H * Sq), torch.zeros(B * H * Sq).cuda()
    Hk.mha_fwd(q, k, v, Or, lr, B, H, Sq, Sk, dh, False, None)
    K.mha_fwd(cu(q), cu(k), cu(v), Og, lg, B, H, Sq, Sk, dh, False, None)
    assert not torch.isnan(Og).any()
    ai = Og.float().cpu().view(B, Sq, H, 32).permute(0, 2, 1, 3).reshape(B * H, -1)
    bi = Or.float().view(B, Sq, H, 32).permute(0, 2, 1, 3).reshape(B * H, -1)
    rel = (ai - bi).abs().amax(1) / bi.abs().amax(1).clamp_min(1e-6)
    assert float(rel.max()) < TOL[dtype], (int(rel.argmax()), float(rel.max()))
    assert err(lg, lr) < 1e-3


@pytest.mark.parametrize("dtype", DTYPES)
def test_mha_decode(K, Hk, dtype):
    B, H, dh, Tmax, klen = 5, 10, 30, 12, 7
    ldc = 3 * H * 32
    cache = headify(rnd((B * Tmax, ldc), torch.float32, 1), 3 * H, dh).to(dtype)
    Q = cache.view(B, Tmax, ldc)[:, klen - 1, : H * 32]
    Or, Og = torch.zeros(B, H * 32, dtype=dtype), torch.zeros(B, H * 32, dtype=dtype).cuda()
    Hk.mha_decode(Q, cache[:, H * 32 : 2 * H * 32], cache[:, 2 * H * 32 :], Or, B, H, dh, Tmax * ldc, Tmax * ldc, klen)
    cg = cu(cache)
    K.mha_decode(cg.view(B, Tmax, ldc)[:, klen - 1, : H * 32], cg[:, H * 32 : 2 * H * 32], cg[:, 2 * H * 32 :], Og, B, H, dh, Tmax * ldc,
                 Tmax * ldc, klen)
    assert err(Og, Or) < TOL[dtype]


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("p", [0.0, 0.4])
@pytest.mark.parametrize("rows,rowmap,yrows", [(301 * 2, (301, 548, 196), 548 * 2), (77, (0, 0, 0), 77)])
def test_add_ln_fwd_bwd(K, Hk, dtype, p, rows, rowmap, yrows):
    d, ld = 300, 320
    x, sub = rnd((rows, ld), dtype, 1), rnd((rows, ld), dtype, 2)
    gamma, beta = 1 + 0.1 * rnd((d,), torch.float32, 3), rnd((d,), torch.float32, 4)
    drop = (p, BIG_SEED, 3) if p > 0 else None
    yr, yg = torch.zeros(yrows, ld, dtype=dtype), torch.full((yrows, ld), 7.0, dtype=dtype).cuda()
    sr, sg = sub.clone(), sub.clone().cuda()
    mr, rr, mg, rg = torch.zeros(rows), torch.zeros(rows), torch.zeros(rows).cuda(), torch.zeros(rows).cuda()
    Hk.add_ln_fwd(x, sr, gamma, beta, yr, mr, rr, d, 1e-5, rowmap, drop)
    K.add_ln_fwd(cu(x), sg, cu(gamma), cu(beta), yg, mg, rg, d, 1e-5, rowmap, drop)
    from hostsim import _rows

    idx = _rows(rowmap, rows)
    assert err(yg[idx.cuda()], yr[idx]) < TOL[dtype]
    assert float(yg[idx.cuda()][:, d:].abs().max()) == 0.0  # pad columns zeroed
    assert err(sg[:, :d], sr[:, :d]) < TOL[dtype]
    assert err(mg, mr) < 1e-3 and err(rg, rr) < 1e-3
    # backward
    dy = rnd((yrows, ld), dtype, 5)
    for acc in (False, True):
        dres_r = rnd((rows, ld), dtype, 6) if acc else torch.zeros(rows, ld, dtype=dtype)
        dres_g = dres_r.clone().cuda()
        dsub_r, dsub_g = torch.zeros(rows, ld, dtype=dtype), torch.full((rows, ld), float("nan"), dtype=dtype).cuda()
        dgr, dbr = rnd((d,), torch.float32, 8), rnd((d,), torch.float32, 9)
        dgg, dbg = dgr.clone().cuda(), dbr.clone().cuda()
        Hk.add_ln_bwd(dy, sr, mr, rr, gamma, dres_r, dsub_r, dgr, dbr, d, rowmap, acc, drop)
        K.add_ln_bwd(cu(dy), cu(sr), cu(mr), cu(rr), cu(gamma), dres_g, dsub_g, dgg, dbg, d, rowmap, acc, drop)
        assert err(dres_g[:, :d], dres_r[:, :d]) < TOL[dtype] * 2
        assert err(dsub_g, dsub_r) < TOL[dtype] * 2
        assert err(dgg, dgr) < 1e-3 and err(dbg, dbr) < 1e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("p", [0.0, 0.5])
@pytest.mark.parametrize("mapped", [False, True])
def test_add_ln_dual(K, Hk, dtype, p, mapped):
    """Two row groups (entity rows | pad | fact rows) with their own gamma/beta, dropout sites and row maps in one launch."""
    d, ld, B, S0, S1, P = 300, 320, 3, 23, 7, 5
    rows0, rows1 = B * S0, B * S1
    row1 = (rows0 + 127) // 128 * 128
    R = row1 + rows1
    M = P + S0 + S1
    x, sub = rnd((R, ld), dtype, 1), rnd((R, ld), dtype, 2)
    gam = [1 + 0.1 * rnd((d,), torch.float32, 3 + i) for i in range(2)]
    bet = [rnd((d,), torch.float32, 5 + i) for i in range(2)]
    drops = ((p, 7, 3), (p, 7, 4)) if p > 0 else (None, None)
    maps = ((S0, M, P), (S1, M, P + S0)) if mapped else ((0, 0, 0), (0, 0, 0))
    yrows = B * M if mapped else R
    yr, yg = torch.zeros(yrows, ld, dtype=dtype), torch.zeros(yrows, ld, dtype=dtype).cuda()
    sr, sg = sub.clone(), sub.clone().cuda()
    mr, rr, mg, rg = torch.zeros(R), torch.zeros(R), torch.zeros(R).cuda(), torch.zeros(R).cuda()
    Hk.add_ln_fwd_dual(x, sr, yr, mr, rr, d, rows0, rows1, row1, gam, bet, maps, drops)
    K.add_ln_fwd_dual(cu(x), sg, yg, mg, rg, d, rows0, rows1, row1, [cu(t) for t in gam], [cu(t) for t in bet], maps, drops)
    real = torch.cat([torch.arange(rows0), torch.arange(row1, R)])
    assert err(yg, yr) < TOL[dtype]
    assert err(sg[real][:, :d], sr[real][:, :d]) < TOL[dtype]
    assert err(mg[real], mr[real]) < 1e-3 and err(rg[real], rr[real]) < 1e-3
    dy = rnd((yrows, ld), dtype, 9)
    for acc in (False, True):
        dres_r = rnd((R, ld), dtype, 6) if acc else torch.zeros(R, ld, dtype=dtype)
        dres_g = dres_r.clone().cuda()
        dsub_r, dsub_g = torch.zeros(R, ld, dtype=dtype), torch.zeros(R, ld, dtype=dtype).cuda()
        dgr, dbr = [rnd((d,), torch.float32, 11 + i) for i in range(2)], [rnd((d,), torch.float32, 13 + i) for i in range(2)]
        dgg, dbg = [t.clone().cuda() for t in dgr], [t.clone().cuda() for t in dbr]
        Hk.add_ln_bwd_dual(dy, sr, mr, rr, dres_r, dsub_r, d, rows0, rows1, row1, gam, dgr, dbr, maps, drops, acc)
        K.add_ln_bwd_dual(cu(dy), cu(sr), cu(mr), cu(rr), dres_g, dsub_g, d, rows0, rows1, row1, [cu(t) for t in gam], dgg, dbg, maps, drops, acc)
        assert err(dres_g[real][:, :d], dres_r[real][:, :d]) < TOL[dtype] * 2
        assert err(dsub_g[real], dsub_r[real]) < TOL[dtype] * 2
        for i in range(2):
            assert err(dgg[i], dgr[i]) < 1e-3 and err(dbg[i], dbr[i]) < 1e-3


def _fused_launched(K, name):
    """the fused entry point really ran (a silent fall-back to the two-kernel path would make these tests vacuous)"""
    return DRYRUN or any(n == name for n, *_ in K.prof)


@pytest.mark.parametrize("p", [0.0, 0.5])
@pytest.mark.parametrize("M,K_,has_x", [(301 * 3, 320, True), (128, 512, True), (1000, 512, True), (77, 320, False), (13056, 320, True)])
def test_gemm_add_ln_fused(K, Hk, monkeypatch, p, M, K_, has_x):
    """Linear + residual + dropout + LayerNorm in one tcgen05 launch (ick_gemm_add_ln_tc) against GEMM -> add_ln on the host."""
    monkeypatch.setenv("ICK_FUSE_LN", "1")  # the fused kernel is opt-in (kernels._ln_fusable)
    dtype, d, ld = torch.bfloat16, 300, 320
    A = rnd((M, K_), dtype, 1)
    W = rnd((ld, K_), dtype, 2, 0.06)
    W[d:] = 0  # d_model padded to 320 rows: pad rows of the packed weight are zero
    bias = torch.cat([rnd((d,), torch.float32, 3), torch.zeros(ld - d)])
    x = rnd((M, ld), dtype, 4) if has_x else None
    if x is not None:
        x[:, d:] = 0
    gamma, beta = 1 + 0.1 * rnd((d,), torch.float32, 5), rnd((d,), torch.float32, 6)
    drop = (p, BIG_SEED, 11) if p > 0 else None
    sr, yr, mr, rr = torch.zeros(M, ld, dtype=dtype), torch.zeros(M, ld, dtype=dtype), torch.zeros(M), torch.zeros(M)
    Hk.gemm_add_ln(A, W, bias, x, sr, gamma, beta, yr, mr, rr, d, drop=drop)
    sg, yg = torch.full((M + 3, ld), 7.0, dtype=dtype).cuda(), torch.full((M + 3, ld), 7.0, dtype=dtype).cuda()  # 3 canary rows
    mg, rg = torch.zeros(M).cuda(), torch.zeros(M).cuda()
    K.prof = []
    K.gemm_add_ln(cu(A), cu(W), cu(bias), cu(x), sg[:M], cu(gamma), cu(beta), yg[:M], mg, rg, d, drop=drop)
    torch.cuda.synchronize() if not DRYRUN else None
    assert _fused_launched(K, "ick_gemm_add_ln_tc")
    K.prof = None
    # the fused kernel normalises the fp32 accumulator, the host reference the bf16-rounded GEMM output: bf16 tolerance
    assert err(sg[:M, :d], sr[:, :d]) < 2e-2 and err(yg[:M, :d], yr[:, :d]) < 2e-2
    assert float(yg[:M, d:].abs().max()) == 0.0 and float(sg[:M, d:].abs().max()) == 0.0  # pad columns are exact zeros
    assert float((sg[M:] - 7.0).abs().max()) == 0.0 and float((yg[M:] - 7.0).abs().max()) == 0.0  # rows behind M untouched
    assert err(mg, mr) < 2e-2 and err(rg, rr) < 2e-2
    # self-consistency at fp32 level: y is LayerNorm of the s the kernel stored (up to the bf16 rounding of s), with its mean / rstd
    sf = sg[:M, :d].float()
    mu, var = sf.mean(-1), sf.var(-1, unbiased=False)
    assert float((mg - mu).abs().max()) < 2e-2 and err(rg, torch.rsqrt(var + 1e-5)) < 2e-2
    y2 = (sf - mg[:, None]) * rg[:, None] * cu(gamma) + cu(beta)
    assert err(yg[:M, :d], y2) < 2e-2


@pytest.mark.parametrize("p", [0.0, 0.5])
def test_gemm_add_ln_fused_dual(K, Hk, monkeypatch, p):
    """two row groups (entity rows | pad | fact rows) with their own weights, LayerNorm parameters and dropout sites"""
    monkeypatch.setenv("ICK_FUSE_LN", "1")
    dtype, d, ld, K_ = torch.bfloat16, 300, 320, 320
    rows0, rows1 = 301 * 2, 51 * 2
    m_split = (rows0 + 127) // 128 * 128
    M = m_split + rows1
    A = rnd((M, K_), dtype, 1)
    Ws = [rnd((ld, K_), dtype, 2 + i, 0.06) for i in range(2)]
    for W in Ws:
        W[d:] = 0
    bs = [torch.cat([rnd((d,), torch.float32, 4 + i), torch.zeros(ld - d)]) for i in range(2)]
    x = rnd((M, ld), dtype, 6)
    x[:, d:] = 0
    gam = [1 + 0.1 * rnd((d,), torch.float32, 7 + i) for i in range(2)]
    bet = [rnd((d,), torch.float32, 9 + i) for i in range(2)]
    drops = ((p, BIG_SEED, 3), (p, BIG_SEED, 4)) if p > 0 else (None, None)
    sr, yr, mr, rr = torch.zeros(M, ld, dtype=dtype), torch.zeros(M, ld, dtype=dtype), torch.zeros(M), torch.zeros(M)
    Hk.gemm_add_ln_dual(A, Ws[0], Ws[1], bs[0], bs[1], x, sr, yr, mr, rr, d, m_split, rows0, gam, bet, drops)
    sg, yg, mg, rg = torch.zeros(M, ld, dtype=dtype).cuda(), torch.zeros(M, ld, dtype=dtype).cuda(), torch.zeros(M).cuda(), torch.zeros(M).cuda()
    K.prof = []
    K.gemm_add_ln_dual(cu(A), cu(Ws[0]), cu(Ws[1]), cu(bs[0]), cu(bs[1]), cu(x), sg, yg, mg, rg, d, m_split, rows0, [cu(t) for t in gam],
                       [cu(t) for t in bet], drops)
    torch.cuda.synchronize() if not DRYRUN else None
    assert _fused_launched(K, "ick_gemm_add_ln_tc")
    K.prof = None
    real = torch.cat([torch.arange(rows0), torch.arange(m_split, M)])
    assert err(sg[real][:, :d], sr[real][:, :d]) < 2e-2 and err(yg[real][:, :d], yr[real][:, :d]) < 2e-2
    assert err(mg[real], mr[real]) < 2e-2 and err(rg[real], rr[real]) < 2e-2
    assert float(yg[real][:, d:].abs().max()) == 0.0


@pytest.mark.parametrize("dual", [False, True])
def test_gemm_rowdot_epilogue(K, Hk, dual):
    """dO = dB W and the attention backward's row term rowsum(dO * O) from the GEMM epilogue (ick_gemm_tn_tc_rowdot)"""
    dtype, H, ld, K_ = torch.bfloat16, 10, 320, 320
    S0, S1, B = 301, 51, 3
    rows0, rows1 = B * S0, B * S1
    m_split = (rows0 + 127) // 128 * 128 if dual else 0
    M = m_split + rows1 if dual else rows0
    A, O = rnd((M, K_), dtype, 1), rnd((M, ld), dtype, 2)
    Ws = [rnd((ld, K_), dtype, 3 + i, 0.06) for i in range(2)]
    Cr, Cg = torch.zeros(M, ld, dtype=dtype), torch.zeros(M, ld, dtype=dtype).cuda()
    dr = [torch.zeros(B * H * S0), torch.zeros(B * H * S1)]
    dg = [torch.full((B * H * S0,), 9.0).cuda(), torch.full((B * H * S1,), 9.0).cuda()]
    kw = dict(W1=Ws[1], m_split=m_split, rows0=rows0, dsum1=dr[1], S1=S1) if dual else {}
    assert Hk.gemm_rowdot(A, Ws[0], Cr, O, dr[0], S0, H, **kw)
    kw = dict(W1=cu(Ws[1]), m_split=m_split, rows0=rows0, dsum1=dg[1], S1=S1) if dual else {}
    assert K.gemm_rowdot(cu(A), cu(Ws[0]), Cg, cu(O), dg[0], S0, H, **kw)
    real = torch.cat([torch.arange(rows0), torch.arange(m_split, M)]) if dual else torch.arange(M)
    assert err(Cg[real], Cr[real]) < 2e-2
    for i in range(2 if dual else 1):
        assert err(dg[i], dr[i]) < 2e-2
    # exact agreement with the standalone row-dot over the dO the kernel itself stored (same bf16 inputs, same summation order)
    prod = (Cg[:rows0].float() * cu(O)[:rows0].float()).view(B, S0, H, 32).sum(-1).permute(0, 2, 1).reshape(-1)
    assert float((dg[0] - prod).abs().max()) <= 1e-5 * max(1.0, float(prod.abs().max()))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,C,Hin,Win,Hout,Wout", [(3, 2048, 8, 8, 14, 14), (2, 100, 7, 7, 14, 14), (1, 64, 10, 6, 4, 5)])
def test_pool_rows(K, Hk, dtype, B, C, Hin, Win, Hout, Wout):
    """AdaptiveAvgPool2d written as the K-major rows of the 1x1-convolution GEMM (Encoder hand-off, G/models.py:43-45)."""
    x = rnd((B, C, Hin, Win), torch.float32, 1)
    ld = (C + 7) // 8 * 8
    rr, rg = torch.zeros(B * Hout * Wout, ld, dtype=dtype), torch.zeros(B * Hout * Wout, ld, dtype=dtype).cuda()
    Hk.pool_rows_fwd(x, rr, B, C, Hin, Win, Hout, Wout)
    K.pool_rows_fwd(cu(x), rg, B, C, Hin, Win, Hout, Wout)
    assert err(rg, rr) < TOL[dtype]


# ---------------------------------------------------------------------------------------------------------------------------
def make_context(variant, B, E, F, V, seed=0):
    from ickb200 import synthetic as syn

    cfg = syn.Config({0: "G", 1: "K", 2: "N"}[variant], B=B, T=14, E=E, F=F, V=V, P=20)
    return cfg, syn.make_batch(cfg, seed=seed)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_entity_fact_caption_kernels(K, Hk, dtype, variant):
    B, E, F, V, D, ld = 3, 23, (0 if variant == 0 else 11), 61, 300, 320
    cfg, batch = make_context(variant, B, E, F, V)
    nf = {0: 4, 1: 6, 2: 5}[variant]
    ntypes = 1000 if variant < 2 else 20
    type_emb = rnd((ntypes, D - nf), torch.float32, 1, 0.1)
    wemb = torch.zeros(V, ld, dtype=dtype)
    wemb[:, :D] = rnd((V, D), dtype, 2, 0.1)
    ent, facts = batch["entities"], batch.get("facts")
    outr, outg = torch.zeros(B * E, ld, dtype=dtype), torch.full((B * E, ld), float("nan"), dtype=dtype).cuda()
    Hk.entity_encode_fwd(ent, facts, type_emb, wemb if variant == 2 else None, outr, variant, B, E, F, D, ntypes, V)
    K.entity_encode_fwd(cu(ent), cu(facts), cu(type_emb), cu(wemb) if variant == 2 else None, outg, variant, B, E, F, D, ntypes, V)
    assert err(outg, outr) < TOL[dtype]
    # backward: type-embedding (and, news, word-embedding) gradients
    n_flat = ntypes * (D - nf) + V * D + 10
    type_off, word_off = 3, 3 + ntypes * (D - nf) + 2
    dEnt = torch.zeros(B * E, ld)
    dEnt[:, :D] = rnd((B * E, D), torch.float32, 3)
    Gr, Gg = torch.zeros(n_flat), torch.zeros(n_flat).cuda()
    dt = 0 if dtype == torch.float32 else 1
    Hk.entity_encode_bwd(dEnt, ent, facts, type_emb, wemb if variant == 2 else None, Gr, type_off, word_off, dt, variant, B, E, F, D, ntypes, V)
    K.entity_encode_bwd(cu(dEnt), cu(ent), cu(facts), cu(type_emb), cu(wemb) if variant == 2 else None, Gg, type_off, word_off, dt, variant, B, E,
                        F, D, ntypes, V)
    assert err(Gg, Gr) < 1e-4
    fact_r = None
    if variant != 0:
        NP = 3000
        pred_emb = rnd((NP, D), torch.float32, 4, 0.1)
        fact_r, fact_g = torch.zeros(B * F, ld, dtype=dtype), torch.full((B * F, ld), float("nan"), dtype=dtype).cuda()
        Hk.fact_encode_fwd(facts, outr, pred_emb, fact_r, B, E, F, D, NP)
        K.fact_encode_fwd(cu(facts), cu(outr), cu(pred_emb), fact_g, B, E, F, D, NP)
        assert err(fact_g, fact_r) < TOL[dtype]
        dFact = torch.zeros(B * F, ld)
        dFact[:, :D] = rnd((B * F, D), torch.float32, 5)
        dEr, dEg = torch.zeros(B * E, ld), torch.zeros(B * E, ld).cuda()
        Pr, Pg = torch.zeros(NP * D + 7), torch.zeros(NP * D + 7).cuda()
        Hk.fact_encode_bwd(dFact, facts, dEr, Pr, 7, B, E, F, D, NP)
        K.fact_encode_bwd(cu(dFact), cu(facts), dEg, Pg, 7, B, E, F, D, NP)
        assert err(dEg, dEr) < 1e-5 and err(Pg, Pr) < 1e-5
    # caption embedder (+ sqrt(d), positional table, dropout), whole sequence and single position
    T = cfg.T
    caps, masks = batch["captions"].clone(), batch["caption_masks"].clone()
    caps[0, 3], masks[0, 3] = V + E + F + 5, 1   # out-of-range pointer -> <unk_ent>
    caps[1, 2], masks[1, 2] = V + 2, 0           # pointer id with mask 0 -> <pad> word row
    pe = rnd((T + 3, D), torch.float32, 6)
    for (t0, Tn), p in (((0, T), 0.0), ((0, T), 0.2), ((5, 1), 0.0)):
        drop = (p, BIG_SEED + 3, 9) if p > 0 else None
        xr, xg = torch.zeros(B * Tn, ld, dtype=dtype), torch.full((B * Tn, ld), float("nan"), dtype=dtype).cuda()
        Hk.caption_embed_fwd(caps, masks, wemb, outr, fact_r, pe, xr, B, T, t0, Tn, V, E, F, D, 0, math.sqrt(D), drop)
        K.caption_embed_fwd(cu(caps), cu(masks), cu(wemb), cu(outr), cu(fact_r), cu(pe), xg, B, T, t0, Tn, V, E, F, D, 0, math.sqrt(D), drop)
        assert err(xg, xr) < TOL[dtype]
    dX = rnd((B * T, ld), dtype, 7)
    drop = (0.2, BIG_SEED + 3, 9)
    dEr, dEg = torch.zeros(B * E, ld), torch.zeros(B * E, ld).cuda()
    dFr = torch.zeros(B * F, ld) if F else None
    dFg = cu(dFr.clone()) if F else None
    Wr, Wg = torch.zeros(V * D + 5), torch.zeros(V * D + 5).cuda()
    Hk.caption_embed_bwd(dX, caps, masks, dEr, dFr, Wr, 5, B, T, V, E, F, D, 0, math.sqrt(D), drop)
    K.caption_embed_bwd(cu(dX), cu(caps), cu(masks), dEg, dFg, Wg, 5, B, T, V, E, F, D, 0, math.sqrt(D), drop)
    assert err(dEg, dEr) < 1e-4 and err(Wg, Wr) < 1e-4
    if F:
        assert err(dFg, dFr) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_pixels(K, Hk, dtype):
    B, D, P, M, ld = 3, 300, 196, 250, 320
    enc = rnd((B, D, P), torch.float32, 1)
    mr, mg = torch.full((B * M, ld), 3.0, dtype=dtype), torch.full((B * M, ld), 3.0, dtype=dtype).cuda()
    Hk.pixels_fwd(enc, mr, B, D, P, M)
    K.pixels_fwd(cu(enc), mg, B, D, P, M)
    assert err(mg, mr) < TOL[dtype] and torch.equal(mg.cpu()[:, D:], mr[:, D:])
    dr, dg = torch.zeros(B, D, P), torch.zeros(B, D, P).cuda()
    Hk.pixels_bwd(mr, dr, B, D, P, M)
    K.pixels_bwd(cu(mr), dg, B, D, P, M)
    assert torch.equal(dg.cpu(), dr)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("lag,Tn,t0", [(0, 14, 0), (1, 1, 6)])
def test_indicators_gate_pointer(K, Hk, dtype, lag, Tn, t0):
    B, E, F, V, D, ld, NP = 4, 23, 17, 61, 300, 320, 3000
    cfg, batch = make_context(1, B, E, F, V, seed=5)
    T = cfg.T
    caps, facts = batch["captions"], batch["facts"].clone()
    facts[:, 3, 2] = facts[:, 1, 2]  # duplicate predicates: the indicator is a SET of predicates (one representative fact each)
    facts[:, 9, 2] = facts[:, 1, 2]
    facts[:, 12, 2] = facts[:, 5, 2]
    ftr, tmr = torch.zeros(B * F, dtype=torch.int32), torch.zeros(B * F, dtype=torch.int32)
    ftg, tmg = ftr.clone().cuda(), tmr.clone().cuda()
    Hk.fact_first_mention(caps, facts, ftr, tmr, B, T, F, V, E)
    K.fact_first_mention(cu(caps), cu(facts), ftg, tmg, B, T, F, V, E)
    assert torch.equal(ftg.cpu(), ftr) and torch.equal(tmg.cpu(), tmr)  # integer work: bit-exact
    ftg2, tmg2 = torch.zeros_like(ftg), torch.zeros_like(tmg)
    K.fact_first_mention(cu(caps), cu(facts), ftg2, tmg2, B, T, F, V, E, NP=NP)  # per-predicate tables instead of the pair search
    assert torch.equal(ftg2.cpu(), ftr) and torch.equal(tmg2.cpu(), tmr)
    assert int((ftr < FIRST_NONE).sum()) > 0
    WpT = torch.zeros(NP, ld)
    WpT[:, :D] = rnd((NP, D), torch.float32, 1, 0.3)
    bias = rnd((D,), torch.float32, 2)
    h = torch.zeros(B * Tn, ld, dtype=dtype)
    h[:, :D] = rnd((B * Tn, D), dtype, 3)
    gr, hgr = torch.zeros(B * Tn, ld, dtype=dtype), torch.zeros(B * Tn, ld, dtype=dtype)
    gg, hgg = torch.full_like(gr, float("nan")).cuda(), torch.full_like(gr, float("nan")).cuda()
    Hk.pred_gate_fwd(tmr, facts, WpT, bias, h, gr, hgr, B, Tn, t0, F, D, NP, lag)
    K.pred_gate_fwd(tmg, cu(facts), cu(WpT), cu(bias), cu(h), gg, hgg, B, Tn, t0, F, D, NP, lag)
    assert err(gg, gr) < TOL[dtype] and err(hgg, hgr) < TOL[dtype]
    # pointer scores into a shared score buffer
    ctx = torch.zeros(B * F, ld, dtype=dtype)
    ctx[:, :D] = rnd((B * F, D), dtype, 4)
    w, b1 = rnd((D,), torch.float32, 5), rnd((1,), torch.float32, 6)
    Wd = V + E + F
    sr, sg = torch.zeros(B * Tn, Wd), torch.zeros(B * Tn, Wd).cuda()
    Hk.pointer_fwd(h, ctx, w, b1, ftr, sr, B, Tn, t0, F, D, V + E, lag)
    K.pointer_fwd(cu(h), cu(ctx), cu(w), cu(b1), ftg, sg, B, Tn, t0, F, D, V + E, lag)
    assert err(sg, sr) < TOL[dtype]
    if Tn == 1:
        return
    # backward kernels (teacher-forced only)
    ldd = (Wd + 7) // 8 * 8
    dS = torch.zeros(B * T, ldd, dtype=dtype)
    dS[:, :Wd] = rnd((B * T, Wd), dtype, 7)
    n_flat = D + 1 + D * NP + 9
    for first in (ftr, None):
        dCr, dCg = rnd((B * F, ld), torch.float32, 8), None
        dCg = dCr.clone().cuda()
        dHr = rnd((B * T, ld), dtype, 9)
        dHg = dHr.clone().cuda()
        Gr, Gg = torch.zeros(n_flat), torch.zeros(n_flat).cuda()
        Hk.pointer_bwd(dS, h, ctx, w, first, dCr, dHr, Gr, 1, 0, B, T, F, D, V + E, 0)
        K.pointer_bwd(cu(dS), cu(h), cu(ctx), cu(w), cu(first), dCg, dHg, Gg, 1, 0, B, T, F, D, V + E, 0)
        assert err(dCg[:, :D], dCr[:, :D]) < TOL[dtype] and err(dHg[:, :D], dHr[:, :D]) < TOL[dtype] * 2
        assert err(Gg, Gr) < (1e-4 if dtype == torch.float32 else 1e-3)
    dhg = torch.zeros(B * T, ld, dtype=dtype)
    dhg[:, :D] = rnd((B * T, D), dtype, 10)
    outs_r = [torch.zeros(B * T, ld, dtype=dtype) for _ in range(2)]
    outs_g = [torch.zeros(B * T, ld, dtype=dtype).cuda() for _ in range(2)]
    Hk.gate_mul_bwd(dhg, h, gr, *outs_r)
    K.gate_mul_bwd(cu(dhg), cu(h), cu(gr), *outs_g)
    assert err(outs_g[0], outs_r[0]) < TOL[dtype] and err(outs_g[1], outs_r[1]) < TOL[dtype]
    Gr, Gg = torch.zeros(n_flat), torch.zeros(n_flat).cuda()
    Hk.pred_gate_bwd(outs_r[0], tmr, facts, Gr, 5, B, T, F, D, NP, 0)
    K.pred_gate_bwd(cu(outs_r[0]), tmg, cu(facts), Gg, 5, B, T, F, D, NP, 0)
    assert err(Gg, Gr) < 1e-4


@pytest.mark.parametrize("S,col0,masked", [(51, 87, True), (301, 64, False), (130, 64, True)])
def test_pointer_tensor_core_path(K, Hk, S, col0, masked):
    """bf16 pointer heads at teacher-forced sizes (T >= 16): the mma.sync kernels of pointer_mma.cu, including a score slice
    that starts at an odd column (col0 = V + E in the reference layout) and slot counts that are not multiples of 16."""
    dtype = torch.bfloat16
    B, T, D, ld = 3, 37, 300, 320
    Wd = col0 + S + 3
    h = torch.zeros(B * T, ld, dtype=dtype)
    h[:, :D] = rnd((B * T, D), dtype, 3)
    ctx = torch.zeros(B * S, ld, dtype=dtype)
    ctx[:, :D] = rnd((B * S, D), dtype, 4)
    w, b1 = rnd((D,), torch.float32, 5), rnd((1,), torch.float32, 6)
    first = torch.randint(0, T + 6, (B * S,), generator=g(11)).int() if masked else None
    sr, sg = torch.zeros(B * T, Wd), torch.zeros(B * T, Wd).cuda()
    Hk.pointer_fwd(h, ctx, w, b1, first, sr, B, T, 0, S, D, col0, 0)
    K.pointer_fwd(cu(h), cu(ctx), cu(w), cu(b1), cu(first), sg, B, T, 0, S, D, col0, 0)
    assert err(sg, sr) < TOL[dtype]
    assert torch.equal(sg.cpu()[:, :col0], sr[:, :col0]) and torch.equal(sg.cpu()[:, col0 + S:], sr[:, col0 + S:])  # nothing else touched
    ldd = (Wd + 7) // 8 * 8
    dS = torch.zeros(B * T, ldd, dtype=dtype)
    dS[:, :Wd] = rnd((B * T, Wd), dtype, 7)
    dCr = rnd((B * S, ld), torch.float32, 8)
    dCg = dCr.clone().cuda()
    dHr = rnd((B * T, ld), dtype, 9)
    dHg = dHr.clone().cuda()
    Gr, Gg = torch.zeros(D + 9), torch.zeros(D + 9).cuda()
    Hk.pointer_bwd(dS, h, ctx, w, first, dCr, dHr, Gr, 1, 0, B, T, S, D, col0, 0)
    K.pointer_bwd(cu(dS), cu(h), cu(ctx), cu(w), cu(first), dCg, dHg, Gg, 1, 0, B, T, S, D, col0, 0)
    assert err(dCg[:, :D], dCr[:, :D]) < TOL[dtype] and err(dHg[:, :D], dHr[:, :D]) < TOL[dtype] * 2
    assert torch.equal(dCg.cpu()[:, D:], dCr[:, D:])  # pad columns untouched
    assert err(Gg, Gr) < 2e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Wd", [10352, 10301, 1038, 2])
def test_ce_row_kernels(K, Hk, dtype, Wd):
    """Packed cross-entropy + its gradient at the reference's score widths: even widths (K / N: 10 352) run the register-resident
    kernel, odd ones (G: 10 301) the scalar one; targets in the first and the last column, pad targets, rows behind the decode length."""
    B, T = 4, 7
    scores = rnd((B * T, Wd), torch.float32, 1, 3.0)
    caps = torch.randint(1, Wd, (B, T), generator=g(2))
    caps[0, 1] = Wd - 1  # last column
    caps[0, 2] = 1
    caps[1, 3:] = 0  # pad targets inside the decode length are ignored
    dl = torch.tensor([6, 6, 4, 0], dtype=torch.int32)
    ldd = (Wd + 7) // 8 * 8
    accr, accg = torch.zeros(2), torch.zeros(2).cuda()
    dr, dg = torch.zeros(B * T, ldd, dtype=dtype), torch.full((B * T, ldd), float("nan"), dtype=dtype).cuda()
    Hk.ce(scores, caps, dl, accr, dr, B, T, Wd, 0)
    K.ce(cu(scores), cu(caps), cu(dl), accg, dg, B, T, Wd, 0)
    assert float(accg[1]) == float(accr[1]) and abs(float(accg[0]) - float(accr[0])) < 1e-4 * float(accr[0])
    assert not torch.isnan(dg.float()).any()
    assert err(dg, dr) < TOL[dtype]
    # the gradient rows sum to zero (softmax - onehot) on valid rows and are exactly zero elsewhere
    rs = dg.float().sum(1).cpu()
    valid = (dr.float().abs().sum(1) > 0)
    assert float(rs[valid].abs().max()) < (1e-4 if dtype == torch.float32 else 5e-2)
    assert float(dg.float().cpu()[~valid].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
def test_ce_adam_misc(K, Hk, dtype):
    B, T, Wd = 5, 9, 1037
    scores = rnd((B * T, Wd), torch.float32, 1, 3.0)
    caps = torch.randint(0, Wd, (B, T), generator=g(2))
    caps[:, 0] = 3
    caps[1, 4:] = 0  # pad targets inside the decode length are ignored
    dl = torch.tensor([8, 8, 5, 3, 0], dtype=torch.int32)
    ldd = (Wd + 7) // 8 * 8
    accr, accg = torch.zeros(2), torch.zeros(2).cuda()
    dr, dg = torch.zeros(B * T, ldd, dtype=dtype), torch.full((B * T, ldd), float("nan"), dtype=dtype).cuda()
    Hk.ce(scores, caps, dl, accr, dr, B, T, Wd, 0)
    K.ce(cu(scores), cu(caps), cu(dl), accg, dg, B, T, Wd, 0)
    assert float(accg[1]) == float(accr[1]) and abs(float(accg[0]) - float(accr[0])) < 1e-3 * float(accr[0])
    assert err(dg, dr) < TOL[dtype]
    # adam + packing
    n = 5000
    p0, gr_, m0, v0 = rnd((n,), torch.float32, 3), rnd((n,), torch.float32, 4, 10.0), rnd((n,), torch.float32, 5, 0.1), rnd((n,), torch.float32, 6).abs()
    dstA = torch.randperm(n, generator=g(7)).int()
    dstA[::5] = -1
    dstB = (n + torch.randperm(n, generator=g(8))).int()
    dstC = torch.randperm(n, generator=g(9)).int()
    dstC[::3] = -1
    cnt = torch.tensor([7.0])
    args = (4e-4, 0.9, 0.999, 1e-8, 1 - 0.9 ** 3, 1 - 0.999 ** 3, 5.0)
    pr, mr_, vr = p0.clone(), m0.clone(), v0.clone()
    pTr, pFr = torch.zeros(2 * n, dtype=dtype), torch.zeros(n)
    Hk.adam_step(pr, gr_, mr_, vr, *args, cnt, 2.0, dstA, dstB, dstC, pTr, pFr, True)
    pg, mg_, vg = p0.clone().cuda(), m0.clone().cuda(), v0.clone().cuda()
    pTg, pFg = torch.zeros(2 * n, dtype=dtype).cuda(), torch.zeros(n).cuda()
    K.adam_step(pg, cu(gr_), mg_, vg, *args, cu(cnt), 2.0, cu(dstA), cu(dstB), cu(dstC), pTg, pFg, True)
    assert err(pg, pr) < 1e-6 and err(mg_, mr_) < 1e-6 and err(vg, vr) < 1e-6
    assert err(pTg, pTr) < (1e-6 if dtype == torch.float32 else 1e-2) and err(pFg, pFr) < 1e-6
    # cast2d / accum / colsum
    src = rnd((37, 50), torch.float32, 10)
    dr2, dg2 = torch.zeros(37, 56, dtype=dtype), torch.full((37, 56), float("nan"), dtype=dtype).cuda()
    Hk.cast2d(src[:, :45], dr2, 45)
    K.cast2d(cu(src)[:, :45], dg2, 45)
    assert torch.equal(dg2.cpu(), dr2)
    a = rnd((1000,), dtype, 11)
    br, bg = rnd((1000,), torch.float32, 12), None
    bg = br.clone().cuda()
    Hk.accum_f32(a, br)
    K.accum_f32(cu(a), bg)
    assert err(bg, br) < 1e-6
    x = rnd((777, 320), dtype, 13)
    cr, cg = torch.zeros(300), torch.zeros(300).cuda()
    Hk.colsum(x, cr, 300)
    K.colsum(cu(x), cg, 300)
    assert err(cg, cr) < 1e-4


def test_greedy_select_state_machine(K, Hk):
    """Scripted score sequences that trigger <end>, and the 1/2/3-token repetition clean-up (G/models.py:418-435)."""
    B, Wd, Tmax, V, E = 6, 40, 12, 30, 6
    scripts = [[5, 5, 5, 5, 7, 29], [1, 2, 1, 2, 1, 2, 9], [1, 2, 3, 1, 2, 3, 1, 2, 3, 4], [31, 37, 31, 37, 3, 29],
               [8, 8, 9, 9, 9, 8, 8, 29], [29]]
    state = {}
    for name, dev in (("r", "cpu"), ("g", DEV)):
        state[name] = dict(output=torch.zeros(B, Tmax, dtype=torch.int64, device=dev), second=torch.zeros(B, Tmax, dtype=torch.int32, device=dev),
                           captions=torch.full((B, Tmax), 28, dtype=torch.int64, device=dev), masks=torch.zeros(B, Tmax, dtype=torch.int64, device=dev),
                           done=torch.zeros(B, dtype=torch.int32, device=dev), margins=torch.zeros(B, Tmax, device=dev))
    for step in range(Tmax):
        sc = rnd((B, Wd), torch.float32, 100 + step)
        for b, s in enumerate(scripts):
            tok = s[step] if step < len(s) else 29
            sc[b, tok] = 50.0
            sc[b, (tok + 3) % Wd] = 40.0
        for name, kern in (("r", Hk), ("g", K)):
            st = state[name]
            kern.greedy_select(sc.to(st["output"].device), Wd, st["output"], st["second"], st["captions"], st["masks"], st["done"], st["margins"], B,
                               step, Tmax, V, E, True, 29)
    for k in ("output", "captions", "masks", "done"):
        assert torch.equal(state["g"][k].cpu(), state["r"][k]), k
    assert int(state["r"]["done"].sum()) == B


# ---- beam-search extension kernels ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("group", [1, 3, 5])
def test_mha_decode_beam(K, Hk, dtype, group):
    NI, H, dh, Tmax, klen, M = 4, 10, 30, 9, 6, 45
    R = NI * group
    ldc = 3 * H * 32
    # self-attention: position-major cache + ancestor-slot table
    cache = headify(rnd((Tmax * R, ldc), torch.float32, 1), 3 * H, dh).to(dtype)
    anc = torch.randint(0, group, (R, Tmax), generator=g(2), dtype=torch.int32)
    Q = cache[(klen - 1) * R : klen * R, : H * 32]
    Or, Og = torch.zeros(R, H * 32, dtype=dtype), torch.zeros(R, H * 32, dtype=dtype).cuda()
    Hk.mha_decode_beam(Q, cache[:, H * 32 : 2 * H * 32], cache[:, 2 * H * 32 :], Or, R, group, H, dh, klen, anc=anc, kpos_stride=R * ldc,
                       vpos_stride=R * ldc)
    cg = cu(cache)
    K.mha_decode_beam(cg[(klen - 1) * R : klen * R, : H * 32], cg[:, H * 32 : 2 * H * 32], cg[:, 2 * H * 32 :], Og, R, group, H, dh, klen,
                      anc=cu(anc), kpos_stride=R * ldc, vpos_stride=R * ldc)
    assert err(Og, Or) < TOL[dtype]
    # cross-attention: the beams of an image share its keys / values
    ldkv = 2 * H * 32
    kv = headify(rnd((NI * M, ldkv), torch.float32, 3), 2 * H, dh).to(dtype)
    q = headify(rnd((R, H * 32), torch.float32, 4), H, dh).to(dtype)
    Or.zero_()
    Og.zero_()
    Hk.mha_decode_beam(q, kv[:, : H * 32], kv[:, H * 32 :], Or, R, group, H, dh, M, kimg_stride=M * ldkv, vimg_stride=M * ldkv)
    kg = cu(kv)
    K.mha_decode_beam(cu(q), kg[:, : H * 32], kg[:, H * 32 :], Og, R, group, H, dh, M, kimg_stride=M * ldkv, vimg_stride=M * ldkv)
    assert err(Og, Or) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_grouped_context_kernels(K, Hk, dtype):
    """caption embed / indicators / gate / pointer heads with `group` rows per image context (beam search)."""
    NI, G, E, F, V, D, ld, NP, T = 3, 5, 23, 17, 61, 300, 320, 3000, 9
    R = NI * G
    cfg, batch = make_context(1, NI, E, F, V, seed=5)
    facts = batch["facts"]
    caps = torch.randint(1, V + E + F, (R, T), generator=g(1))
    caps[:, 0] = V - 2
    masks = torch.where(caps >= V + E, 2, torch.where(caps >= V, 1, 0))
    ent = torch.zeros(NI * E, ld, dtype=dtype)
    ent[:, :D] = rnd((NI * E, D), dtype, 2)
    fct = torch.zeros(NI * F, ld, dtype=dtype)
    fct[:, :D] = rnd((NI * F, D), dtype, 3)
    wemb = torch.zeros(V, ld, dtype=dtype)
    wemb[:, :D] = rnd((V, D), dtype, 4)
    pe = rnd((T, D), torch.float32, 5)
    t0 = 4
    xr, xg = torch.zeros(R, ld, dtype=dtype), torch.full((R, ld), float("nan"), dtype=dtype).cuda()
    Hk.caption_embed_fwd(caps, masks, wemb, ent, fct, pe, xr, R, T, t0, 1, V, E, F, D, 0, math.sqrt(D), group=G)
    K.caption_embed_fwd(cu(caps), cu(masks), cu(wemb), cu(ent), cu(fct), cu(pe), xg, R, T, t0, 1, V, E, F, D, 0, math.sqrt(D), group=G)
    assert err(xg, xr) < TOL[dtype]
    ftr, tmr = torch.zeros(R * F, dtype=torch.int32), torch.zeros(R * F, dtype=torch.int32)
    ftg, tmg = ftr.clone().cuda(), tmr.clone().cuda()
    Hk.fact_first_mention(caps, facts, ftr, tmr, R, T, F, V, E, group=G)
    K.fact_first_mention(cu(caps), cu(facts), ftg, tmg, R, T, F, V, E, group=G)
    assert torch.equal(ftg.cpu(), ftr) and torch.equal(tmg.cpu(), tmr)
    assert int((ftr < FIRST_NONE).sum()) > 0
    WpT = torch.zeros(NP, ld)
    WpT[:, :D] = rnd((NP, D), torch.float32, 6, 0.3)
    bias = rnd((D,), torch.float32, 7)
    h = torch.zeros(R, ld, dtype=dtype)
    h[:, :D] = rnd((R, D), dtype, 8)
    gr, hgr = torch.zeros(R, ld, dtype=dtype), torch.zeros(R, ld, dtype=dtype)
    gg, hgg = torch.full_like(gr, float("nan")).cuda(), torch.full_like(gr, float("nan")).cuda()
    Hk.pred_gate_fwd(tmr, facts, WpT, bias, h, gr, hgr, R, 1, t0, F, D, NP, 1, group=G)
    K.pred_gate_fwd(tmg, cu(facts), cu(WpT), cu(bias), cu(h), gg, hgg, R, 1, t0, F, D, NP, 1, group=G)
    assert err(gg, gr) < TOL[dtype] and err(hgg, hgr) < TOL[dtype]
    w, b1 = rnd((D,), torch.float32, 9), rnd((1,), torch.float32, 10)
    Wd = V + E + F
    sr, sg = torch.zeros(R, Wd), torch.zeros(R, Wd).cuda()
    Hk.pointer_fwd(h, ent, w, b1, None, sr, R, 1, t0, E, D, V, 1, group=G)
    Hk.pointer_fwd(h, fct, w, b1, ftr, sr, R, 1, t0, F, D, V + E, 1, group=G)
    K.pointer_fwd(cu(h), cu(ent), cu(w), cu(b1), None, sg, R, 1, t0, E, D, V, 1, group=G)
    K.pointer_fwd(cu(h), cu(fct), cu(w), cu(b1), ftg, sg, R, 1, t0, F, D, V + E, 1, group=G)
    assert err(sg, sr) < TOL[dtype]


@pytest.mark.parametrize("G", [1, 3, 5])
def test_beam_select_state_machine(K, Hk, G):
    """Random score tables with <end> pushed into the top candidates at scripted steps: live-beam count, histories, ancestor
    tables, cumulative scores and the best completed caption must match the host restatement exactly (fp32 scores: 1e-5)."""
    NI, Wd, Tmax, V, E, end, pad = 7, 47, 8, 30, 6, 29, 0
    R = NI * G
    state = {}
    for name, dev in (("r", "cpu"), ("g", DEV)):
        own = (torch.arange(R, dtype=torch.int32) % G).unsqueeze(1).expand(R, Tmax).contiguous()
        state[name] = dict(tok=[torch.full((R, Tmax), 28, dtype=torch.int64, device=dev) for _ in range(2)],
                           msk=[torch.zeros(R, Tmax, dtype=torch.int64, device=dev) for _ in range(2)],
                           anc=[own.clone().to(dev) for _ in range(2)], cum=torch.zeros(R, device=dev),
                           ksel=torch.full((NI,), G, dtype=torch.int32, device=dev), best=torch.full((NI,), float("-inf"), device=dev),
                           result=torch.full((NI, Tmax), pad, dtype=torch.int64, device=dev))
    for step in range(Tmax):
        sc = rnd((R, Wd), torch.float32, 200 + step, 2.0)
        for img in range(NI):
            if img == 6:
                sc[img * G : (img + 1) * G, end] = -50.0  # never completes: best live beam at the last step
            elif (step + img) % 3 == 1:
                sc[img * G + (step % G), end] = 6.0  # one beam of the image very likely ends here
        cur, nxt = step & 1, (step + 1) & 1
        for name, kern in (("r", Hk), ("g", K)):
            st = state[name]
            kern.beam_select(sc.to(st["cum"].device), Wd, st["cum"], st["ksel"], st["tok"][cur], st["msk"][cur], st["tok"][nxt], st["msk"][nxt],
                             st["anc"][cur], st["anc"][nxt], st["best"], st["result"], NI, G, step, Tmax, V, E, True, end, pad)
        kr = state["r"]["ksel"]
        assert torch.equal(state["g"]["ksel"].cpu(), kr), step
        for img in range(NI):  # live rows only (dead slots hold stale data by design)
            rows = slice(img * G, img * G + int(kr[img]))
            upto = min(step + 2, Tmax)
            for key in ("tok", "msk", "anc"):
                assert torch.equal(state["g"][key][nxt][rows, :upto].cpu(), state["r"][key][nxt][rows, :upto]), (key, step, img)
            assert torch.allclose(state["g"]["cum"][rows].cpu(), state["r"]["cum"][rows], atol=1e-5)
    assert torch.equal(state["g"]["result"].cpu(), state["r"]["result"])
    assert torch.allclose(state["g"]["best"].cpu(), state["r"]["best"], atol=1e-5)
    assert int((state["r"]["ksel"] < G).sum()) > 0 and torch.isfinite(state["r"]["best"]).all()


@pytest.mark.parametrize("kind", ["normal", "quantised", "constant", "short"])
def test_beam_row_topk_candidates_at_full_width(K, kind):
    """Stage 1 of beam_select at the reference's score width (10 352 columns, the register-resident kernel): the per-row candidate
    lists (value = cum + log_softmax, column) must be the row's G best in the order (value descending, column ascending).  "quantised"
    and "constant" rows are full of ties: the per-warp threshold lists overflow and the kernel must take its full-rounds path."""
    if DRYRUN:
        pytest.skip("reads the kernel's workspace layout")
    G, NI, Tmax, V, E = 5, 6, 4, 10000, 301
    Wd = 10352 if kind != "short" else 700
    R = NI * G
    sc = rnd((R, Wd), torch.float32, 77, 3.0)
    if kind == "quantised":
        sc = torch.round(sc * 2) / 2
    elif kind == "constant":
        sc = torch.full((R, Wd), 1.25)
        sc[3, 77] = 1.5
    cum = rnd((R,), torch.float32, 78)
    own = (torch.arange(R, dtype=torch.int32) % G).unsqueeze(1).expand(R, Tmax).contiguous()
    tok = [torch.full((R, Tmax), 28, dtype=torch.int64, device=DEV) for _ in range(2)]
    msk = [torch.zeros(R, Tmax, dtype=torch.int64, device=DEV) for _ in range(2)]
    anc = [own.clone().to(DEV) for _ in range(2)]
    ws = torch.zeros(R * G * 2, dtype=torch.float32, device=DEV)
    K.beam_select(cu(sc), Wd, cu(cum).clone(), torch.full((NI,), G, dtype=torch.int32, device=DEV), tok[0], msk[0], tok[1], msk[1], anc[0], anc[1],
                  torch.full((NI,), float("-inf"), device=DEV), torch.zeros(NI, Tmax, dtype=torch.int64, device=DEV), NI, G, 1, Tmax, V, E, True,
                  V - 1, 0, workspace=ws)
    cand_v = ws[: R * G].cpu().view(R, G)
    cand_i = ws[R * G :].cpu().view(torch.int32).view(R, G)
    lp = torch.log_softmax(sc.double(), dim=1)
    order = torch.sort(-sc, dim=1, stable=True).indices[:, :G]  # value descending, ties by ascending column
    want_i = order + (torch.arange(R) % G).unsqueeze(1) * Wd
    want_v = (cum.double().unsqueeze(1) + lp.gather(1, order)).float()
    assert torch.equal(cand_i.long(), want_i), kind
    assert torch.allclose(cand_v, want_v, atol=2e-5, rtol=1e-6)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("channels_last", [False, True])
def test_image_prep_matches_the_reference_host_pipeline(K, dtype, channels_last):
    """fp16 HDF5 pixels -> normalised encoder input (G/datasets.py:44 + G/train.py:139-147): fp32 output bit-identical."""
    from oracle import decoder_oracle as orc

    raw = (torch.rand(5, 3, 24, 40, generator=g(3)) * 255).half()
    raw[0, 0, 0, :4] = torch.tensor([0.0, 255.0, 1.0, 254.5]).half()
    ref = orc.prepare_images(raw)
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    out = torch.empty(raw.shape, dtype=dtype, device=DEV, memory_format=fmt)
    from ickb200.data import IMAGENET_MEAN, IMAGENET_STD

    K.image_prep(cu(raw), out, IMAGENET_MEAN, IMAGENET_STD, channels_last=channels_last)
    if dtype == torch.float32:
        assert torch.equal(out.cpu(), ref)
    else:
        assert torch.equal(out.cpu(), ref.to(torch.bfloat16))  # one rounding of the same fp32 value
    if not DRYRUN:
        from ickb200.data import prepare_images

        got = prepare_images(cu(raw), dtype, channels_last)
        assert got.is_contiguous(memory_format=fmt) and torch.equal(got.cpu(), out.cpu())


@pytest.mark.parametrize("klen", [548, 497, 64, 71])
def test_mha_decode_tma_streaming_path(K, Hk, klen):
    """Per-step cross-attention over contiguous K|V rows (the per-layer memory buffers of the decode loops): bf16, klen >= 64
    takes the TMA-staged kernel; against the host restatement, and against the CUDA-core kernel on a non-contiguous copy."""
    B, H, dh = 9, 10, 30
    ldkv = 2 * H * 32
    dtype = torch.bfloat16
    kv = headify(rnd((B * klen, ldkv), torch.float32, 11), 2 * H, dh).to(dtype)
    q = headify(rnd((B, H * 32), torch.float32, 12), H, dh).to(dtype)
    Or = torch.zeros(B, H * 32, dtype=dtype)
    Hk.mha_decode(q, kv[:, : H * 32], kv[:, H * 32 :], Or, B, H, dh, klen * ldkv, klen * ldkv, klen)
    kg, qg = cu(kv), cu(q)
    guard = torch.full((B + 2, H * 32 + 16), 7.0, dtype=dtype, device=DEV)  # canary rows / columns around the output
    Og = guard[1 : B + 1, : H * 32]
    K.mha_decode(qg, kg[:, : H * 32], kg[:, H * 32 :], Og, B, H, dh, klen * ldkv, klen * ldkv, klen)
    assert err(Og, Or) < TOL[dtype]
    assert bool((guard[0] == 7).all()) and bool((guard[-1] == 7).all()) and bool((guard[:, H * 32 :] == 7).all())
    wide = torch.zeros(B * klen, ldkv + 64, dtype=dtype, device=DEV)  # same rows, padded stride -> the CUDA-core kernel
    wide[:, :ldkv] = kg
    O2 = torch.zeros_like(Og)
    K.mha_decode(qg, wide[:, : H * 32], wide[:, H * 32 : ldkv], O2, B, H, dh, klen * (ldkv + 64), klen * (ldkv + 64), klen)
    assert err(Og, O2) < 1e-2


@pytest.mark.parametrize("G", [5, 1, 3, 8])
@pytest.mark.parametrize("klen,NI", [(548, 37), (598, 9), (65, 330), (80, 5), (257, 64)])
def test_mha_decode_beam_tma_tensor_core_path(K, Hk, G, klen, NI):
    """Beam-search step cross-attention: the G beams of an image against its contiguous memory K|V rows run on the TMA-streamed
    tensor-core kernel (mha_decode_tma_mma_kernel: bf16, more than 64 cached positions).  Against the host restatement, per row (a
    wrong stage tenant or a broken fragment layout corrupts single rows / heads), with canaries around the output, and against the
    CUDA-core kernel on a copy with a padded stride."""
    H, dh = 10, 30
    R = NI * G
    ldkv = 2 * H * 32
    dtype = torch.bfloat16
    kv = headify(rnd((NI * klen, ldkv), torch.float32, 21, 1.5), 2 * H, dh).to(dtype)
    q = headify(rnd((R, H * 32), torch.float32, 22, 1.5), H, dh).to(dtype)
    Or = torch.zeros(R, H * 32, dtype=dtype)
    Hk.mha_decode_beam(q, kv[:, : H * 32], kv[:, H * 32 :], Or, R, G, H, dh, klen, kimg_stride=klen * ldkv, vimg_stride=klen * ldkv)
    kg, qg = cu(kv), cu(q)
    guard = torch.full((R + 2, H * 32 + 16), 7.0, dtype=dtype, device=DEV)
    Og = guard[1 : R + 1, : H * 32]
    K.mha_decode_beam(qg, kg[:, : H * 32], kg[:, H * 32 :], Og, R, G, H, dh, klen, kimg_stride=klen * ldkv, vimg_stride=klen * ldkv)
    assert not torch.isnan(Og.float()).any()
    ai, bi = Og.float().cpu().view(R * H, 32), Or.float().view(R * H, 32)
    rel = (ai - bi).abs().amax(1) / bi.abs().amax(1).clamp_min(1e-6)
    assert float(rel.max()) < TOL[dtype], (int(rel.argmax()), float(rel.max()))
    assert float(Og.float().view(R, H, 32)[:, :, dh:].abs().max()) == 0.0  # pad lanes of every head stay zero
    assert bool((guard[0] == 7).all()) and bool((guard[-1] == 7).all()) and bool((guard[:, H * 32 :] == 7).all())
    wide = torch.zeros(NI * klen, ldkv + 64, dtype=dtype, device=DEV)  # same rows, padded stride -> the CUDA-core kernel
    wide[:, :ldkv] = kg
    O2 = torch.zeros(R, H * 32, dtype=dtype, device=DEV)
    K.mha_decode_beam(qg, wide[:, : H * 32], wide[:, H * 32 : ldkv], O2, R, G, H, dh, klen, kimg_stride=klen * (ldkv + 64),
                      vimg_stride=klen * (ldkv + 64))
    assert err(Og, O2) < 1e-2


@pytest.mark.parametrize("rows", [625, 37, 16])
@pytest.mark.parametrize("ffn,proj_n", [(False, 320), (True, 960), (True, 0)])
def test_decode_chain(K, Hk, rows, ffn, proj_n):
    """Fused row-wise tail of a decoder layer (out-proj + LN [+ FFN + LN] [+ next projection]) against the host restatement."""
    D, DP, FFP = 300, 320, 512
    dtype = torch.bfloat16

    def w(n, k, kreal, seed):
        m = torch.zeros(n, k)
        m[:, :kreal] = rnd((n, kreal), torch.float32, seed, 0.08)
        return m.to(dtype)

    def rows_of(seed, ld=DP):
        x = torch.zeros(rows, ld)
        x[:, :D] = rnd((rows, D), torch.float32, seed)
        return x.to(dtype)

    a = headify(rnd((rows, DP), torch.float32, 1), 10, 30).to(dtype)
    x = rows_of(2)
    Wo, bo = w(DP, DP, DP, 3), torch.cat([rnd((D,), torch.float32, 4, 0.1), torch.zeros(DP - D)])
    Wo[D:] = 0
    g1, be1 = 1 + rnd((D,), torch.float32, 5, 0.1), rnd((D,), torch.float32, 6, 0.1)
    f = None
    if ffn:
        W1, b1 = w(FFP, DP, D, 7), rnd((FFP,), torch.float32, 8, 0.1)
        W2, b2 = w(DP, FFP, FFP, 9), torch.cat([rnd((D,), torch.float32, 10, 0.1), torch.zeros(DP - D)])
        W2[D:] = 0
        f = (W1, b1, W2, b2, 1 + rnd((D,), torch.float32, 11, 0.1), rnd((D,), torch.float32, 12, 0.1))
    yr, yg = torch.zeros(rows, DP, dtype=dtype), torch.full((rows, DP), float("nan"), dtype=dtype).to(DEV)
    pr = pg = None
    pj_r = pj_g = None
    if proj_n:
        Wn, bn = w(proj_n, DP, D, 13), rnd((proj_n,), torch.float32, 14, 0.1)
        big_r, big_g = torch.zeros(rows, 3 * proj_n, dtype=dtype), torch.zeros(rows, 3 * proj_n, dtype=dtype).to(DEV)
        pr, pg = big_r[:, proj_n : 2 * proj_n], big_g[:, proj_n : 2 * proj_n]  # strided rows, like a cache slice
        pj_r, pj_g = (Wn, bn, pr), (cu(Wn), cu(bn), pg)
    Hk.decode_chain(a, x, Wo, bo, g1, be1, yr, D, ffn=f, proj=pj_r)
    K.decode_chain(cu(a), cu(x), cu(Wo), cu(bo), cu(g1), cu(be1), yg, D, ffn=tuple(cu(t) for t in f) if f else None, proj=pj_g)
    assert torch.isfinite(yg.float()).all()
    assert err(yg, yr) < TOL[dtype] and float(yg[:, D:].float().abs().max()) == 0.0
    if proj_n:
        assert err(pg, pr) < TOL[dtype]
        assert float(big_g[:, :proj_n].float().abs().max()) == 0.0 and float(big_g[:, 2 * proj_n :].float().abs().max()) == 0.0  # canaries
