"""
GPU: parity at the BASELINE shapes (G T=32 E=301; K T=102 E=301 F=51; N T=52 E=101 F=301; V=10000; batch 8).

The CUDA path (drop-in module -> C ABI -> sm_100a kernels) against
  * tests/golden/golden_base_{G,K,N}.npz - outputs of the UNMODIFIED reference modules at these shapes (make_golden.py --baseline),
  * the CPU oracle run live on the box's host cores on the same inputs (full score tensors, every gradient).
Tolerances are north_star's: fp32 scores 1e-4 relative, token-identical greedy decode; bf16 logits 2e-2 relative and the
loss within 1e-3 ABSOLUTE (weights at the reference's init scale, so the absolute bound means what it says).
These shapes reach the code paths the toy shapes do not: dQ from the stored dS^T for the 301-token self-attention, 256-wide
vocabulary tiles, more than one k-tile in the FFN, the multi-chunk key loop of the 548 / 598-token cross-attention.
"""
import numpy as np
import pytest
import torch

from helpers import batch_args, build_module, load_base_golden, nmax_err, oracle_drop_fn, oracle_params, spec_for
from ickb200 import synthetic as syn
from oracle import decoder_oracle as orc
from test_oracle_golden import check_grads_against_base_golden, check_scores_against_base_golden

pytestmark = pytest.mark.gpu


def to_dev(batch):
    out = dict(batch)
    for k in ("captions", "encoder_out", "caption_masks", "caption_lengths", "facts"):
        if k in out:
            out[k] = out[k].cuda()
    return out


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_fp32_forward_backward_vs_reference_golden(variant):
    cfg = syn.BASE_PARITY_CONFIGS[variant]
    g = load_base_golden(variant)
    dec = build_module(cfg, "cuda", torch.float32, profile="reference").eval()
    batch = to_dev(syn.make_batch(cfg, seed=int(g["seed"])))
    batch["encoder_out"].requires_grad_(True)
    scores, caps, dl = dec(*batch_args(cfg, batch))
    assert np.array_equal(caps.cpu().numpy(), g["captions_sorted"]) and dl == g["decode_lengths"].tolist()
    v = check_scores_against_base_golden(scores, g, cfg.V, 1e-4)
    assert float((v["s_argmax"] == g["s_argmax"]).mean()) > 0.999
    loss = orc.caption_loss(scores, caps, dl)  # train.py's pack + CrossEntropyLoss(ignore_index=<pad>), torch ops on the device
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4
    loss.backward()
    ge = batch["encoder_out"].grad.cpu()
    assert nmax_err(ge[:, ::29, ::11], g["grad_encoder_out_rows"]) < 2e-3
    assert abs(float(ge.double().norm()) - float(g["gnorm_encoder_out"])) <= 2e-3 * float(g["gnorm_encoder_out"])
    check_grads_against_base_golden(((k, p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in dec.named_parameters()), g, 2e-3)


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_fp32_predict_token_identical_to_reference(variant):
    cfg = syn.BASE_PARITY_CONFIGS[variant]
    g = load_base_golden(variant)
    dec = build_module(cfg, "cuda", torch.float32, profile="reference").eval()
    pb = syn.make_batch(cfg, seed=int(g["predict_seed"]))
    T = int(g["predict_max_len"])
    n = g["predict_tokens"].shape[0]
    facts = pb["facts"][:n].cuda() if cfg.has_facts else None
    out, margins = dec.predict_batch(pb["encoder_out"][:n].cuda(), T, pb["entities"][:n], facts, return_margins=True)
    got, ref = out.cpu().numpy(), g["predict_tokens"]
    assert float(g["predict_min_margin"].min()) > 1e-4  # the reference's own top-1 / top-2 gaps: no fp32 ties in these captions
    if not np.array_equal(got, ref):
        b, t = np.argwhere(got != ref)[0]
        raise AssertionError(f"first divergence image {b} step {t}: got {got[b, t]} ref {ref[b, t]}, margin {float(margins[b, t]):.3e}")
    # the graph-captured loop and the reference's own batch-1 call signature
    assert np.array_equal(dec.predict_batch(pb["encoder_out"][:n].cuda(), T, pb["entities"][:n], facts).cpu().numpy(), ref)
    one = dec.predict(pb["encoder_out"][:1].cuda(), T, pb["entities"][:1], facts[:1] if facts is not None else None)
    assert tuple(one.shape) == (T, 1) and one.reshape(-1).tolist() == ref[0].tolist()


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_bf16_logits_and_loss_vs_reference_golden(variant):
    """north_star: bf16 logits within 2e-2 relative, loss within 1e-3 absolute."""
    cfg = syn.BASE_PARITY_CONFIGS[variant]
    g = load_base_golden(variant)
    dec = build_module(cfg, "cuda", torch.bfloat16, profile="reference").eval()
    batch = to_dev(syn.make_batch(cfg, seed=int(g["seed"])))
    scores, caps, dl = dec(*batch_args(cfg, batch))
    check_scores_against_base_golden(scores, g, cfg.V, 2e-2)
    eng = dec._engine
    acc, _ = eng.loss(scores.detach(), caps, torch.tensor(dl, dtype=torch.int32, device="cuda"))
    loss_fused = float(acc[0] / acc[1])  # the fused masked-CE kernel
    loss_torch = float(orc.caption_loss(scores.detach(), caps, dl))  # train.py's recipe on the same scores
    assert abs(loss_fused - float(g["loss"])) < 1e-3, (loss_fused, float(g["loss"]))
    assert abs(loss_torch - float(g["loss"])) < 1e-3
    # every gradient against the live fp32 oracle: norms within 3 %, direction (cosine) above 0.995 for the big matrices
    orc.caption_loss(scores, caps, dl).backward()
    p = oracle_params(cfg, requires_grad=True, profile="reference")
    cpu_batch = syn.make_batch(cfg, seed=int(g["seed"]))
    rs, rc, rdl = orc.forward(spec_for(cfg), p, *batch_args(cfg, cpu_batch))
    assert nmax_err(scores.detach().cpu(), rs.detach()) < 2e-2  # the FULL score tensor, not only the fixture's slices
    orc.caption_loss(rs, rc, rdl).backward()
    for k, prm in dec.named_parameters():
        ref = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        rn = float(ref.double().norm())
        if rn < 1e-7:
            continue
        got = prm.grad.float().cpu()
        assert abs(float(got.double().norm()) - rn) <= 3e-2 * rn, k
        if ref.numel() >= 300:
            cos = float((got.double() * ref.double()).sum() / (got.double().norm() * ref.double().norm()))
            assert cos > 0.995, (k, cos)


def _train_mode_grads(cfg, dtype, ps, torch_seed):
    """One train-mode forward + backward of the CUDA module and of the oracle fed the kernels' own dropout masks."""
    torch.manual_seed(torch_seed)  # the dropout seed derives from torch.initial_seed(): pin the mask draw whatever ran before
    dec = build_module(cfg, "cuda", dtype, dropouts=(ps["dec"], ps["enc"], ps["pos"]), profile="reference").train()
    batch_cpu = syn.make_batch(cfg, seed=23)
    scores, caps, dl = dec(*batch_args(cfg, to_dev(batch_cpu)))
    seed = (int(torch.initial_seed()) * 1000003 + dec._step) & 0x7FFFFFFF
    p = oracle_params(cfg, requires_grad=True, profile="reference")
    ref_scores, _, _ = orc.forward(spec_for(cfg), p, *batch_args(cfg, batch_cpu), drop=oracle_drop_fn(seed, ps))
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert nmax_err(scores.detach().cpu(), ref_scores.detach()) < tol
    orc.caption_loss(scores, caps, dl).backward()
    orc.caption_loss(ref_scores, caps.cpu(), dl).backward()
    out = []
    for k, prm in dec.named_parameters():
        ref = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        got = (prm.grad if prm.grad is not None else torch.zeros_like(prm)).float().cpu()
        out.append((k, got, ref))
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_train_mode_dropout_at_baseline_shapes_vs_oracle_with_same_masks(variant, dtype):
    """Train mode at full size: the oracle is given the kernels' own dropout masks (hash of seed, site, element), so the
    forward and every gradient are compared exactly, not statistically."""
    cfg = syn.BASE_PARITY_CONFIGS[variant].with_batch(4)
    ps = dict(dec=0.5, enc=0.5, pos=0.1)  # the reference's defaults (G/models.py:219)
    if dtype == torch.float32:
        # EVERY element of EVERY gradient within 2e-3 of the gradient's magnitude.  One effect is not a kernel error and is
        # handled explicitly: a ReLU unit whose pre-activation is within fp32 rounding of zero (0.2-0.6 million FFN pre-activations
        # per layer and pass, kept values doubled by the p = 0.5 dropout in front) may get ReLU'(0) = 0 from one side and 1 from
        # the other.  That toggles ONE (token, unit) term: a whole row of that layer's linear1.weight moves by a few per cent and
        # everything upstream of the token by ~0.4 % in l2 (tools/grad_err.py, gpurun_out/r02d: mask draw 1234 of K has exactly one
        # such unit in decoder layer 2; draw 7 agrees to 1e-6 everywhere).  So: up to three mask draws, at least one must be
        # clean in the strict sense, and a draw that is not clean must look like a flip (every parameter within 1 % in l2).
        clean = False
        report = []
        for torch_seed in (1234, 7, 99):
            worst = None
            for k, got, ref in _train_mode_grads(cfg, dtype, ps, torch_seed):
                lim = 2e-3 * max(float(ref.abs().max()), 1e-6) + 1e-7
                nbad = int(((got - ref).abs() > lim).sum())
                if nbad:
                    rel = float((got - ref).double().norm() / max(float(ref.double().norm()), 1e-12))
                    assert rel < 1e-2, (torch_seed, k, nbad, rel)  # larger than one ReLU'(0) flip can explain
                    if worst is None or rel > worst[2]:
                        worst = (k, nbad, rel)
            report.append((torch_seed, worst))
            if worst is None:
                clean = True
                break
        assert clean, report
        return
    for k, got, ref in _train_mode_grads(cfg, dtype, ps, 1234):
        # bf16 activations under p = 0.5 dropout (kept values doubled) through three layers: single parameters land 2-6 % off
        # in norm depending on the mask draw (measured on B200); the direction is the sharper test
        rn = float(ref.double().norm())
        if rn > 1e-7:
            assert abs(float(got.double().norm()) - rn) <= 8e-2 * rn, k
            if ref.numel() >= 300:
                cos = float((got.double() * ref.double()).sum() / (got.double().norm() * ref.double().norm()))
                assert cos > 0.99, (k, cos)


def _oracle_recipe(cfg, batch, steps, profile, lr=4e-4):
    """train.py's step on the oracle: loss.backward(); ut.clip_gradient(+-5); Adam.step()  (G/train.py:275-292, G/utils.py:75-85)."""
    p = oracle_params(cfg, requires_grad=True, profile=profile)
    names = [k for k in p if k != "pos_encoder.pe"]
    opt = torch.optim.Adam([p[k] for k in names], lr=lr)
    losses = []
    for _ in range(steps):
        scores, caps, dl = orc.forward(spec_for(cfg), p, *batch_args(cfg, batch))
        loss = orc.caption_loss(scores, caps, dl)
        opt.zero_grad()
        loss.backward()
        for k in names:
            if p[k].grad is not None:
                p[k].grad.data.clamp_(-5.0, 5.0)
        opt.step()
        losses.append(float(loss))
    return p, opt, losses


@pytest.mark.parametrize("case", ["small", "baseline_K"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_gpu_fused_train_step_equals_reference_recipe(case, use_graph):
    """The fused GPU step (forward, masked CE, hand-written backward, clamp, Adam, re-pack; optionally one CUDA graph) against
    train.py's recipe executed with torch on the oracle: two steps, fp32, dropout off."""
    from ickb200.trainer import Trainer

    if case == "small":
        cfg, profile = syn.SMALL_CONFIGS["K"], "test"
        batch = syn.make_batch(cfg, seed=4, equal_lengths=False)
    else:
        cfg, profile = syn.BASE_PARITY_CONFIGS["K"].with_batch(4), "reference"
        batch = syn.make_batch(cfg, seed=24)
    dec = build_module(cfg, "cuda", torch.float32, dropouts=(0.0, 0.0, 0.0), profile=profile).train()
    tr = Trainer(dec, lr=4e-4, grad_clip=5.0, use_graph=use_graph)
    accs = [tr.train_step(*batch_args(cfg, to_dev(batch))).clone() for _ in range(2)]
    p, opt, losses = _oracle_recipe(cfg, batch, 2, profile)
    for a, l in zip(accs, losses):
        assert abs(float(a[0] / a[1]) - l) < 1e-4
    eng = dec._engine
    for k, prm in dec.named_parameters():
        ref = p[k].detach()
        st = opt.state[p[k]]
        m_ref, v_ref = st["exp_avg"], st["exp_avg_sq"]
        d = (prm.detach().cpu() - ref).abs()
        # Adam normalises by sqrt(v): where |g| ~ eps the update direction is fp32 noise, so weights are compared at a fraction of
        # the 2*lr total step and the moments (linear / quadratic in g) tightly
        sig = v_ref.sqrt() > 1e-3 * max(float(v_ref.sqrt().max()), 1e-12)
        assert float(d.max()) < 1e-3, k
        if bool(sig.any()):
            assert float((d[sig] < 2e-5).float().mean()) > 0.98, k
        assert float((eng.param(k, tr.m).cpu() - m_ref).abs().max()) <= 2e-3 * float(m_ref.abs().max()) + 1e-9, k
        assert float((eng.param(k, tr.v).cpu() - v_ref).abs().max()) <= 4e-3 * float(v_ref.abs().max()) + 1e-12, k
    lin = eng.lin["fc_vocab"]
    assert torch.equal(lin.W[:, : cfg.D], dec._get("fc_vocab.weight").detach())


# ---- two GPUs over NCCL against one GPU ---------------------------------------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_nccl_equals_single_gpu(tmp_path):
    """Two NCCL ranks with half of the batch each end with the parameters of one GPU with the whole batch (gradient all-reduce
    before the clamp, normalisation by the global kept-token count): eager, graph-captured and with the region-wise
    overlapped all-reduce.  The ranks run under torchrun in a subprocess with a hard time limit (a hung collective must not eat
    the GPU budget)."""
    import os
    import subprocess
    import sys

    from ickb200.trainer import Trainer

    cfg, profile = syn.SMALL_CONFIGS["K"].with_batch(4), "test"
    full = syn.make_batch(cfg, seed=4, equal_lengths=False)
    dec1 = build_module(cfg, "cuda:0", torch.float32, dropouts=(0.0, 0.0, 0.0), profile=profile).train()
    tr1 = Trainer(dec1, lr=4e-4, grad_clip=5.0)
    accs1 = [tr1.train_step(*batch_args(cfg, full)).cpu().tolist() for _ in range(2)]
    out_path = tmp_path / "nccl_out.pt"
    here = os.path.dirname(os.path.abspath(__file__))
    port = 33500 + (os.getpid() % 2000)
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(here, "nccl_equiv_worker.py"), str(out_path)],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-3000:]
    out = torch.load(out_path, weights_only=False)
    for mode, (params2, accs2) in out.items():
        for a1, a2 in zip(accs1, accs2):
            assert a1[1] == a2[1] and abs(a1[0] - a2[0]) < 1e-3 * abs(a2[0]), mode
        n_close = n_all = 0
        for k, prm in dec1.named_parameters():
            d = (prm.detach().cpu() - params2[k]).abs()
            assert float(d.max()) < 1e-3, (k, mode)
            n_close += int((d < 2e-5).sum())
            n_all += d.numel()
        assert n_close > 0.999 * n_all, mode
