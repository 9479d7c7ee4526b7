"""
CPU: the host-side orchestration (engine forward + hand-written backward chain, packing plan and gradient index maps,
module sort/autograd/flat-parameter plumbing, greedy-decode loop) run over tests/hostsim.py and compared with the
oracle and the golden vectors.  No CUDA kernel runs here; the `-m gpu` tests repeat these checks on the real kernels.
"""
import numpy as np
import pytest
import torch

from helpers import batch_args, build_module, load_golden, module_cls, nmax_err, oracle_drop_fn, oracle_params, spec_for
from hostsim import HostKernels
from ickb200 import models as M, synthetic as syn
from oracle import decoder_oracle as orc


@pytest.fixture(autouse=True)
def host_kernels():
    M.DecoderTransformer._test_kernel_factory = HostKernels
    yield
    M.DecoderTransformer._test_kernel_factory = None


def test_product_refuses_cpu_without_seam():
    M.DecoderTransformer._test_kernel_factory = None
    cfg = syn.SMALL_CONFIGS["G"]
    dec = build_module(cfg, "cpu")
    batch = syn.make_batch(cfg, seed=1)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        dec(*batch_args(cfg, batch))


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_state_dict_keys_match_reference(variant):
    cfg = syn.SMALL_CONFIGS[variant]
    dec = build_module(cfg, "cpu")
    g = load_golden(variant)
    ref_keys = {k[len("gnorm_"):] for k in g if k.startswith("gnorm_")}
    assert {k for k, _ in dec.named_parameters()} == ref_keys
    sd = dec.state_dict()
    assert "pos_encoder.pe" in sd and tuple(sd["pos_encoder.pe"].shape) == (5000, 1, cfg.D)
    if variant != "G":
        assert sd["fact_encoder.predicate_embedding.weight"].data_ptr() == sd["predicate_embedding.weight"].data_ptr()


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_forward_backward_vs_golden(variant):
    cfg = syn.SMALL_CONFIGS[variant]
    g = load_golden(variant)
    dec = build_module(cfg, "cpu").eval()
    batch = syn.make_batch(cfg, seed=1)
    batch["encoder_out"].requires_grad_(True)
    scores, caps, dl = dec(*batch_args(cfg, batch))
    assert np.array_equal(caps.numpy(), g["captions_sorted"]) and dl == g["decode_lengths"].tolist()
    assert nmax_err(scores.detach(), g["scores"]) < 1e-4
    loss = orc.caption_loss(scores, caps, dl)
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4
    loss.backward()
    assert nmax_err(batch["encoder_out"].grad, g["grad_encoder_out"]) < 1e-3
    for k, p in dec.named_parameters():
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        ref_norm = float(g[f"gnorm_{k}"])
        assert abs(float(gr.double().norm()) - ref_norm) <= 2e-3 * max(ref_norm, 1e-6), k
        if f"grad_{k}" in g and ref_norm > 1e-12:
            assert nmax_err(gr, g[f"grad_{k}"]) < 2e-3, k


@pytest.mark.parametrize("variant", ["G", "K"])
def test_train_mode_dropout_matches_oracle_with_injected_masks(variant):
    cfg = syn.SMALL_CONFIGS[variant]
    ps = dict(dec=0.3, enc=0.4, pos=0.1)
    dec = build_module(cfg, "cpu", dropouts=(ps["dec"], ps["enc"], ps["pos"])).train()
    batch = syn.make_batch(cfg, seed=3)
    scores, caps, dl = dec(*batch_args(cfg, batch))
    seed = (int(torch.initial_seed()) * 1000003 + dec._step) & 0x7FFFFFFF
    p = oracle_params(cfg, requires_grad=True)
    ref_scores, _, _ = orc.forward(spec_for(cfg), p, *batch_args(cfg, batch), drop=oracle_drop_fn(seed, ps))
    assert nmax_err(scores.detach(), ref_scores.detach()) < 1e-4
    orc.caption_loss(scores, caps, dl).backward()
    orc.caption_loss(ref_scores, caps, dl).backward()
    for k, prm in dec.named_parameters():
        ref = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        got = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        assert float((got - ref).abs().max()) <= 2e-3 * max(float(ref.abs().max()), 1e-6) + 1e-7, k


def test_fused_loss_matches_train_py_loss():
    cfg = syn.SMALL_CONFIGS["K"]
    dec = build_module(cfg, "cpu").eval()
    batch = syn.make_batch(cfg, seed=1, equal_lengths=False)
    with torch.no_grad():
        scores, caps, dl = dec(*batch_args(cfg, batch))
    eng = dec._engine
    acc, ds = eng.loss(scores, caps, torch.tensor(dl, dtype=torch.int32))
    ref = orc.caption_loss(scores, caps, dl)
    assert abs(float(acc[0] / acc[1]) - float(ref)) < 1e-5
    s = scores.clone().requires_grad_(True)
    orc.caption_loss(s, caps, dl).backward()
    W = scores.shape[-1]
    assert nmax_err(ds[:, :W].view_as(s) / acc[1], s.grad) < 1e-5


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_predict_tokens_vs_golden(variant):
    cfg = syn.SMALL_CONFIGS[variant]
    g = load_golden(variant)
    dec = build_module(cfg, "cpu").eval()
    pb = syn.make_batch(cfg, seed=int(g["predict_seed"]))
    T = int(g["predict_max_len"])
    out = dec.predict_batch(pb["encoder_out"], T, pb["entities"], pb.get("facts"))
    assert out.tolist() == g["predict_tokens"].tolist()
    one = dec.predict(pb["encoder_out"][:1], T, pb["entities"][:1], pb["facts"][:1] if cfg.has_facts else None)
    assert tuple(one.shape) == (T, 1) and one.reshape(-1).tolist() == g["predict_tokens"][0].tolist()


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_beam_search_tokens_vs_golden(variant):
    """Extension (the reference decodes greedily): engine.beam_decode's orchestration - position-major cache + ancestor table,
    shared image context, double-buffered histories - over the host kernels against the captions that the tutorial beam search
    produces with the unmodified reference modules as the scoring function (tests/golden/make_golden_beam.py)."""
    import os

    import numpy as np

    from helpers import GOLDEN_DIR

    cfg = syn.SMALL_CONFIGS[variant]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"golden_beam_{variant}.npz")))
    T, k, B = int(g["max_len"]), int(g["beam"]), int(g["batch"])
    pb = syn.make_batch(cfg.with_batch(B), seed=int(g["seed"]))
    for j, bias in enumerate(g["end_bias"].tolist()):
        dec = build_module(cfg, "cpu").eval()
        with torch.no_grad():
            dec.state_dict()["fc_vocab.bias"][cfg.V - 1] += bias
        out, score = dec.beam_search_batch(pb["encoder_out"], T, pb["entities"], pb.get("facts"), beam_size=k, return_scores=True)
        ok = g[f"margins_{j}"] > 1e-4
        assert ok.sum() >= B - 1
        assert out[ok].tolist() == g[f"tokens_{j}"][ok].tolist()
        assert np.allclose(score.numpy()[ok], g[f"scores_{j}"][ok], atol=1e-3)


@pytest.mark.parametrize("variant", ["G", "K"])
def test_fused_decode_chain_orchestration(variant, monkeypatch):
    """The decode loops with the row-wise layer tails as single ick_decode_chain launches (opt-in on the GPU: ICK_DECODE_CHAIN=1),
    driven over the host kernels in fp32: greedy tokens and beam captions must equal the golden ones of the unfused orchestration."""
    import os

    import numpy as np

    from helpers import GOLDEN_DIR

    monkeypatch.setenv("ICK_DECODE_CHAIN", "force")
    cfg = syn.SMALL_CONFIGS[variant]
    g = load_golden(variant)
    dec = build_module(cfg, "cpu").eval()
    pb = syn.make_batch(cfg, seed=int(g["predict_seed"]))
    T = int(g["predict_max_len"])
    calls0 = dec._ensure_engine().K.calls
    out = dec.predict_batch(pb["encoder_out"], T, pb["entities"], pb.get("facts"))
    assert out.tolist() == g["predict_tokens"].tolist()
    fused_calls = dec._ensure_engine().K.calls - calls0
    monkeypatch.setenv("ICK_DECODE_CHAIN", "0")
    calls0 = dec._ensure_engine().K.calls
    dec.predict_batch(pb["encoder_out"], T, pb["entities"], pb.get("facts"))
    assert fused_calls < dec._ensure_engine().K.calls - calls0  # fewer launches
    monkeypatch.setenv("ICK_DECODE_CHAIN", "force")
    gb = dict(np.load(os.path.join(GOLDEN_DIR, f"golden_beam_{variant}.npz")))
    Tb, k, B = int(gb["max_len"]), int(gb["beam"]), int(gb["batch"])
    pbb = syn.make_batch(cfg.with_batch(B), seed=int(gb["seed"]))
    outb = dec.beam_search_batch(pbb["encoder_out"], Tb, pbb["entities"], pbb.get("facts"), beam_size=k)
    ok = gb["margins_0"] > 1e-4
    assert outb[ok].tolist() == gb["tokens_0"][ok].tolist()


def test_module_pickles_like_reference_checkpoints(tmp_path):
    cfg = syn.SMALL_CONFIGS["G"]
    dec = build_module(cfg, "cpu").eval()
    batch = syn.make_batch(cfg, seed=1)
    with torch.no_grad():
        s0, _, _ = dec(*batch_args(cfg, batch))
    path = tmp_path / "ckpt.pth.tar"
    torch.save({"decoder": dec}, path)  # G/utils.py:32-46 pickles whole modules
    dec2 = torch.load(path, weights_only=False)["decoder"]
    with torch.no_grad():
        s1, _, _ = dec2(*batch_args(cfg, batch))
    assert torch.equal(s0, s1)


@pytest.mark.parametrize("kind", ["pair", "attn"])
@pytest.mark.parametrize("p", [0.1, 0.3, 0.5])
def test_dropout_hash_statistics(p, kind):
    """The counter hashes behind every dropout mask (csrc/common.cuh, ported in dropout_ref.py) - the pair hash of the
    activation sites and the bit-parallel keep words of the attention probabilities: keep rate within sampling error of 1-p, and
    no visible correlation between the two columns of a pair, neighbouring pairs, neighbouring rows or neighbouring key groups."""
    from dropout_ref import attn_drop_mul, drop_mul

    rows, cols = 2048, 640
    k = ((drop_mul if kind == "pair" else attn_drop_mul)(p, 1234, 7, rows, cols) > 0).double()
    thr = int(p * 32768)
    assert abs(float(k.mean()) - (1 - thr / 32768)) < 4 * (p * (1 - p) / (rows * cols)) ** 0.5
    kc = k - k.mean()

    def corr(a, b):
        return float((a * b).mean() / ((a * a).mean() * (b * b).mean()).sqrt())

    lim = 5.0 / (rows * cols / 2) ** 0.5
    assert abs(corr(kc[:, 0::2], kc[:, 1::2])) < lim       # the two 15-bit fields of one hash
    assert abs(corr(kc[:, 0:-2:2], kc[:, 2::2])) < lim     # consecutive pairs
    assert abs(corr(kc[:-1], kc[1:])) < lim                # consecutive rows
    assert abs(corr(kc[:, :-8], kc[:, 8:])) < lim          # the stride an MMA fragment lane sees
    assert abs(corr(kc[:, :-32], kc[:, 32:])) < lim        # the same bit of consecutive keep words
    assert abs(corr(kc[:, :-1], kc[:, 1:])) < lim          # bits i and 16 + i / i + 1 of one word
    # each row / column keeps its own share close to 1-p
    assert float(k.mean(1).std()) < 1.3 * (p * (1 - p) / cols) ** 0.5
    assert float(k.mean(0).std()) < 1.3 * (p * (1 - p) / rows) ** 0.5


def test_beam_search_edge_cases():
    """One image, one step, widths 1 and 8 (the widest supported), and equality with the oracle on each."""
    from helpers import oracle_params, spec_for
    from oracle import decoder_oracle as orc

    cfg = syn.SMALL_CONFIGS["K"]
    dec = build_module(cfg, "cpu").eval()
    p = oracle_params(cfg)
    pb = syn.make_batch(cfg.with_batch(2), seed=5)
    for T, k in ((1, 1), (1, 8), (5, 8), (4, 2)):
        out, score = dec.beam_search_batch(pb["encoder_out"][:1], T, pb["entities"][:1], pb["facts"][:1], beam_size=k, return_scores=True)
        with torch.no_grad():
            ref, margin = orc.beam_search(spec_for(cfg), p, pb["encoder_out"][:1], T, pb["entities"][:1], pb["facts"][:1], beam_size=k,
                                          return_margin=True)
        assert tuple(out.shape) == (1, T) and torch.isfinite(score).all()
        assert margin < 1e-4 or out[0].tolist() == ref.tolist(), (T, k)
    with pytest.raises(AssertionError):
        dec.beam_search_batch(pb["encoder_out"][:1], 3, pb["entities"][:1], pb["facts"][:1], beam_size=9)


def test_encoder_pretrained_weights_missing_is_loud(monkeypatch):
    """Encoder() asks for the ImageNet weights like the reference (G/models.py:24); without network or cache that must raise (or,
    with ICKB200_ALLOW_RANDOM_TRUNK=1, warn) instead of silently training on a random trunk."""
    import torchvision

    from ickb200.geo_aware import Encoder

    def boom(*a, **k):
        if k.get("weights") is not None or k.get("pretrained"):
            raise OSError("no network")
        return orig(*a, **k)

    orig = torchvision.models.resnet101
    monkeypatch.setattr(torchvision.models, "resnet101", boom)
    monkeypatch.delenv("ICKB200_ALLOW_RANDOM_TRUNK", raising=False)
    with pytest.raises(RuntimeError, match="pretrained=False"):
        Encoder()
    monkeypatch.setenv("ICKB200_ALLOW_RANDOM_TRUNK", "1")
    with pytest.warns(UserWarning, match="RANDOM trunk"):
        Encoder()
