"""
CPU: the oracle restatement (oracle/decoder_oracle.py) against the golden vectors produced by the unmodified
reference modules (tests/golden/make_golden.py) — forward scores, train.py loss, parameter gradients, greedy predict.
"""
import numpy as np
import pytest
import torch

from helpers import batch_args, load_golden, nmax_err, oracle_params, spec_for
from ickb200 import synthetic as syn
from oracle import decoder_oracle as orc


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_inputs_and_weights_regenerate(variant):
    cfg = syn.SMALL_CONFIGS[variant]
    g = load_golden(variant)
    batch = syn.make_batch(cfg, seed=1)
    for k, v in batch.items():
        assert np.array_equal(v.numpy(), g[f"in_{k}"]), k
    p = oracle_params(cfg)
    chk = sum(float(v.double().abs().sum()) for k, v in sorted(p.items()) if not k.endswith("pe"))
    assert abs(chk - float(g["weights_checksum"])) <= 1e-9 * abs(chk)


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_forward_loss_grads(variant):
    cfg = syn.SMALL_CONFIGS[variant]
    g = load_golden(variant)
    p = oracle_params(cfg, requires_grad=True)
    batch = syn.make_batch(cfg, seed=1)
    batch["encoder_out"].requires_grad_(True)
    scores, caps, dl = orc.forward(spec_for(cfg), p, *batch_args(cfg, batch))
    assert np.array_equal(caps.numpy(), g["captions_sorted"])
    assert dl == g["decode_lengths"].tolist()
    # fp32, tolerance 1e-4 relative (north_star), measured ~1e-6
    assert nmax_err(scores.detach(), g["scores"]) < 1e-4
    loss = orc.caption_loss(scores, caps, dl)
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    assert nmax_err(batch["encoder_out"].grad, g["grad_encoder_out"]) < 1e-3
    for k, v in p.items():
        if k == "pos_encoder.pe":
            continue
        gr = v.grad if v.grad is not None else torch.zeros_like(v)
        ref_norm = float(g[f"gnorm_{k}"])
        assert abs(float(gr.double().norm()) - ref_norm) <= 1e-3 * max(ref_norm, 1e-6), k
        if f"grad_{k}" in g:
            assert nmax_err(gr, g[f"grad_{k}"]) < 1e-3 or ref_norm < 1e-12, k
        else:
            rows = gr.reshape(gr.shape[0], -1)[:: max(1, gr.shape[0] // 7)][:, :64]
            ref = g[f"gradrows_{k}"]
            assert float((rows.double() - torch.as_tensor(ref).double()).abs().max()) <= 1e-3 * max(
                float(np.abs(ref).max()), 1e-6) + 1e-9, k


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_predict_tokens(variant):
    cfg = syn.SMALL_CONFIGS[variant]
    g = load_golden(variant)
    p = oracle_params(cfg)
    pb = syn.make_batch(cfg, seed=int(g["predict_seed"]))
    T = int(g["predict_max_len"])
    with torch.no_grad():
        for i in range(min(2, cfg.B)):  # the rest of the golden images are covered by the GPU tests
            out, margins = orc.predict(
                spec_for(cfg), p, pb["encoder_out"][i : i + 1], T, pb["entities"][i : i + 1],
                pb["facts"][i : i + 1] if cfg.has_facts else None, return_margins=True)
            assert out.reshape(-1).tolist() == g["predict_tokens"][i].tolist(), (i, min(margins))


def test_indicator_predict_mode_has_no_lag():
    # K/models.py:406-409: with out_length == 1 every position counts, including the last one
    cfg = syn.SMALL_CONFIGS["K"]
    sp = spec_for(cfg)
    facts = torch.zeros(1, cfg.F, 3, dtype=torch.long)
    facts[0, :, 1] = torch.arange(cfg.F) % cfg.E
    facts[0, :, 2] = torch.arange(cfg.F)
    caps = torch.full((1, 5), cfg.V - 2)
    caps[0, 4] = cfg.V + 3
    eb1, pi1 = orc.context_indicators(sp, caps, facts, cfg.E, 1)
    assert eb1[0, 0, 3] == 1 and pi1[0, 0, 3] == 1 and eb1.sum() == 1
    ebT, _ = orc.context_indicators(sp, caps, facts, cfg.E, 5)
    assert ebT.sum() == 0  # teacher-forced: only strictly later positions, and there are none


def load_beam_golden(variant):
    import os

    from helpers import GOLDEN_DIR

    return dict(np.load(os.path.join(GOLDEN_DIR, f"golden_beam_{variant}.npz")))


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_beam_search_vs_reference_scored_golden(variant):
    """Extension (the reference has no beam search): the oracle's beam search against captions computed with the unmodified
    reference modules as the scoring function (tests/golden/make_golden_beam.py)."""
    cfg = syn.SMALL_CONFIGS[variant]
    g = load_beam_golden(variant)
    T, k = int(g["max_len"]), int(g["beam"])
    pb = syn.make_batch(cfg.with_batch(int(g["batch"])), seed=int(g["seed"]))
    for j, bias in enumerate(g["end_bias"].tolist()):
        p = oracle_params(cfg)
        p["fc_vocab.bias"][cfg.V - 1] += bias
        with torch.no_grad():
            for i in (0, 3, 4):
                out, margin = orc.beam_search(spec_for(cfg), p, pb["encoder_out"][i : i + 1], T, pb["entities"][i : i + 1],
                                              pb["facts"][i : i + 1] if cfg.has_facts else None, beam_size=k, return_margin=True)
                assert abs(margin - float(g[f"margins_{j}"][i])) < 1e-4
                assert out.tolist() == g[f"tokens_{j}"][i].tolist(), (j, i, margin)


def test_beam_of_one_is_greedy_without_cleanup():
    cfg = syn.SMALL_CONFIGS["K"]
    g = load_golden("K")
    p = oracle_params(cfg)
    pb = syn.make_batch(cfg, seed=int(g["predict_seed"]))
    with torch.no_grad():
        out = orc.beam_search(spec_for(cfg), p, pb["encoder_out"][:1], 4, pb["entities"][:1], pb["facts"][:1], beam_size=1)
    # step 3 repeats token 27 and predict()'s clean-up rewrites it; the first three steps are plain argmax
    assert out.tolist()[:3] == g["predict_tokens"][0][:3].tolist() and out.tolist()[3] == out.tolist()[2]


def test_image_prep_oracle_is_the_reference_host_pipeline():
    """oracle.prepare_images against the very calls the reference makes: torch.FloatTensor(imgs[i] / 255.) on an fp16 numpy
    array (G/datasets.py:44) then torchvision's transforms.Normalize (G/train.py:139-147)."""
    import torchvision.transforms as T

    rng = np.random.default_rng(0)
    a = (rng.random((3, 3, 16, 24)) * 255).astype(np.float16)
    norm = T.Compose([T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ref = torch.stack([norm(torch.FloatTensor(a[i] / 255.0)) for i in range(a.shape[0])])
    assert torch.equal(orc.prepare_images(a), ref)


# ---- BASELINE shapes (E, F, T, V of SURVEY.md §8d, batch 8): fixtures of the unmodified reference, reference-scale weights --------
def check_scores_against_base_golden(scores, g, V, tol, what=""):
    """Strided grid over all columns, all pointer columns at every third position (normalised max error), row logsumexp / max
    (absolute, on the logit scale) of `scores` against the BASELINE-shape fixture."""
    from helpers import score_views

    v = score_views(scores, V)
    scale = float(np.abs(g["s_grid"]).max())
    assert nmax_err(v["s_grid"], g["s_grid"]) < tol, what
    assert float(np.abs(v["s_ptr"] - g["s_ptr"]).max()) < tol * scale, what
    assert float(np.abs(v["s_lse"] - g["s_lse"]).max()) < tol * max(scale, 1.0), what
    assert float(np.abs(v["s_rowmax"] - g["s_rowmax"]).max()) < tol * scale, what
    return v


def check_grads_against_base_golden(named_grads, g, tol):
    for k, gr in named_grads:
        gr = gr.detach().cpu().float()
        ref_norm = float(g[f"gnorm_{k}"])
        assert abs(float(gr.double().norm()) - ref_norm) <= tol * max(ref_norm, 1e-6), k
        if f"grad_{k}" in g:
            assert nmax_err(gr, g[f"grad_{k}"]) < tol or ref_norm < 1e-12, k
        elif f"gradrows_{k}" in g:
            rows = gr.reshape(gr.shape[0], -1)[:: max(1, gr.shape[0] // 7)][:, :64]
            ref = g[f"gradrows_{k}"]
            assert float((rows.double() - torch.as_tensor(ref).double()).abs().max()) <= tol * max(float(np.abs(ref).max()), 1e-6) + 1e-9, k


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_baseline_shape_forward_loss_grads(variant):
    from helpers import load_base_golden

    cfg = syn.BASE_PARITY_CONFIGS[variant]
    g = load_base_golden(variant)
    p = oracle_params(cfg, requires_grad=True, profile="reference")
    chk = sum(float(v.detach().double().abs().sum()) for k, v in sorted(p.items()) if not k.endswith("pe"))
    assert abs(chk - float(g["weights_checksum"])) <= 1e-9 * abs(chk)
    batch = syn.make_batch(cfg, seed=int(g["seed"]))
    for k, v in batch.items():
        assert abs(float(v.double().abs().sum()) - float(g[f"insum_{k}"])) <= 1e-9 * max(1.0, float(g[f"insum_{k}"])), k
    batch["encoder_out"].requires_grad_(True)
    scores, caps, dl = orc.forward(spec_for(cfg), p, *batch_args(cfg, batch))
    assert np.array_equal(caps.numpy(), g["captions_sorted"]) and dl == g["decode_lengths"].tolist()
    v = check_scores_against_base_golden(scores, g, cfg.V, 1e-4)
    assert float((v["s_argmax"] == g["s_argmax"]).mean()) > 0.999
    loss = orc.caption_loss(scores, caps, dl)
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    ge = batch["encoder_out"].grad
    assert nmax_err(ge[:, ::29, ::11], g["grad_encoder_out_rows"]) < 1e-3
    assert abs(float(ge.double().norm()) - float(g["gnorm_encoder_out"])) <= 1e-3 * float(g["gnorm_encoder_out"])
    check_grads_against_base_golden(((k, v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()
                                     if k != "pos_encoder.pe"), g, 1e-3)


@pytest.mark.parametrize("variant", ["G", "K", "N"])
def test_baseline_shape_predict_tokens(variant):
    from helpers import load_base_golden

    cfg = syn.BASE_PARITY_CONFIGS[variant]
    g = load_base_golden(variant)
    p = oracle_params(cfg, profile="reference")
    pb = syn.make_batch(cfg, seed=int(g["predict_seed"]))
    T = int(g["predict_max_len"])
    with torch.no_grad():  # one image here (0.4 s of un-cached decoding); the GPU tests cover all four
        out = orc.predict(spec_for(cfg), p, pb["encoder_out"][:1], T, pb["entities"][:1], pb["facts"][:1] if cfg.has_facts else None)
    assert out.reshape(-1).tolist() == g["predict_tokens"][0].tolist()
