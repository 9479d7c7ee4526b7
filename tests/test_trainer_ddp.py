"""
CPU: the fused train step (trainer.Trainer) over the host simulation:
  * one step equals the reference recipe  loss.backward(); clip_gradient(+-5); Adam.step()  done with torch on the oracle,
  * two gloo ranks, each with half of the batch, end with the same parameters as one process with the whole batch
    (gradient all-reduce before the clamp, normalisation by the global kept-token count — SURVEY.md §8e).
"""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import batch_args, build_module, nmax_err, oracle_params, spec_for
from hostsim import HostKernels
from ickb200 import models as M, synthetic as syn
from ickb200.trainer import Trainer
from oracle import decoder_oracle as orc


def _trainer_step(cfg, batch, steps=2, distributed=False, pg=None, overlap=False):
    M.DecoderTransformer._test_kernel_factory = HostKernels
    dec = build_module(cfg, "cpu", dropouts=(0.0, 0.0, 0.0)).train()
    tr = Trainer(dec, lr=4e-4, grad_clip=5.0, distributed=distributed, process_group=pg, overlap_allreduce=overlap)
    accs = []
    for _ in range(steps):
        accs.append(tr.train_step(*batch_args(cfg, batch)).clone())
    return dec, accs, tr


def test_fused_step_equals_reference_recipe():
    cfg = syn.SMALL_CONFIGS["K"]
    batch = syn.make_batch(cfg, seed=4, equal_lengths=False)
    dec, accs, tr = _trainer_step(cfg, batch, steps=2)
    # reference recipe on the oracle: G/train.py:275-292 with ut.clip_gradient (G/utils.py:75-85)
    p = oracle_params(cfg, requires_grad=True)
    names = [k for k in p if k != "pos_encoder.pe"]
    opt = torch.optim.Adam([p[k] for k in names], lr=4e-4)
    losses = []
    for _ in range(2):
        scores, caps, dl = orc.forward(spec_for(cfg), p, *batch_args(cfg, batch))
        loss = orc.caption_loss(scores, caps, dl)
        opt.zero_grad()
        loss.backward()
        for k in names:
            if p[k].grad is not None:
                p[k].grad.data.clamp_(-5.0, 5.0)
        opt.step()
        losses.append(float(loss))
    for a, l in zip(accs, losses):
        assert abs(float(a[0] / a[1]) - l) < 1e-4
    eng = dec._engine
    for k, prm in dec.named_parameters():
        ref = p[k].detach()
        # Adam normalises by sqrt(v): where |g| ~ eps the update direction is fp32 noise, so the weights are compared at a
        # fraction of the 2*lr total step, and the first/second moments (linear/quadratic in g) are compared tightly.
        st = opt.state[p[k]]
        m_ref, v_ref = st["exp_avg"], st["exp_avg_sq"]
        d = (prm.detach() - ref).abs()
        sig = v_ref.sqrt() > 1e-3 * max(float(v_ref.sqrt().max()), 1e-12)  # e.g. the key bias has an exactly-zero true gradient
        assert float(d.max()) < 1e-3, k
        if bool(sig.any()):
            assert float((d[sig] < 2e-5).float().mean()) > 0.98, k
        assert float((eng.param(k, tr.m) - m_ref).abs().max()) <= 2e-3 * float(m_ref.abs().max()) + 1e-9, k
        assert float((eng.param(k, tr.v) - v_ref).abs().max()) <= 4e-3 * float(v_ref.abs().max()) + 1e-12, k
    # the packed operand copies follow the master weights after the fused step
    lin = eng.lin["fc_vocab"]
    assert torch.allclose(lin.W[:, : cfg.D], dec._get("fc_vocab.weight").detach(), atol=0)


def _worker(rank, world, port, cfg, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    full = syn.make_batch(cfg, seed=4, equal_lengths=False)
    n = cfg.B // world
    shard = {k: v[rank * n : (rank + 1) * n] for k, v in full.items()}
    dec, accs, _ = _trainer_step(cfg.with_batch(n), shard, steps=2, distributed=True)
    dec_o, accs_o, _ = _trainer_step(cfg.with_batch(n), shard, steps=2, distributed=True, overlap=True)  # region-wise all-reduce
    same = all(torch.equal(a, b) for a, b in zip(accs, accs_o)) and all(
        torch.allclose(p1.detach(), p2.detach(), atol=1e-7) for (_, p1), (_, p2) in zip(dec.named_parameters(), dec_o.named_parameters()))
    if rank == 0:
        q.put(({k: v.detach().numpy().copy() for k, v in dec.named_parameters()}, [a.tolist() for a in accs], same))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process():
    cfg = syn.SMALL_CONFIGS["K"].with_batch(4)
    full = syn.make_batch(cfg, seed=4, equal_lengths=False)
    dec1, accs1, _ = _trainer_step(cfg, full, steps=2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, q)) for r in range(2)]
    for p in procs:
        p.start()
    params2, accs2, overlap_same = q.get(timeout=600)
    assert overlap_same  # the region-wise (overlapped) all-reduce gives the update of the single collective
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    for a1, a2 in zip(accs1, accs2):
        assert abs(float(a1[1]) - a2[1]) == 0 and abs(float(a1[0]) - a2[0]) < 1e-3 * abs(a2[0])
    n_close = n_all = 0
    for k, prm in dec1.named_parameters():
        d = (prm.detach() - torch.from_numpy(params2[k])).abs()
        assert float(d.max()) < 1e-3, k  # never more than the 2*lr two Adam steps can move a weight
        n_close += int((d < 2e-5).sum())
        n_all += d.numel()
    assert n_close > 0.999 * n_all  # all but noise-gradient elements agree to fp32 summation order


def _gen_worker(rank, world, port, cfg, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from ickb200.trainer import generate_sharded

    M.DecoderTransformer._test_kernel_factory = HostKernels
    dec = build_module(cfg, "cpu").eval()
    pb = syn.make_batch(cfg, seed=11)
    greedy = generate_sharded(dec, pb["encoder_out"], 6, pb["entities"], pb.get("facts"))
    beam = generate_sharded(dec, pb["encoder_out"], 6, pb["entities"], pb.get("facts"), beam_size=3)
    if rank == 0:
        q.put((greedy.tolist(), beam.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_generation_two_ranks_equals_single_process():
    """Inference shards images i % world == rank with no data-path communication: 5 images over 2 gloo ranks
    (a ragged split) give the single-process captions, greedy and beam."""
    from ickb200.trainer import generate_sharded

    cfg = syn.SMALL_CONFIGS["K"].with_batch(5)
    M.DecoderTransformer._test_kernel_factory = HostKernels
    dec = build_module(cfg, "cpu").eval()
    pb = syn.make_batch(cfg, seed=11)
    ref_g = generate_sharded(dec, pb["encoder_out"], 6, pb["entities"], pb.get("facts")).tolist()
    ref_b = generate_sharded(dec, pb["encoder_out"], 6, pb["entities"], pb.get("facts"), beam_size=3).tolist()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gen_worker, args=(r, 2, port, cfg, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_g, got_b = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert got_g == ref_g and got_b == ref_b


def test_validate_step_is_the_eval_mode_loss():
    """Trainer.validate_step = train.py's validate() body (G/train.py:317-386): decoder in eval mode, same packed CE, no update."""
    cfg = syn.SMALL_CONFIGS["K"]
    batch = syn.make_batch(cfg, seed=6, equal_lengths=False)
    M.DecoderTransformer._test_kernel_factory = HostKernels
    dec = build_module(cfg, "cpu").train()  # default dropouts 0.5/0.5/0.1: validate_step must not apply them
    tr = Trainer(dec, lr=4e-4, grad_clip=5.0)
    before = {k: v.detach().clone() for k, v in dec.named_parameters()}
    acc = tr.validate_step(*batch_args(cfg, batch))
    p = oracle_params(cfg)
    with torch.no_grad():
        scores, caps, dl = orc.forward(spec_for(cfg), p, *batch_args(cfg, batch))
        ref = orc.caption_loss(scores, caps, dl)
    assert abs(float(acc[0] / acc[1]) - float(ref)) < 1e-4
    assert float(acc[1]) == float(sum(1 for b, n in enumerate(dl) for t in range(n) if int(caps[b, t + 1]) != 0))
    for k, v in dec.named_parameters():
        assert torch.equal(v.detach(), before[k]), k


def test_trainer_state_dict_resumes_exactly(tmp_path):
    """Checkpoint / resume of the fused optimizer (the reference pickles its optimizer, G/utils.py:32-46, G/train.py:102-129): two
    steps, save decoder + trainer state, reload into fresh objects, one more step == three uninterrupted steps, bit for bit
    (train-mode dropout included: the mask stream is a function of seed base + step)."""
    cfg = syn.SMALL_CONFIGS["G"]
    batch = syn.make_batch(cfg, seed=8, equal_lengths=False)
    M.DecoderTransformer._test_kernel_factory = HostKernels
    torch.manual_seed(1234)
    dec_a = build_module(cfg, "cpu", dropouts=(0.3, 0.3, 0.1)).train()
    tr_a = Trainer(dec_a, lr=4e-4, grad_clip=5.0)
    for _ in range(3):
        acc_a = tr_a.train_step(*batch_args(cfg, batch)).clone()
    torch.manual_seed(1234)
    dec_b = build_module(cfg, "cpu", dropouts=(0.3, 0.3, 0.1)).train()
    tr_b = Trainer(dec_b, lr=4e-4, grad_clip=5.0)
    for _ in range(2):
        tr_b.train_step(*batch_args(cfg, batch))
    path = tmp_path / "ckpt.pth.tar"
    torch.save({"decoder": dec_b, "trainer": tr_b.state_dict()}, path)
    ck = torch.load(path, weights_only=False)
    dec_c = ck["decoder"].train()
    tr_c = Trainer(dec_c, lr=1.0, grad_clip=None)  # deliberately wrong hyper-parameters: load_state_dict must restore them
    tr_c.load_state_dict(ck["trainer"])
    acc_c = tr_c.train_step(*batch_args(cfg, batch)).clone()
    assert torch.equal(acc_a, acc_c)
    for (k, pa), (_, pc) in zip(dec_a.named_parameters(), dec_c.named_parameters()):
        assert torch.equal(pa.detach(), pc.detach()), k


def _trim_worker(rank, world, port, cfg, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    full = _ragged_batch(cfg)
    n = cfg.B // world
    shard = {k: v[rank * n : (rank + 1) * n] for k, v in full.items()}
    M.DecoderTransformer._test_kernel_factory = HostKernels
    torch.manual_seed(100 + rank)  # ranks seeded DIFFERENTLY on purpose: the Trainer must broadcast rank 0's parameters
    dec = build_module(cfg.with_batch(n), "cpu", dropouts=(0.0, 0.0, 0.0), seed=rank).train()
    tr = Trainer(dec, lr=4e-4, grad_clip=5.0, distributed=True, trim_padding=True)
    local_w = tr.trimmed_width(shard["captions"])
    inp = tr.prepare(*batch_args(cfg.with_batch(n), shard))
    acc = tr.step(inp).clone()
    q.put((rank, local_w, int(inp.captions.shape[1]), acc.tolist(), int(tr.seed_base),
           {k: v.detach().numpy().copy() for k, v in dec.named_parameters()} if rank == 0 else None))
    dist.barrier()
    dist.destroy_process_group()


def _ragged_batch(cfg):
    full = syn.make_batch(cfg, seed=4)
    full["captions"][:2, 9:] = 0  # rank 0's shard: short captions (width 16 after rounding) ...
    full["caption_masks"][:2, 9:] = 0
    full["captions"][2:, 30:] = 0  # ... rank 1's: long ones (width 32)
    full["caption_masks"][2:, 30:] = 0
    return full


def test_trim_padding_width_is_a_collective_decision():
    """with trim_padding under data parallelism the trimmed width must be agreed by all ranks (graphs are
    keyed by the caption shape; a rank meeting a new width alone would capture - and all-reduce once more - while its peers
    replay).  Two gloo ranks whose LOCAL widths differ (16 vs 32) must both run at 32, and the update must equal the
    single-process step on the whole batch.  Ranks are built from different seeds: rank 0's parameters must win."""
    cfg = syn.Config("K", B=4, T=40, E=13, F=7, V=57, P=20)
    full = _ragged_batch(cfg)
    M.DecoderTransformer._test_kernel_factory = HostKernels
    dec1 = build_module(cfg, "cpu", dropouts=(0.0, 0.0, 0.0), seed=0).train()
    tr1 = Trainer(dec1, lr=4e-4, grad_clip=5.0, trim_padding=True)
    acc1 = tr1.train_step(*batch_args(cfg, full)).clone()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_trim_worker, args=(r, 2, port, cfg, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    (_, w0, used0, acc_a, seed0, params), (_, w1, used1, acc_b, seed1, _) = res
    assert (w0, w1) == (16, 32) and used0 == used1 == 32
    assert acc_a == acc_b and acc_a[1] == float(acc1[1]) and abs(acc_a[0] - float(acc1[0])) < 1e-3 * abs(float(acc1[0]))
    assert seed0 != seed1  # rank-mixed dropout seeds: the shards must not draw identical masks
    n_close = n_all = 0
    for k, prm in dec1.named_parameters():
        d = (prm.detach() - torch.from_numpy(params[k])).abs()
        assert float(d.max()) < 1e-3, k
        n_close += int((d < 2e-5).sum())
        n_all += d.numel()
    assert n_close > 0.999 * n_all


def test_writes_through_data_need_repack_and_mode_switch_repacks():
    """`.data` writes (the reference's own init idiom, G/models.py:264-272) do not bump the autograd version; the
    public repack() - and every train()/eval() switch - refreshes the packed operand copies."""
    cfg = syn.SMALL_CONFIGS["G"]
    M.DecoderTransformer._test_kernel_factory = HostKernels
    dec = build_module(cfg, "cpu").eval()
    batch = syn.make_batch(cfg, seed=1)
    s0, _, _ = dec(*batch_args(cfg, batch))
    dec._get("fc_vocab.weight").data.mul_(2.0)
    dec._get("fc_vocab.bias").data.zero_()
    dec.repack()
    s1, _, _ = dec(*batch_args(cfg, batch))
    V = cfg.V
    assert nmax_err(s1[..., :V], 2.0 * (s0[..., :V] - syn.det_weights({"fc_vocab.bias": (V,)})["fc_vocab.bias"])) < 1e-5
    dec._get("fc_vocab.weight").data.mul_(0.5)
    dec.train().eval()  # a mode switch alone is enough
    s2, _, _ = dec(*batch_args(cfg, batch))
    assert nmax_err(s2[..., :V] , s0[..., :V] - syn.det_weights({"fc_vocab.bias": (V,)})["fc_vocab.bias"]) < 1e-5


def test_load_state_dict_and_frozen_set_invalidate_captured_steps():
    cfg = syn.SMALL_CONFIGS["G"]
    M.DecoderTransformer._test_kernel_factory = HostKernels
    dec = build_module(cfg, "cpu").train()
    tr = Trainer(dec, lr=4e-4)
    tr._graphs[("stale",)] = (object(), object())
    tr.load_state_dict(tr.state_dict())
    assert not tr._graphs and tr._graph is None
    assert "word_embedding.weight" not in tr._frozen()
    dec.fine_tune_embeddings(False)  # G/models.py:282-289, called after the Trainer was built
    assert "word_embedding.weight" in tr._frozen()
    before = dec._get("word_embedding.weight").detach().clone()
    tr.train_step(*batch_args(cfg, syn.make_batch(cfg, seed=1)))
    assert torch.equal(dec._get("word_embedding.weight").detach(), before)
