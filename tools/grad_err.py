import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import *
from ickb200 import synthetic as syn
from oracle import decoder_oracle as orc
variant='K'; dtype=torch.float32
cfg = syn.BASE_PARITY_CONFIGS[variant].with_batch(4)
ps = dict(dec=0.5, enc=0.5, pos=0.1)
for sd in (1234, 7):
  for pset in (ps, dict(dec=0.0,enc=0.0,pos=0.0)):
    torch.manual_seed(sd)
    dec = build_module(cfg, "cuda", dtype, dropouts=(pset["dec"], pset["enc"], pset["pos"]), profile="reference").train()
    b = syn.make_batch(cfg, seed=23)
    bd = dict(b)
    for k in ("captions","encoder_out","caption_masks","caption_lengths","facts"): bd[k]=bd[k].cuda()
    scores, caps, dl = dec(*batch_args(cfg, bd))
    seed = (int(torch.initial_seed()) * 1000003 + dec._step) & 0x7FFFFFFF
    p = oracle_params(cfg, requires_grad=True, profile="reference")
    ref_scores,_,_ = orc.forward(spec_for(cfg), p, *batch_args(cfg, b), drop=oracle_drop_fn(seed, pset))
    print("seed",sd,"p",pset["dec"],"scores nmax", nmax_err(scores.detach().cpu(), ref_scores.detach()))
    orc.caption_loss(scores, caps, dl).backward(); orc.caption_loss(ref_scores, caps.cpu(), dl).backward()
    errs=[]
    for k, prm in dec.named_parameters():
        ref = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        got = (prm.grad if prm.grad is not None else torch.zeros_like(prm)).float().cpu()
        errs.append((float((got-ref).abs().max())/max(float(ref.abs().max()),1e-9), k))
    errs.sort(reverse=True)
    print("   worst:", [(round(e,5),k) for e,k in errs[:4]])
    # where do the out-of-tolerance elements of the worst parameters sit?
    for e, k in errs[:3]:
        prm = dict(dec.named_parameters())[k]
        ref = p[k].grad; got = prm.grad.float().cpu()
        d = (got - ref).abs(); tol = 2e-3 * float(ref.abs().max()) + 1e-7
        bad = d > tol
        print("   ", k, tuple(ref.shape), "bad", int(bad.sum()), "of", bad.numel(), "max", float(d.max()), "tol", tol,
              "rel-l2", float((got - ref).norm() / ref.norm()))
