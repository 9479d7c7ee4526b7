"""Debug driver for the fused bf16 attention backward: one shape against the host simulation.
    python tools/attn_bwd_check.py B H Sq Sk causal p"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ickb200  # noqa
from hostsim import HostKernels
from ickb200.kernels import CudaKernels

B, H, Sq, Sk, causal, p = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), bool(int(sys.argv[5])), float(sys.argv[6])
dh, bf = 30, torch.bfloat16
K, Hk = CudaKernels(), HostKernels()
g = torch.Generator().manual_seed(1)


def heads(rows):
    x = torch.randn(rows, H * 32, generator=g).view(rows, H, 32)
    x[:, :, dh:] = 0
    return x.view(rows, H * 32).to(bf)


q, k, v, do = heads(B * Sq), heads(B * Sk), heads(B * Sk), heads(B * Sq)
if len(sys.argv) > 7:  # sharper score distributions (exercises the lazily rescaled accumulator of the tcgen05 forward)
    q = (q.float() * float(sys.argv[7])).to(bf)
drop = (p, 7, 9) if p > 0 else None
o, lse = torch.zeros(B * Sq, H * 32, dtype=bf), torch.zeros(B * H * Sq)
Hk.mha_fwd(q, k, v, o, lse, B, H, Sq, Sk, dh, causal, drop)
og, lg = torch.full((B * Sq, H * 32), float("nan"), dtype=bf).cuda(), torch.zeros(B * H * Sq).cuda()
K.mha_fwd(q.cuda(), k.cuda(), v.cuda(), og, lg, B, H, Sq, Sk, dh, causal, drop)
torch.cuda.synchronize()
of = og.float().cpu()
print("O  nan", int(torch.isnan(of).sum()), "max err", round(float((of - o.float()).abs().max() / o.float().abs().max()), 5), "fro err",
      round(float((of - o.float()).norm() / o.float().norm()), 5), "| lse max abs err", round(float((lg.cpu() - lse).abs().max()), 5))
ref = [torch.zeros(B * Sq, H * 32, dtype=bf), torch.zeros(B * Sk, H * 32, dtype=bf), torch.zeros(B * Sk, H * 32, dtype=bf)]
ds = torch.zeros(B * H * Sq)
Hk.mha_bwd(q, k, v, o, do, lse, ds, *ref, B, H, Sq, Sk, dh, causal, drop)
out = [torch.full_like(r, float("nan")).cuda() for r in ref]
dsg = torch.zeros(B * H * Sq).cuda()
K.mha_bwd(q.cuda(), k.cuda(), v.cuda(), o.cuda(), do.cuda(), lse.cuda(), dsg, *out, B, H, Sq, Sk, dh, causal, drop)
torch.cuda.synchronize()
for a, b, name in zip(out, ref, ["dQ", "dK", "dV"]):
    a = a.float().cpu()
    e = float((a - b.float()).abs().max() / b.float().abs().max())
    fro = float((a - b.float()).norm() / b.float().norm())
    print(name, "nan", int(torch.isnan(a).sum()), "max err", round(e, 5), "fro err", round(fro, 5), "norm ratio", round(float(a.norm() / b.float().norm()), 5))
