"""Times the bf16 attention forward / backward at the shapes of the benched step (B = 128, H = 10): CUDA events, L2 flushed by size."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ickb200  # noqa
from ickb200.kernels import CudaKernels

K = CudaKernels()
B, H, dh = 128, 10, 30
bf = torch.bfloat16


def run(Sq, Sk, causal, p, iters=20):
    q = torch.randn(B * Sq, 320, device="cuda").to(bf)
    k, v = (torch.randn(B * Sk, 320, device="cuda").to(bf) for _ in range(2))
    o, do = torch.empty_like(q), torch.randn_like(q)
    lse, ds = torch.empty(B * H * Sq, device="cuda"), torch.empty(B * H * Sq, device="cuda")
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(k)
    drop = (p, 1, 2) if p > 0 else None
    res = []
    for name, fn in (("fwd", lambda: K.mha_fwd(q, k, v, o, lse, B, H, Sq, Sk, dh, causal, drop)),
                     ("bwd", lambda: K.mha_bwd(q, k, v, o, do, lse, ds, dq, dk, dv, B, H, Sq, Sk, dh, causal, drop))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(f"{name} {e0.elapsed_time(e1) / iters * 1000:.0f} us")
    scores = B * H * Sq * Sk * (0.5 if causal else 1.0)
    print(f"Sq={Sq} Sk={Sk} causal={causal} p={p}: " + ", ".join(res) + f"  ({scores / 1e6:.0f} M scores)")


for shape in ((301, 301, False), (102, 548, False), (102, 102, True), (51, 51, False)):
    for p in (0.0, 0.5, 0.3):
        run(*shape, p)
