"""Small driver for ncu captures of the bf16 attention kernels at the entity-encoder shape (B=128, H=10, S=301)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ickb200  # noqa
from ickb200.kernels import CudaKernels

K = CudaKernels()
B, H, S, dh = 128, 10, 301, 30
bf = torch.bfloat16
q, k, v = (torch.randn(B * S, 320, device="cuda").to(bf) for _ in range(3))
o, do = torch.empty_like(q), torch.randn_like(q)
lse, ds = torch.empty(B * H * S, device="cuda"), torch.empty(B * H * S, device="cuda")
dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
for _ in range(3):
    K.mha_fwd(q, k, v, o, lse, B, H, S, S, dh, False, (0.5, 1, 2))
    K.mha_bwd(q, k, v, o, do, lse, ds, dq, dk, dv, B, H, S, S, dh, False, (0.5, 1, 2))
torch.cuda.synchronize()
print("ok")
