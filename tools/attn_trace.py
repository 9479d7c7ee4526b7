"""Timeline of CTA 0 of one tcgen05 attention-backward launch (debug build: ICK_EXTRA_NVCC_FLAGS=-DICK_TB_TRACE python -m ickb200.build --force).
    python tools/attn_trace.py [Sq Sk causal p]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ickb200  # noqa
from ickb200 import _lib
from ickb200.kernels import CudaKernels

K = CudaKernels()
lib = ctypes.CDLL(_lib.get().path)
Sq, Sk, causal, p = (int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3])), float(sys.argv[4])) if len(sys.argv) > 4 else (301, 301, False, 0.5)
B, H, dh, bf = 128, 10, 30, torch.bfloat16
q = torch.randn(B * Sq, 320, device="cuda").to(bf)
k, v = (torch.randn(B * Sk, 320, device="cuda").to(bf) for _ in range(2))
o, do = torch.randn_like(q), torch.randn_like(q)
lse, ds = torch.zeros(B * H * Sq, device="cuda"), torch.empty(B * H * Sq, device="cuda")
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(k)
for _ in range(3):
    K.mha_bwd(q, k, v, o, do, lse, ds, dq, dk, dv, B, H, Sq, Sk, dh, causal, (p, 1, 2) if p > 0 else None)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (6 * 4096))()
lib.ick_debug_tb_trace_read(buf)
ev = []
for role in range(4):
    for i in range(2047):
        tag, t = buf[role * 4096 + 2 * i], buf[role * 4096 + 2 * i + 1]
        if tag == 0xFFFFFFFFFFFFFFFF:
            break
        ev.append((t, role, tag >> 32, tag & 0xFFFFFFFF))
ev.sort()
t0 = ev[0][0]
names = {1: "SP: wait set free", 2: "SP: set free", 3: "SP: issued+committed", 4: "GR: wait tiles", 5: "GR: tiles written", 6: "GR: issued+committed",
         7: "WG: wait S/dP", 8: "WG: S/dP ready", 9: "WG: wait tile buffers", 10: "WG: buffers free", 11: "WG: rows written"}
limit = int(os.environ.get("TRACE_EVENTS", "260"))
for t, role, e, x in ev[:limit]:
    print(f"{(t - t0):8d} cyc  {' ' * 30 * role}{names.get(e, e)} t{x}")
