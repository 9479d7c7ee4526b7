"""Timeline of CTA 0 of one tcgen05 GEMM launch (debug build: ICK_EXTRA_NVCC_FLAGS=-DICK_TRACE python -m ickb200.build --force)."""
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
M, N, Kd = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (38528, 320, 320)))
bf = torch.bfloat16
A = torch.randn(M, Kd, device="cuda").to(bf)
W = (torch.randn(N, Kd, device="cuda") * 0.05).to(bf)
C = torch.empty(M, N, device="cuda", dtype=bf)
bias = None if "nobias" in sys.argv else torch.randn(N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
NOFLUSH = "noflush" in sys.argv  # keep L2 warm (operands and bias resident, as inside the step)
buf = (ctypes.c_ulonglong * 8192)()
zero = torch.zeros(1)
for it in range(4):
    if not NOFLUSH:
        flush.zero_()
    torch.cuda.synchronize()
    K.gemm(A, W, C, bias=bias)
    torch.cuda.synchronize()
lib.ick_debug_trace_read(buf, 4096)
ev = []
for role in range(3):
    for i in range(1023):
        tag, t = buf[role * 2048 + 2 * i], buf[role * 2048 + 2 * i + 1]
        if tag == 0xFFFFFFFFFFFFFFFF:
            break
        ev.append((t, tag >> 32, tag & 0xFFFFFFFF))
ev.sort()
t0 = ev[0][0]
names = {1: "P start", 2: "P: stage free -> issue kb", 3: "M: tile start (tmem free)", 4: "M: operands landed kb", 5: "E: wait acc",
         6: "E: acc ready", 7: "E: chunk in regs", 8: "E: chunk store issued", 9: "E: done"}
col = {1: 0, 2: 0, 3: 1, 4: 1, 5: 2, 6: 2, 7: 2, 8: 2, 9: 2}
for t, e, x in ev:
    print(f"{(t - t0):8d} cyc  {' ' * 34 * col.get(e, 0)}{names.get(e, e)} {x}")
