"""GEMM + add_ln_fwd against the fused ick_gemm_add_ln_tc at the train step's shapes (CUDA events, back to back, rotating buffers):
python tools/fuse_ln_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ickb200  # noqa: F401,E402
from ickb200.kernels import CudaKernels  # noqa: E402

K = CudaKernels()
dev = torch.device("cuda", 0)
d, ld = 300, 320


def run(M, Kd, dual, fused, reps=30, nbuf=6):
    os.environ["ICK_FUSE_LN"] = "1" if fused else "0"
    g = torch.Generator(device="cpu").manual_seed(0)
    rows0 = 38528 if dual else M
    m_split = (rows0 + 127) // 128 * 128
    Mt = m_split + 6528 if dual else M
    bufs = []
    for _ in range(nbuf):
        A = (torch.randn(Mt, Kd, generator=g) * 0.5).to(torch.bfloat16).to(dev)
        x = torch.randn(Mt, ld, generator=g).to(torch.bfloat16).to(dev)
        s = torch.empty(Mt, ld, dtype=torch.bfloat16, device=dev)
        y = torch.empty(Mt, ld, dtype=torch.bfloat16, device=dev)
        bufs.append((A, x, s, y))
    W = [(torch.randn(ld, Kd, generator=g) * 0.05).to(torch.bfloat16).to(dev) for _ in range(2)]
    b = [torch.zeros(ld, device=dev) for _ in range(2)]
    gam = [torch.ones(d, device=dev) for _ in range(2)]
    bet = [torch.zeros(d, device=dev) for _ in range(2)]
    mean, rstd = torch.empty(Mt, device=dev), torch.empty(Mt, device=dev)
    drops = ((0.5, 1234, 3), (0.5, 1234, 4))

    def once(i):
        A, x, s, y = bufs[i % nbuf]
        if dual:
            K.gemm_add_ln_dual(A, W[0], W[1], b[0], b[1], x, s, y, mean, rstd, d, m_split, rows0, gam, bet, drops)
        else:
            K.gemm_add_ln(A, W[0], b[0], x, s, gam[0], bet[0], y, mean, rstd, d, drop=drops[0])

    for i in range(3):
        once(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for i in range(reps):
        once(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for M, Kd, dual in ((13056, 320, False), (13056, 512, False), (45056, 320, True), (45056, 512, True), (38528, 320, False), (625, 320, False)):
    a, f = run(M, Kd, dual, False), run(M, Kd, dual, True)
    print(f"M={M} K={Kd} dual={dual}: GEMM + LayerNorm {a:.1f} us, fused {f:.1f} us")
