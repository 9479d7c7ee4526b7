"""Per-shape timing of the hot kernels with CUDA events (run on the GPU box): python tools/microbench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ickb200  # noqa
from ickb200.kernels import CudaKernels

K = CudaKernels()
dev = "cuda"
bf = torch.bfloat16


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


B = 128
shapes = {
    "ent qkv": (B * 301, 960, 320), "ent out": (B * 301, 320, 320), "ent ffn1": (B * 301, 512, 320), "ent ffn2": (B * 301, 320, 512),
    "dec qkv": (B * 102, 960, 320), "dec out": (B * 102, 320, 320), "dec ffn1": (B * 102, 512, 320),
    "mem kv_all": (B * 548, 1920, 320), "mem dkv dgrad": (B * 548, 320, 1920), "vocab fwd(f32 out)": (B * 102, 10000, 320),
    "vocab dgrad": (B * 102, 320, 10000), "fact qkv": (B * 51, 960, 320),
}
print("== gemm_tn_tc")
for name, (M, N, Kd) in shapes.items():
    A = torch.randn(M, Kd, device=dev).to(bf)
    W = (torch.randn(N, Kd, device=dev) * 0.05).to(bf)
    C = torch.empty(M, N, device=dev, dtype=torch.float32 if "f32" in name else bf)
    bias = torch.randn(N, device=dev)
    ms = timeit(lambda: K.gemm(A, W, C, bias=bias))
    by = (M * Kd + N * Kd) * 2 + M * N * C.element_size()
    print(f"{name:22s} M={M:6d} N={N:5d} K={Kd:5d}  {ms*1e3:8.1f} us  {2*M*N*Kd/ms/1e9:8.1f} TFLOP/s  {by/ms/1e6:8.1f} GB/s")
print("== wgrad_tc")
for name, (M, N, Kd) in shapes.items():
    if "dgrad" in name:
        continue
    dY = torch.randn(M, N, device=dev).to(bf)
    X = torch.randn(M, Kd, device=dev).to(bf)
    G = torch.zeros(N * Kd + N, device=dev)
    rowoff = (torch.arange(N, device=dev) * Kd).int()
    colmap = torch.arange(Kd, device=dev).int()
    biasoff = (N * Kd + torch.arange(N, device=dev)).int()
    ms = timeit(lambda: K.wgrad(dY, X, G, rowoff, colmap, biasoff))
    by = (M * Kd + M * N) * 2
    print(f"{name:22s} M={M:6d} N={N:5d} K={Kd:5d}  {ms*1e3:8.1f} us  {2*M*N*Kd/ms/1e9:8.1f} TFLOP/s  {by/ms/1e6:8.1f} GB/s")
print("== attention (bf16 mma)")
for name, (Sq, Sk, causal) in {"entity self": (301, 301, False), "cross": (102, 548, False), "dec self causal": (102, 102, True), "fact self": (51, 51, False)}.items():
    H, dh = 10, 30
    q = torch.randn(B * Sq, 320, device=dev).to(bf)
    k = torch.randn(B * Sk, 320, device=dev).to(bf)
    v = torch.randn(B * Sk, 320, device=dev).to(bf)
    o = torch.empty_like(q)
    lse = torch.empty(B * H * Sq, device=dev)
    for p in (0.0, 0.5):
        drop = (p, 1, 2) if p else None
        ms = timeit(lambda: K.mha_fwd(q, k, v, o, lse, B, H, Sq, Sk, dh, causal, drop))
        fl = 4 * B * H * Sq * Sk * 32 * (0.5 if causal else 1)
        print(f"fwd {name:16s} p={p}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
        do = torch.randn_like(q)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ds = torch.empty(B * H * Sq, device=dev)
        ms = timeit(lambda: K.mha_bwd(q, k, v, o, do, lse, ds, dq, dk, dv, B, H, Sq, Sk, dh, causal, drop))
        print(f"bwd {name:16s} p={p}  {ms*1e3:8.1f} us  {2.5*fl/ms/1e9:7.1f} TFLOP/s")

print("== encoder hand-off (pool 8x8 -> 14x14 rows, 1x1 conv 2048 -> 300 as a GEMM, transpose to (B,300,196))")
feats = torch.randn(B, 2048, 8, 8, device=dev)
rows = torch.empty(B * 196, 2048, device=dev, dtype=bf)
Wc = (torch.randn(300, 2048, device=dev) * 0.02).to(bf)
bc = torch.randn(300, device=dev)
y = torch.zeros(B * 196, 304, device=dev, dtype=bf)
out = torch.empty(B, 300, 196, device=dev)
t_pool = timeit(lambda: K.pool_rows_fwd(feats, rows, B, 2048, 8, 8, 14, 14))
t_gemm = timeit(lambda: K.gemm(rows, Wc, y[:, :300], bias=bc))
t_tr = timeit(lambda: K.pixels_bwd(y, out, B, 300, 196, 196))
fl = 2 * B * 196 * 2048 * 300
print(f"pool_rows {t_pool*1e3:6.1f} us ({(feats.numel()*4 + rows.numel()*2)/t_pool/1e6:6.0f} GB/s)   conv1 gemm {t_gemm*1e3:6.1f} us "
      f"({fl/t_gemm/1e9:6.1f} TFLOP/s)   transpose {t_tr*1e3:6.1f} us   total {(t_pool+t_gemm+t_tr)*1e3:6.1f} us for {B} images")
conv = torch.nn.Conv2d(2048, 300, 1).to(dev)
pool = torch.nn.AdaptiveAvgPool2d((14, 14))
with torch.no_grad():
    t_ref = timeit(lambda: conv(pool(feats)).view(B, 300, -1))
print(f"stock PyTorch (AdaptiveAvgPool2d + cuDNN Conv2d, TF32) {t_ref*1e3:6.1f} us")
