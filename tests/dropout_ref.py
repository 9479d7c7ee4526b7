"""Test-side port of the kernels' counter-based dropout hash (csrc/common.cuh: ick_hash / ick_drop_mul)."""
import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def ick_hash(seed: int, site: int, idx: np.ndarray) -> np.ndarray:
    idx = idx.astype(np.uint64)
    lo = idx & M32
    hi = (idx >> np.uint64(32)) & M32
    h = ((lo * np.uint64(0x9E3779B1)) & M32) ^ (((hi + np.uint64((site * 0x7F4A7C15) & 0xFFFFFFFF)) & M32) * np.uint64(0x85EBCA77) & M32) \
        ^ np.uint64(seed & 0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & M32
    h ^= h >> np.uint64(16)
    return h


def drop_mul(p: float, seed: int, site: int, numel: int) -> torch.Tensor:
    """multiplier (0 or 1/(1-p)) for flat element indices 0..numel-1, as float32"""
    if p <= 0.0:
        return torch.ones(numel, dtype=torch.float32)
    t = float(p) * 4294967296.0
    thr = 4294967295 if t >= 4294967295.0 else int(t)
    h = ick_hash(seed, site, np.arange(numel, dtype=np.uint64))
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(h >= np.uint64(thr), inv, np.float32(0.0)).astype(np.float32))
