"""Test-side port of the kernels' counter-based dropout hash (csrc/common.cuh: ick_rowmix / ick_pairhash / ick_keep).

An element is addressed as (row, col): row = flat index over the leading dimensions, col = index in the last one.  One
32-bit hash serves the column pair (2k, 2k+1) as two 15-bit uniform fields (bits 0-14 / 16-30) compared against
thr = floor(p * 32768)."""
import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def _u(x):
    return np.uint64(x & 0xFFFFFFFF)


def ick_rowmix(seed: int, site: int, row: np.ndarray) -> np.ndarray:
    row = row.astype(np.uint64)
    lo = row & M32
    hi = (row >> np.uint64(32)) & M32
    h = _u(seed) ^ _u(site * 0x7F4A7C15) ^ ((lo * np.uint64(0x9E3779B1)) & M32) ^ ((hi * np.uint64(0x85EBCA77)) & M32)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & M32
    h ^= h >> np.uint64(16)
    return h


def ick_pairhash(rowmix: np.ndarray, col: np.ndarray) -> np.ndarray:
    h = (rowmix + (col.astype(np.uint64) >> np.uint64(1)) * np.uint64(0x9E3779B1)) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M32
    return h


def drop_mul(p: float, seed: int, site: int, rows: int, cols: int) -> torch.Tensor:
    """multiplier (0 or 1/(1-p)) for a (rows, cols) tensor, as float32"""
    if p <= 0.0:
        return torch.ones(rows, cols, dtype=torch.float32)
    t = float(p) * 32768.0
    thr = 32767 if t >= 32767.0 else int(t)
    rm = ick_rowmix(seed, site, np.arange(rows, dtype=np.uint64))[:, None]
    col = np.arange(cols, dtype=np.uint64)[None, :]
    h = ick_pairhash(rm, col)
    f = np.where((col & np.uint64(1)) != 0, h >> np.uint64(16), h) & np.uint64(0x7FFF)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(f >= np.uint64(thr), inv, np.float32(0.0)).astype(np.float32))
