"""Test-side port of the kernels' counter-based dropout hash (csrc/common.cuh: ick_rowmix / ick_pairhash / ick_keep).

An element is addressed as (row, col): row = flat index over the leading dimensions, col = index in the last one.  One
32-bit hash serves the column pair (2k, 2k+1) as two 15-bit uniform fields (bits 0-14 / 16-30) compared against
thr = floor(p * 32768).

Attention-probability sites (site names ending in ".attn") use the bit-parallel definition instead (csrc/common.cuh:
ick_keepword): per probability row and group of 32 keys, up to 16 hashed words form 32 independent 16-bit uniform numbers (one
bit plane per word, most significant first); key k keeps its probability iff U < t16 = 65536 - 2 * thr, its number being the
one at bit (k >> 1) + 16 * (k & 1) of the words."""
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


def _fmix32(h: np.ndarray) -> np.ndarray:
    h = h & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & M32
    h ^= h >> np.uint64(16)
    return h


def attn_drop_mul(p: float, seed: int, site: int, rows: int, cols: int) -> torch.Tensor:
    """multiplier (0 or 1/(1-p)) for the (rows, cols) = (B*H*Sq, Sk) attention probabilities, as float32"""
    if p <= 0.0:
        return torch.ones(rows, cols, dtype=torch.float32)
    t = float(p) * 32768.0
    thr = 32767 if t >= 32767.0 else int(t)
    t16 = 65536 - 2 * thr
    rm = ick_rowmix(seed, site, np.arange(rows, dtype=np.uint64))[:, None]
    ngroups = (cols + 31) // 32
    kg = np.arange(ngroups, dtype=np.uint64)[None, :]
    lt = np.zeros((rows, ngroups), dtype=np.uint64)
    eq = np.full((rows, ngroups), 0xFFFFFFFF, dtype=np.uint64)
    for i in range(16):
        if t16 & (0xFFFF >> i) == 0:
            break
        w = _fmix32(rm + ((kg * np.uint64(16) + np.uint64(i)) * np.uint64(0x9E3779B1) & M32))
        if (t16 >> (15 - i)) & 1:
            lt |= eq & (~w & M32)
            eq &= w
        else:
            eq &= ~w & M32
    col = np.arange(cols, dtype=np.uint64)
    bit = ((col & np.uint64(31)) >> np.uint64(1)) + ((col & np.uint64(1)) << np.uint64(4))
    keep = (lt[:, (col >> np.uint64(5)).astype(np.int64)] >> bit[None, :]) & np.uint64(1)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(keep != 0, inv, np.float32(0.0)).astype(np.float32))
