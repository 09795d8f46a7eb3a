"""Deterministic synthetic biome maps for the parity tests and the benchmark.

Counter-based, so numpy (host) and torch (device) produce the same map without shared RNG state:
    id(chunk, y, x) = (splitmix64(seed ^ chunk<<40 ^ cy<<20 ^ cx) >> 33) mod B
with (cy, cx) = (y, x) for the "uniform" distribution (iid ids, the reference benchmark's protocol,
SuperTest+/SuperAlgorithm+/STPTestHistogram.cpp:342-343), (y//64, x//64) for "blocky" (clustered biomes),
"rare" = id 0 with p ~ 0.98 else uniform, "stripes" = vertical stripes of width 1 with period 2r+3 (bin birth/death
stress for the ordering rule).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

SEED = 0x5EED0001
_M64 = (1 << 64) - 1


@dataclass(frozen=True)
class Workload:
    name: str
    map_size: Tuple[int, int]      # W, H of one chunk
    nn: Tuple[int, int]            # chunk neighbourhood
    radius: int
    biomes: int
    chunks: int
    dist: str = "uniform"

    @property
    def total(self) -> Tuple[int, int]:
        return self.map_size[0] * self.nn[0], self.map_size[1] * self.nn[1]

    @property
    def pixels(self) -> int:
        return self.map_size[0] * self.map_size[1] * self.chunks


# BASELINE.json configs
CONFIGS = {
    "C1": Workload("C1", (512, 512), (3, 3), 32, 8, 1),
    "C2": Workload("C2", (1024, 1024), (3, 3), 64, 32, 1),
    "C3": Workload("C3", (512, 512), (3, 3), 64, 64, 256),
    "C4": Workload("C4", (2048, 2048), (3, 3), 64, 32, 1),
}


def _splitmix64_np(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _cell_keys_np(h: int, w: int, dist: str, radius: int):
    y = np.arange(h, dtype=np.uint64)[:, None]
    x = np.arange(w, dtype=np.uint64)[None, :]
    if dist == "blocky":
        return y // np.uint64(64), x // np.uint64(64)
    if dist == "stripes":
        return np.zeros_like(y), x % np.uint64(2 * radius + 3)
    return y, x


def make_map_np(wl: Workload, chunk: int = 0, seed: int = SEED) -> np.ndarray:
    """One merged neighbourhood map, uint16 [TotalMapSize.y, TotalMapSize.x]."""
    tw, th = wl.total
    with np.errstate(over="ignore"):
        cy, cx = _cell_keys_np(th, tw, wl.dist, wl.radius)
        key = np.uint64(seed & _M64) ^ (np.uint64(chunk) << np.uint64(40)) ^ (cy << np.uint64(20)) ^ cx
        v = _splitmix64_np(np.broadcast_to(key, (th, tw)).copy()) >> np.uint64(33)
        ids = v % np.uint64(wl.biomes)
        if wl.dist == "rare":
            y = np.arange(th, dtype=np.uint64)[:, None]
            x = np.arange(tw, dtype=np.uint64)[None, :]
            k2 = np.uint64((seed ^ 0xABCDEF) & _M64) ^ (np.uint64(chunk) << np.uint64(40)) ^ (y << np.uint64(20)) ^ x
            u = _splitmix64_np(np.broadcast_to(k2, (th, tw)).copy()) >> np.uint64(33)
            ids = np.where(u % np.uint64(100) < np.uint64(98), np.uint64(0), ids)
    return ids.astype(np.uint16)


def make_maps_torch(wl: Workload, first_chunk: int, n_chunks: int, device, seed: int = SEED):
    """n_chunks merged maps as one uint16 tensor [n, TotalMapSize.y, TotalMapSize.x] generated on `device`;
    bit-identical to make_map_np for every chunk index."""
    import torch

    tw, th = wl.total

    def i64(v: int) -> int:  # python int -> two's complement int64
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(t, k: int):  # logical shift right on int64
        return (t >> k) & ((1 << (64 - k)) - 1)

    def splitmix(z):
        z = z + i64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * i64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * i64(0x94D049BB133111EB)
        return z ^ lsr(z, 31)

    y = torch.arange(th, dtype=torch.int64, device=device)[:, None]
    x = torch.arange(tw, dtype=torch.int64, device=device)[None, :]
    if wl.dist == "blocky":
        cy, cx = y // 64, x // 64
    elif wl.dist == "stripes":
        cy, cx = torch.zeros_like(y), x % (2 * wl.radius + 3)
    else:
        cy, cx = y, x
    cell = (cy << 20) ^ cx
    out = torch.empty((n_chunks, th, tw), dtype=torch.uint16, device=device)
    for i in range(n_chunks):
        chunk = first_chunk + i
        key = cell ^ i64(seed ^ (chunk << 40))
        ids = lsr(splitmix(key.expand(th, tw)), 33) % wl.biomes
        if wl.dist == "rare":
            k2 = ((y << 20) ^ x) ^ i64((seed ^ 0xABCDEF) ^ (chunk << 40))
            u = lsr(splitmix(k2), 33) % 100
            ids = torch.where(u < 98, torch.zeros_like(ids), ids)
        out[i] = ids.to(torch.uint16)
    return out


def algorithmic_bytes(wl: Workload, n_bins: int) -> int:
    """B_alg of BASELINE.md section 4 summed over the batch: halo-extended u16 input read once, u32 offsets and
    8-byte bins written once."""
    w, h = wl.map_size
    r = wl.radius
    per_chunk = 2 * (w + 2 * r) * (h + 2 * r) + 4 * (w * h + 1)
    return per_chunk * wl.chunks + 8 * n_bins
