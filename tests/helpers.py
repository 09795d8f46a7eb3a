"""Shared comparison helpers for the parity tests."""
from __future__ import annotations

import numpy as np


def split_result(hist):
    """STPSingleHistogram view -> (items u16, weights f32, offsets u32) copies."""
    return hist.Bin["Item"].copy(), hist.Bin["Weight"].copy(), hist.HistogramStartOffset.copy()


def assert_same(got, want, what=""):
    """Bit-exact comparison of (items, weights, offsets); on failure names the first differing pixel."""
    gi, gw, go = got
    wi, ww, wo = want
    assert go.shape == wo.shape, f"{what}: offsets shape {go.shape} vs {wo.shape}"
    if not np.array_equal(go, wo):
        p = int(np.flatnonzero(go != wo)[0])
        raise AssertionError(f"{what}: offsets differ first at index {p}: got {go[max(0, p - 2):p + 3]} want {wo[max(0, p - 2):p + 3]}")
    assert gi.shape == wi.shape, f"{what}: bins {gi.shape} vs {wi.shape}"
    if not np.array_equal(gi, wi):
        b = int(np.flatnonzero(gi != wi)[0])
        px = int(np.searchsorted(wo, b, side="right") - 1)
        lo, hi = int(wo[px]), int(wo[px + 1])
        raise AssertionError(f"{what}: items differ at bin {b} (pixel {px}): got {gi[lo:hi]} want {wi[lo:hi]}")
    gb, wb = gw.view(np.uint32), ww.view(np.uint32)
    if not np.array_equal(gb, wb):
        b = int(np.flatnonzero(gb != wb)[0])
        px = int(np.searchsorted(wo, b, side="right") - 1)
        raise AssertionError(f"{what}: weight bits differ at bin {b} (pixel {px}): got {gw[b]!r} want {ww[b]!r}")


def make_generator_tables(seed, n_table, grad_size=12, max_octave=8):
    """Inputs of the heightfield consumer: biome property table, 512-byte permutation, unit-gradient table. The reference
    derives the last two from std::mt19937_64 + std::shuffle (STPPermutationGenerator.cpp:40-93), which is standard-
    library defined, so they are treated as inputs here: any permutation of 0..255 repeated twice, gradients evenly
    spread on the unit circle."""
    rng = np.random.default_rng(seed)
    table = np.zeros(n_table, dtype=[("Scale", "<f4"), ("Octave", "<u4"), ("Persistence", "<f4"), ("Lacunarity", "<f4"),
                                     ("Depth", "<f4"), ("Variation", "<f4")])
    table["Scale"] = rng.uniform(40.0, 900.0, n_table)
    table["Octave"] = rng.integers(1, max_octave + 1, n_table)
    table["Persistence"] = rng.uniform(0.3, 0.7, n_table)
    table["Lacunarity"] = rng.uniform(1.7, 2.6, n_table)
    table["Depth"] = rng.uniform(0.0, 1.0, n_table)
    table["Variation"] = rng.uniform(0.05, 1.0, n_table)
    perm = np.tile(rng.permutation(256).astype(np.uint8), 2)
    angle = np.arange(grad_size, dtype=np.float64) * (2.0 * np.pi / grad_size) + float(rng.uniform(0, 1))
    grad = np.stack([np.cos(angle), np.sin(angle)], axis=1).astype(np.float32)
    return table, perm, grad
