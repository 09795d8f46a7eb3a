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
