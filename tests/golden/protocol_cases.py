"""The problem sizes of the reference's own benchmark of the filter (SuperTest+/SuperAlgorithm+/STPTestHistogram.cpp:
215-226 settings, :251-264 dimension sweep, :266-279 radius sweep, :281-295 sample-range sweep): 3x3 neighbourhoods of
iid random maps `bounded(rangeMax)`. The reference draws them from nanobench's Rng; here a seeded numpy generator stands
in (the protocol's shapes are what is pinned, outputs of the reference build on these inputs are stored as digests)."""
import hashlib

import numpy as np

NEIGHBOUR = (3, 3)
DEFAULT_DIM, DEFAULT_RADIUS, DEFAULT_RANGE = 192, 16, 5

PROTOCOL = (
    [dict(name=f"dimension {16 << i}x{16 << i}", dim=16 << i, r=DEFAULT_RADIUS, rng_max=DEFAULT_RANGE) for i in range(7)]
    + [dict(name=f"radius {2 * 3 ** i}", dim=DEFAULT_DIM, r=2 * 3 ** i, rng_max=DEFAULT_RANGE) for i in range(5)]
    + [dict(name=f"range [0, {m}]", dim=DEFAULT_DIM, r=DEFAULT_RADIUS, rng_max=m) for m in (2, 8, 15, 30)]
)


def make_protocol_map(index):
    case = PROTOCOL[index]
    rng = np.random.default_rng(0xBE7C + index)
    side = case["dim"] * NEIGHBOUR[0]
    return np.ascontiguousarray(rng.integers(0, case["rng_max"], (side, side)), dtype=np.uint16)


def digest(result):
    """One SHA-256 over items, weight bits and offsets of a filter result."""
    items, weights, offsets = result
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(items, dtype=np.uint16).tobytes())
    h.update(np.ascontiguousarray(weights, dtype=np.float32).view(np.uint32).tobytes())
    h.update(np.ascontiguousarray(offsets, dtype=np.uint32).tobytes())
    return h.hexdigest()
