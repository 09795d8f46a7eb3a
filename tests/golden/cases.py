"""Seeded input cases shared by the golden-vector generator and the tests (inputs are regenerated, outputs are stored)."""
import numpy as np

# w, h: chunk map size; nn: neighbourhood; r: radius; b: biome count; kind: map structure
CASES = [
    dict(w=4, h=4, nn=(3, 3), r=2, b=4, kind="iid", seed=1),
    dict(w=16, h=12, nn=(3, 3), r=4, b=6, kind="iid", seed=2),
    dict(w=12, h=20, nn=(3, 3), r=12, b=9, kind="blocky", seed=3),
    dict(w=32, h=32, nn=(3, 3), r=16, b=40, kind="iid", seed=4),
    dict(w=24, h=24, nn=(3, 3), r=24, b=3, kind="rare", seed=5),
    dict(w=40, h=16, nn=(3, 3), r=6, b=12, kind="stripes", seed=6),
    dict(w=16, h=40, nn=(3, 3), r=10, b=12, kind="hstripes", seed=7),
    dict(w=20, h=20, nn=(5, 3), r=20, b=15, kind="blocky", seed=8),
    dict(w=64, h=48, nn=(3, 3), r=32, b=200, kind="iid", seed=9),
    dict(w=48, h=64, nn=(3, 3), r=8, b=1000, kind="iid", seed=10),
    dict(w=33, h=17, nn=(3, 3), r=2, b=2, kind="blocky", seed=11),
    dict(w=30, h=30, nn=(3, 3), r=14, b=7, kind="sparse_ids", seed=12),
    dict(w=64, h=64, nn=(3, 3), r=64, b=64, kind="iid", seed=13),
    dict(w=96, h=40, nn=(3, 3), r=20, b=10, kind="blocky", seed=14),
    # radius above 126 (window counts need more than 8 bits) and the reference benchmark's radius 162
    dict(w=20, h=16, nn=(15, 17), r=128, b=40, kind="iid", seed=15),
    dict(w=24, h=18, nn=(15, 19), r=162, b=5, kind="blocky", seed=16),
    dict(w=20, h=16, nn=(15, 17), r=128, b=600, kind="rare", seed=17),
]


def make_case(case):
    rng = np.random.default_rng(case["seed"])
    tw, th = case["w"] * case["nn"][0], case["h"] * case["nn"][1]
    b, kind = case["b"], case["kind"]
    if kind == "iid":
        m = rng.integers(0, b, (th, tw))
    elif kind == "blocky":
        bs = int(rng.integers(2, 9))
        m = rng.integers(0, b, (th // bs + 1, tw // bs + 1)).repeat(bs, 0).repeat(bs, 1)[:th, :tw]
    elif kind == "rare":
        m = np.where(rng.random((th, tw)) < 0.95, 0, rng.integers(0, b, (th, tw)))
    elif kind == "stripes":
        m = np.broadcast_to((np.arange(tw) % int(rng.integers(2, 12)))[None, :] % b, (th, tw))
    elif kind == "hstripes":
        m = np.broadcast_to((np.arange(th) % int(rng.integers(2, 12)))[:, None] % b, (th, tw))
    elif kind == "sparse_ids":
        ids = np.array([0, 7, 300, 4095, 65535, 12345, 32768], dtype=np.int64)
        m = ids[rng.integers(0, len(ids), (th, tw))]
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(m, dtype=np.uint16)
