#!/usr/bin/env python
"""Randomised differential test on a GPU: random shapes, radii (8-bit ring, 16-bit ring, wide path), biome counts
(K = 1..8 and beyond), map kinds and batch sizes against the CPU oracle, bit for bit. usage: python tests/fuzz_gpu.py [cases] [seed] [big]
(test infrastructure: run by tests/test_fuzz_gpu.py; the oracle is the checker)"""
import os, sys, time
_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
sys.path.insert(0, _HERE)
import numpy as np
import oracle
import superterrainplus_b200 as shf
from helpers import assert_same
from test_parity_gpu import random_map

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
big = len(sys.argv) > 3 and sys.argv[3] == "big"   # larger maps: the row / column segmented kernels of small calls
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
filt = shf.STPSingleHistogramFilter()
FB = shf.STPSingleHistogramFilter.STPFilterBuffer
buf = FB(0xFF)
kinds = ["iid", "blocky", "rare", "stripes", "hstripes"]
plans = {}
t0 = time.time()
for case in range(n_cases):
    w, h = (int(rng.integers(4, 200)), int(rng.integers(4, 160))) if not big else (int(rng.integers(64, 700)), int(rng.integers(64, 600)))
    rmax = int(rng.choice([16, 64, 126, 200, 254, 300]))
    r = 2 * int(rng.integers(1, rmax // 2 + 1))
    biomes = int(rng.choice([1, 2, 5, 20, 33, 64, 65, 100, 129, 200, 256, 257, 400]))
    kind = kinds[int(rng.integers(0, len(kinds)))]
    nn = (2 * ((r + w - 1) // w) + 1, 2 * ((r + h - 1) // h) + 1)
    if (w * nn[0]) * (h * nn[1]) > (1_500_000 if not big else 6_000_000):   # keep the oracle quick
        continue
    n = int(rng.choice([1, 1, 2, 3]))
    maps = [random_map(rng, w, h, biomes, kind, nn) for _ in range(n)]
    info = shf.STPNearestNeighbourInformation((w, h), nn, (w * nn[0], h * nn[1]))
    want = [oracle.run_port(m, (w, h), nn, r) for m in maps]
    # the first call on a shape takes the checked path, a repeated one runs ahead with the previous plan; a third call
    # with other maps of the same shape (possibly more distinct values or more bins than the buffers hold) must notice
    for rep, batch in enumerate((maps, maps, [random_map(rng, w, h, int(rng.choice([biomes, 2 * biomes + 3])), kind, nn)
                                              for _ in range(n)])):
        hist = filt.runBatch(batch, info, buf, r)
        base = buf.chunkBase()
        per = w * h + 1
        for i, m in enumerate(batch):
            got = (hist.Bin["Item"][base[i]:base[i + 1]].copy(), hist.Bin["Weight"][base[i]:base[i + 1]].copy(),
                   hist.HistogramStartOffset[i * per:(i + 1) * per].copy())
            ref = want[i] if rep < 2 else oracle.run_port(m, (w, h), nn, r)
            assert_same(got, ref, f"case {case} call {rep}: {w}x{h} r={r} B={biomes} {kind} nn={nn} chunk {i}/{n}")
    p = buf.lastPlan()
    key = (p["k_sets"], 16 if 2 * r + 1 > 255 and p["k_sets"] else 8 if p["k_sets"] else 0)
    plans[key] = plans.get(key, 0) + 1
print(f"fuzz ok: {n_cases} cases in {time.time() - t0:.1f} s; (k_sets, ring bits) -> cases: {dict(sorted(plans.items()))}")
