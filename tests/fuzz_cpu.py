#!/usr/bin/env python
"""Randomised differential run on the CPU: the C restatement (oracle/shf_oracle.c) against the reference's own compiled filter
(oracle/_ref/libshf_ref.so; needs /root/reference to have been present at build time), bit for bit, on random shapes, radii,
biome counts and map kinds -- the same case generator as tests/fuzz_gpu.py. usage: python tests/fuzz_cpu.py [cases] [seed]
(test infrastructure; 1500 cases take ~100 s)"""
import os, sys, time
_HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(_HERE)); sys.path.insert(0, _HERE)
import numpy as np
import oracle
from helpers import assert_same
from test_parity_gpu import random_map
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 777)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kinds = ["iid", "blocky", "rare", "stripes", "hstripes"]
t0=time.time(); done=0
for case in range(n_cases):
    w, h = int(rng.integers(4, 200)), int(rng.integers(4, 160))
    rmax = int(rng.choice([16, 64, 126, 200, 254, 300]))
    r = 2 * int(rng.integers(1, rmax // 2 + 1))
    biomes = int(rng.choice([1, 2, 5, 20, 33, 64, 65, 100, 129, 200, 256, 257, 400, 3000]))
    kind = kinds[int(rng.integers(0, len(kinds)))]
    nn = (2 * ((r + w - 1) // w) + 1, 2 * ((r + h - 1) // h) + 1)
    if (w * nn[0]) * (h * nn[1]) > 1_500_000: continue
    m = random_map(rng, w, h, biomes, kind, nn)
    a = oracle.run_port(m, (w, h), nn, r)
    b = oracle.run_reference(m, (w, h), nn, r)
    assert_same(a, b, f"case {case}: {w}x{h} r={r} B={biomes} {kind} nn={nn}")
    done+=1
print(f"port == reference on {done} random cases in {time.time()-t0:.0f} s")
