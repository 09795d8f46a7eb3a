"""Generates tests/golden/biome_vectors.npz: outputs of the REFERENCE's own biome-map producer (STPLayerChainBuilder of
SuperDemo+/World/Layers/STPAllLayers.cpp, compiled from /root/reference into oracle/_ref/libbiome_ref.so by
oracle/Makefile) for the cases listed in BIOME_CASES. Run in the build container (where /root/reference exists):

    python tests/golden/make_biome_golden.py

The .npz travels to the GPU box, /root/reference does not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# (global seed, (offset x, offset z), width, height)
BIOME_CASES = [
    (6666, (-256, 1024), 512, 512),
    (1, (0, 0), 300, 200),
    (0xDEADBEEFCAFE, (-100000, 77777), 257, 129),
    (42, (5000, -9000), 64, 700),
    (2 ** 63 + 12345, (-3, -5), 33, 17),
    (987654321, (123456, 654321), 640, 640),
    (7, (-1536, -1536), 1536, 1536),
    (20261017, (1 << 24, -(1 << 23)), 96, 1200),
]


def main():
    import oracle
    from oracle import biome

    oracle.build()
    out = {}
    for i, (seed, off, w, h) in enumerate(BIOME_CASES):
        out[f"map_{i}"] = biome.biome_reference(w, h, off, seed)
        print(i, seed, off, w, h, np.unique(out[f"map_{i}"], return_counts=True))
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "biome_vectors.npz"), **out)


if __name__ == "__main__":
    main()
