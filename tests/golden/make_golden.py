"""Generates tests/golden/ref_vectors.npz by running the REFERENCE's own STPSingleHistogramFilter (compiled from
/root/reference into oracle/_ref/libshf_ref.so by oracle/Makefile) on seeded inputs. Run in the build container, where
/root/reference exists; the committed .npz lets the oracle be pinned on machines without the reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from golden.cases import CASES, make_case  # noqa: E402


def main():
    oracle.build()
    assert oracle.have_reference(), "oracle/_ref/libshf_ref.so missing: /root/reference is needed to (re)generate"
    out = {}
    for i, case in enumerate(CASES):
        m = make_case(case)
        w, h, nn, r = case["w"], case["h"], case["nn"], case["r"]
        ser = oracle.run_reference(m, (w, h), nn, r, exec_type=0x00)
        par = oracle.run_reference(m, (w, h), nn, r, exec_type=0xFF)
        for a, b in zip(ser, par):  # the reference's two execution types agree bit for bit
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), case
        out[f"items_{i}"] = par[0]
        out[f"weights_{i}"] = par[1]
        out[f"offsets_{i}"] = par[2]
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(CASES), "cases")


if __name__ == "__main__":
    main()
