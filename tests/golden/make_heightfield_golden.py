"""Generates heightfield_ref.npz: outputs of the REFERENCE's own device code (STPSimplexNoise.cu + the histogram wrapper,
compiled from /root/reference into oracle/_ref/libshf_ref_height.so) on seeded inputs. Needs a GPU, so it runs on the GPU
box (gpurun) and writes into gpurun_out/; the file is then committed under tests/golden/.

    python tests/golden/make_heightfield_golden.py gpurun_out/heightfield_ref.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from golden.heightfield_cases import HEIGHT_CASES, make_height_case  # noqa: E402


def main():
    out = {}
    worst = 0.0
    for i, case in enumerate(HEIGHT_CASES):
        m, table, perm, grad = make_height_case(case)
        size = (case["w"], case["h"])
        items, weights, offsets = oracle.run_port(m, size, (3, 3), case["r"])
        ref = oracle.heightfield_reference(items, weights, offsets, size, table, perm, grad, case["offset"])
        port = oracle.heightfield_port(items, weights, offsets, size, table, perm, grad, case["offset"])
        worst = max(worst, float(np.abs(ref - port).max()))
        out[f"height_{i}"] = ref
    np.savez_compressed(sys.argv[1], **out)
    print(sys.argv[1], len(HEIGHT_CASES), "cases; max |reference - C restatement| =", worst)


if __name__ == "__main__":
    main()
