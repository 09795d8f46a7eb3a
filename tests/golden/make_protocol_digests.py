"""Generates tests/golden/protocol_digests.json: SHA-256 digests of the outputs of the REFERENCE's own compiled filter
(oracle/_ref/libshf_ref.so, built from /root/reference by oracle/Makefile) on the problem sizes of the reference's
benchmark protocol (tests/golden/protocol_cases.py). Run in the build container, where /root/reference exists.

    python tests/golden/make_protocol_digests.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from golden.protocol_cases import NEIGHBOUR, PROTOCOL, digest, make_protocol_map  # noqa: E402


def main():
    oracle.build()
    assert oracle.have_reference(), "oracle/_ref/libshf_ref.so missing: /root/reference is needed to (re)generate"
    out = []
    for i, case in enumerate(PROTOCOL):
        m = make_protocol_map(i)
        dim = (case["dim"], case["dim"])
        ser = oracle.run_reference(m, dim, NEIGHBOUR, case["r"], exec_type=0x00)
        par = oracle.run_reference(m, dim, NEIGHBOUR, case["r"], exec_type=0xFF)
        for a, b in zip(ser, par):  # the reference's two execution types agree bit for bit
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), case
        out.append({"name": case["name"], "bins": int(len(par[0])), "sha256": digest(par)})
        print(out[-1])
    path = os.path.join(ROOT, "tests", "golden", "protocol_digests.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path)


if __name__ == "__main__":
    main()
