"""Pins the CPU oracle (oracle/shf_oracle.c, a restatement of the reference filter) to the reference itself:
  1. the reference's own known-answer vector (STPTestHistogram.cpp:44-94),
  2. outputs of the reference's own compiled filter stored in tests/golden/ref_vectors.npz (made by make_golden.py),
  3. the reference build live, when oracle/_ref/libshf_ref.so is present,
  4. the independent closed-form order rule of SURVEY.md Appendix A.5,
  5. digests of the reference build's outputs on the problem sizes of the reference's own benchmark protocol
     (STPTestHistogram.cpp:215-295: dimension, radius and sample-range sweeps; tests/golden/protocol_digests.json).
No GPU involved."""
import json
import os

import numpy as np
import pytest

from golden import reference_vector as gv
from golden.cases import CASES, make_case
from golden.protocol_cases import NEIGHBOUR, PROTOCOL, digest, make_protocol_map
from helpers import assert_same

GOLDEN_NPZ = os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz")
PROTOCOL_DIGESTS = os.path.join(os.path.dirname(__file__), "golden", "protocol_digests.json")


def test_reference_known_answer_vector(oracle_mod):
    items, weights, offsets = oracle_mod.run_port(gv.TEXTURE, gv.DIMENSION, gv.NEIGHBOUR, gv.RADIUS)
    assert list(offsets) == list(range(0, 65, 4))
    for px, bins in gv.EXPECTED.items():
        lo, hi = int(offsets[px]), int(offsets[px + 1])
        assert list(items[lo:hi]) == [b[0] for b in bins]
        np.testing.assert_allclose(weights[lo:hi], [b[1] for b in bins], rtol=gv.WEIGHT_RTOL)
    flat = [b for px in gv.FULL_COUNTS for b in px]
    assert list(items) == [b[0] for b in flat]
    inv = np.float32(1.0) / np.float32(25.0)
    assert np.array_equal(weights.view(np.uint32), (np.array([b[1] for b in flat], np.float32) * inv).view(np.uint32))


@pytest.mark.parametrize("radius", gv.BAD_RADII)
def test_reference_error_cases(oracle_mod, radius):
    with pytest.raises(oracle_mod.OracleError) as err:
        oracle_mod.run_port(gv.TEXTURE, gv.DIMENSION, gv.NEIGHBOUR, radius)
    assert err.value.status == 1  # STPNumericDomainError
    with pytest.raises(oracle_mod.OracleError):
        oracle_mod.closed_form(gv.TEXTURE, gv.DIMENSION, gv.NEIGHBOUR, radius)


@pytest.mark.parametrize("index", range(len(CASES)))
def test_port_matches_stored_reference_outputs(oracle_mod, index):
    case = CASES[index]
    stored = np.load(GOLDEN_NPZ)
    want = (stored[f"items_{index}"], stored[f"weights_{index}"], stored[f"offsets_{index}"])
    got = oracle_mod.run_port(make_case(case), (case["w"], case["h"]), case["nn"], case["r"])
    assert_same(got, want, f"golden case {index} {case}")


@pytest.mark.parametrize("index", range(len(CASES)))
def test_port_matches_live_reference(oracle_mod, index):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libshf_ref.so not built here (needs /root/reference)")
    case = CASES[index]
    m = make_case(case)
    got = oracle_mod.run_port(m, (case["w"], case["h"]), case["nn"], case["r"])
    for exec_type in (0x00, 0xFF):
        want = oracle_mod.run_reference(m, (case["w"], case["h"]), case["nn"], case["r"], exec_type=exec_type)
        assert_same(got, want, f"live reference, exec {exec_type:#x}, case {index}")


@pytest.mark.parametrize("index", [0, 1, 2, 4, 5, 6, 7, 10])
def test_port_matches_closed_form(oracle_mod, index):
    case = CASES[index]
    m = make_case(case)
    got = oracle_mod.run_port(m, (case["w"], case["h"]), case["nn"], case["r"])
    want = oracle_mod.closed_form(m, (case["w"], case["h"]), case["nn"], case["r"])
    assert_same(got, want, f"closed form, case {index}")


def test_reference_session_invariants(oracle_mod):
    """type() echoes construction, a fresh buffer is empty, size() = (bins, W*H+1), reruns agree (TestHist:150-161)."""
    if not oracle_mod.have_reference():
        pytest.skip("reference build not present")
    for exec_type in (0x00, 0xFF):
        s = oracle_mod.ReferenceSession(exec_type)
        assert s.type() == exec_type
        assert s.size() == (0, 0, 3)
        a = s.run(gv.TEXTURE, gv.DIMENSION, gv.NEIGHBOUR, gv.RADIUS)
        b = s.run(gv.TEXTURE, gv.DIMENSION, gv.NEIGHBOUR, gv.RADIUS)
        assert_same(a, b, "rerun")
        assert s.size()[:2] == (64, 17)
        s.close()
    with pytest.raises(oracle_mod.OracleError) as err:
        oracle_mod.ReferenceSession(0x42)
    assert err.value.status == 2  # STPInvalidEnum


@pytest.mark.parametrize("index", range(len(PROTOCOL)))
def test_port_on_reference_benchmark_protocol(oracle_mod, index):
    """The shapes the reference benchmarks its filter on (16x16 .. 1024x1024 at radius 16; radius 2 .. 162 at 192x192;
    sample ranges 2 .. 30): the restatement against the stored digest of the reference build's output, and against the
    reference build live (both execution types) when it is present."""
    case = PROTOCOL[index]
    with open(PROTOCOL_DIGESTS) as f:
        stored = json.load(f)[index]
    assert stored["name"] == case["name"]
    m = make_protocol_map(index)
    dim = (case["dim"], case["dim"])
    got = oracle_mod.run_port(m, dim, NEIGHBOUR, case["r"])
    assert len(got[0]) == stored["bins"] and digest(got) == stored["sha256"], case["name"]
    if oracle_mod.have_reference():
        for exec_type in (0x00, 0xFF):
            assert_same(got, oracle_mod.run_reference(m, dim, NEIGHBOUR, case["r"], exec_type=exec_type),
                        f"live reference, exec {exec_type:#x}, {case['name']}")


def test_random_differential_restatement_vs_reference_build(oracle_mod):
    """tests/fuzz_cpu.py: 120 random cases (shapes, radii up to 300, up to 3000 distinct values, five map kinds), the
    restatement against the reference's own compiled filter, bit for bit."""
    import subprocess
    import sys

    if not oracle_mod.have_reference():
        pytest.skip("reference build not present")
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "fuzz_cpu.py"), "120", "20261017"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "port == reference on" in out.stdout, out.stdout[-500:] + out.stderr[-1500:]
