"""Pins the CPU restatement of the heightfield consumer (oracle/shf_heightfield_oracle.c) without a GPU:
against outputs of the reference's own device code stored in tests/golden/heightfield_ref.npz (made on a GPU box by
tests/golden/make_heightfield_golden.py), and through properties of the formula."""
import os

import numpy as np
import pytest

from golden.heightfield_cases import HEIGHT_CASES, make_height_case
from helpers import make_generator_tables

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "heightfield_ref.npz")
# the reference leaves multiply-add contraction to NVRTC; the restatement fixes one rounding sequence
TOL_REFERENCE = 2e-5


@pytest.mark.parametrize("index", range(len(HEIGHT_CASES)))
def test_port_matches_stored_reference_device_output(oracle_mod, index):
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/heightfield_ref.npz not generated yet (needs a GPU box)")
    case = HEIGHT_CASES[index]
    m, table, perm, grad = make_height_case(case)
    size = (case["w"], case["h"])
    items, weights, offsets = oracle_mod.run_port(m, size, (3, 3), case["r"])
    got = oracle_mod.heightfield_port(items, weights, offsets, size, table, perm, grad, case["offset"])
    want = np.load(GOLDEN)[f"height_{index}"]
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= TOL_REFERENCE, np.abs(got - want).max()


def test_single_biome_histogram_is_plain_fractal_noise(oracle_mod):
    """One bin of weight 1 per pixel: height = fractal * Variation + Depth, inside [Depth, Depth + Variation]; a biome
    without table entry contributes nothing; zero octaves give Depth (0/0 saturates to 0 like __saturatef)."""
    w, h = 20, 12
    table, perm, grad = make_generator_tables(3, 4)
    table["Octave"][2] = 0
    for biome in range(4):
        items = np.full(w * h, biome, dtype=np.uint16)
        weights = np.ones(w * h, dtype=np.float32)
        offsets = np.arange(w * h + 1, dtype=np.uint32)
        out = oracle_mod.heightfield_port(items, weights, offsets, (w, h), table, perm, grad, (3.0, 4.0))
        lo, hi = float(table["Depth"][biome]), float(table["Depth"][biome] + table["Variation"][biome])
        assert out.min() >= lo - 1e-6 and out.max() <= hi + 1e-6
        if biome == 2:
            assert np.array_equal(out, np.full((h, w), table["Depth"][2], dtype=np.float32))
        else:
            assert out.std() > 0
    out = oracle_mod.heightfield_port(np.full(w * h, 9, np.uint16), np.ones(w * h, np.float32),
                                      np.arange(w * h + 1, dtype=np.uint32), (w, h), table, perm, grad, (0.0, 0.0))
    assert not out.any()


def test_bin_order_matters_only_in_the_last_bits(oracle_mod):
    """The sum runs in bin order (STPSingleHistogramWrapper.inl:13-19): permuting a pixel's bins may change low bits only."""
    w, h = 8, 8
    table, perm, grad = make_generator_tables(4, 3)
    rng = np.random.default_rng(0)
    weights = rng.dirichlet(np.ones(3), w * h).astype(np.float32)
    offsets = (np.arange(w * h + 1) * 3).astype(np.uint32)
    a = oracle_mod.heightfield_port(np.tile([0, 1, 2], w * h).astype(np.uint16), weights.reshape(-1), offsets, (w, h),
                                    table, perm, grad, (0.0, 0.0))
    b = oracle_mod.heightfield_port(np.tile([2, 1, 0], w * h).astype(np.uint16), weights[:, ::-1].reshape(-1), offsets,
                                    (w, h), table, perm, grad, (0.0, 0.0))
    assert np.abs(a - b).max() < 1e-6
