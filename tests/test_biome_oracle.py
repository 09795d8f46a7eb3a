"""The CPU restatement of the biome-map producer (oracle/biome_oracle.c) pinned to the reference: to the stored outputs
of the reference's own compiled layer chain (tests/golden/biome_vectors.npz, made by make_biome_golden.py) and, where
that build is present (oracle/_ref/libbiome_ref.so), to the reference live on random regions."""
import os

import numpy as np
import pytest

from golden.make_biome_golden import BIOME_CASES


@pytest.fixture(scope="module")
def biome(oracle_mod):
    from oracle import biome as b

    return b


@pytest.mark.parametrize("index", range(len(BIOME_CASES)))
def test_port_against_stored_reference_maps(biome, index):
    seed, off, w, h = BIOME_CASES[index]
    stored = np.load(os.path.join(os.path.dirname(__file__), "golden", "biome_vectors.npz"))[f"map_{index}"]
    assert stored.shape == (h, w)
    assert np.array_equal(biome.biome_port(w, h, off, seed), stored)


def test_port_against_live_reference(biome):
    if not biome.have_biome_reference():
        pytest.skip("oracle/_ref/libbiome_ref.so not built here")
    rng = np.random.default_rng(2026)
    for _ in range(25):
        seed = int(rng.integers(0, 2 ** 63))
        off = (int(rng.integers(-10 ** 6, 10 ** 6)), int(rng.integers(-10 ** 6, 10 ** 6)))
        w, h = int(rng.integers(1, 500)), int(rng.integers(1, 300))
        assert np.array_equal(biome.biome_port(w, h, off, seed), biome.biome_reference(w, h, off, seed)), (seed, off, w, h)


def test_map_is_a_pure_function_of_world_coordinates(biome):
    """What makes the dense level-by-level evaluation on the device legitimate: overlapping requests agree cell by cell
    (the reference's per-layer cache never changes a value, STPLayer.cpp:77-92)."""
    a = biome.biome_port(200, 120, (-50, 30), 99)
    b = biome.biome_port(90, 70, (20, 60), 99)
    assert np.array_equal(a[30:100, 70:160], b)


def test_other_chains_and_ids(biome):
    """Chains other than the demo's (custom ids, fuzzy scales, a short chain) stay self-consistent and use only the ids
    given."""
    chain = biome.chain_array([(biome.CONTINENT, 11), (biome.SCALE_FUZZY, 12), (biome.LAND, 13), (biome.ISLAND, 14),
                               (biome.SCALE_NORMAL, 15), (biome.VORONOI, 16)])
    ids = (10, 20, 30, 40, 50, 60, 70)
    m = biome.biome_port(150, 90, (-33, 12), 5, chain=chain, ids=ids)
    assert set(np.unique(m)) <= {10, 20, 30}
    assert len(np.unique(m)) >= 2
