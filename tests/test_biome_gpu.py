"""GPU parity of the device biome-map producer (shf_biome_factory_*) through the C ABI:
bit-exact against the stored outputs of the reference's own compiled layer chain, against the CPU restatement on random
regions, chains and ids, and -- the point of the row -- the producer feeding the filter with no host copy of the map."""
import os

import numpy as np
import pytest

from golden.make_biome_golden import BIOME_CASES
from helpers import assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def filt(shf):
    return shf.STPSingleHistogramFilter()


@pytest.fixture(scope="module")
def biome(oracle_mod):
    from oracle import biome as b

    return b


def produce(shf, filt, dim, offsets, seed, chain=None, ids=None):
    import torch

    offsets = np.asarray(offsets, dtype=np.int32).reshape(-1, 2)
    out = torch.zeros((len(offsets), dim[1], dim[0]), dtype=torch.int16, device="cuda")
    if chain is None:
        fac = shf.STPLayerChainBuilder(filt, dim, seed, ids)
    else:
        fac = shf.STPBiomeFactory(filt, dim, chain, seed, ids)
    fac(out.data_ptr(), offsets, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    fac.close()
    return out.cpu().numpy().view(np.uint16)


@pytest.mark.parametrize("index", range(len(BIOME_CASES)))
def test_against_stored_reference_maps(shf, filt, index):
    seed, off, w, h = BIOME_CASES[index]
    stored = np.load(os.path.join(os.path.dirname(__file__), "golden", "biome_vectors.npz"))[f"map_{index}"]
    got = produce(shf, filt, (w, h), [off], seed)[0]
    assert np.array_equal(got, stored), f"case {index}: {int((got != stored).sum())} cells differ"


def test_random_regions_against_cpu(shf, filt, biome):
    rng = np.random.default_rng(77)
    for _ in range(20):
        seed = int(rng.integers(0, 2 ** 63))
        off = (int(rng.integers(-10 ** 6, 10 ** 6)), int(rng.integers(-10 ** 6, 10 ** 6)))
        w, h = int(rng.integers(1, 600)), int(rng.integers(1, 400))
        got = produce(shf, filt, (w, h), [off], seed)[0]
        want = biome.biome_reference(w, h, off, seed) if biome.have_biome_reference() else biome.biome_port(w, h, off, seed)
        assert np.array_equal(got, want), (seed, off, w, h)


def test_batch_of_offsets_and_windows(shf, filt, biome):
    """Many maps per call, written as windows of larger images (row stride / map stride), e.g. only the (W+2r)^2 cells
    the filter reads of every merged neighbourhood map."""
    import torch

    w, h, seed = 160, 96, 31337
    offs = [(-500 + 97 * i, 40 - 61 * i) for i in range(7)]
    img_w, img_h = 200, 120
    img = torch.full((len(offs), img_h, img_w), -1, dtype=torch.int16, device="cuda")
    fac = shf.STPLayerChainBuilder(filt, (w, h), seed)
    x0, z0 = 24, 9
    fac(img.data_ptr() + 2 * (z0 * img_w + x0), offs, row_stride=img_w, map_stride=img_w * img_h)
    torch.cuda.synchronize()
    fac.close()
    got = img.cpu().numpy().view(np.uint16)
    for i, off in enumerate(offs):
        assert np.array_equal(got[i, z0:z0 + h, x0:x0 + w], biome.biome_port(w, h, off, seed)), i
        frame = got[i].copy()
        frame[z0:z0 + h, x0:x0 + w] = 0xFFFF
        assert (frame == 0xFFFF).all(), "cells outside the window were written"


def test_other_chains_and_ids(shf, filt, biome):
    K = shf.STPLayerKind
    chain = [(K.Continent, 11), (K.ScaleFuzzy, 12), (K.Land, 13), (K.Island, 14), (K.ScaleNormal, 15), (K.Voronoi, 16)]
    ids = (10, 20, 30, 40, 50, 60, 70)
    got = produce(shf, filt, (150, 90), [(-33, 12)], 5, chain=chain, ids=ids)[0]
    want = biome.biome_port(150, 90, (-33, 12), 5, chain=biome.chain_array([(int(k), s) for k, s in chain]), ids=ids)
    assert np.array_equal(got, want)
    # a tree: two descendants share one ascendant; a layer not reachable from the root is skipped
    tree = [(K.Continent, 3), (K.ScaleNormal, 4, 0), (K.Island, 9, 0), (K.Land, 5, 1), (K.Voronoi, 6, 3)]
    got = produce(shf, filt, (70, 50), [(9, -200)], 17, chain=tree, ids=(0, 1, 3, 0, 0, 0, 0))[0]
    arr = np.zeros(len(tree), dtype=biome.LAYER_DTYPE)
    for i, t in enumerate(tree):
        arr[i] = (int(t[0]), t[2] if len(t) > 2 else max(i - 1, 0), t[1])
    assert np.array_equal(got, biome.biome_port(70, 50, (9, -200), 17, chain=arr))


def test_argument_errors(shf, filt):
    K = shf.STPLayerKind
    with pytest.raises(shf.STPNumericDomainError):     # STPBiomeFactory.cpp:20-22
        shf.STPLayerChainBuilder(filt, (0, 5), 1)
    with pytest.raises(shf.STPInvalidEnum):
        shf.STPBiomeFactory(filt, (4, 4), [(K.Continent, 1), (99, 2)], 1, (0, 1, 3, 0, 0, 0, 0))
    with pytest.raises(ValueError):
        shf.STPBiomeFactory(filt, (4, 4), [(K.Continent, 1), (K.Land, 2, 1)], 1, (0, 1, 3, 0, 0, 0, 0))


def test_producer_feeds_the_filter_without_a_host_copy(shf, filt, biome, oracle_mod):
    """biome maps are produced in HBM and filtered where they lie. Three neighbourhoods of
    3 x 3 chunks of 128 x 128; the producer writes only the (W+2r) x (H+2r) window of every merged map that the filter
    reads. Expected = the CPU filter on the CPU-produced merged maps."""
    import torch

    w, h, r, seed = 128, 128, 32, 424242
    tw, th = 3 * w, 3 * h
    chunk_origins = [(0, 0), (-4 * w, 7 * h), (1000 * w, -3 * h)]           # world coordinate of every centre chunk
    merged = torch.zeros((len(chunk_origins), th, tw), dtype=torch.int16, device="cuda")
    fac = shf.STPLayerChainBuilder(filt, (w + 2 * r, h + 2 * r), seed)
    stream = torch.cuda.current_stream().cuda_stream
    fac(merged.data_ptr() + 2 * ((h - r) * tw + (w - r)), [(x - r, z - r) for x, z in chunk_origins], row_stride=tw,
        map_stride=tw * th, stream=stream)
    info = shf.STPNearestNeighbourInformation((w, h), (3, 3), (tw, th))
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    filt.runDevice(merged.data_ptr(), tw * th, len(chunk_origins), info, buf, r, stream)
    torch.cuda.synchronize()
    fac.close()
    from superterrainplus_b200.api import DeviceArrayView
    bins_p, offs_p = buf.readDevice()
    base = buf.chunkBase()
    per = w * h + 1
    for i, (x, z) in enumerate(chunk_origins):
        cpu_map = biome.biome_port(tw, th, (x - w, z - h), seed)          # the whole merged map, on the CPU
        want = oracle_mod.run_port(cpu_map, (w, h), (3, 3), r)
        lo, hi = int(base[i]), int(base[i + 1])
        raw = torch.as_tensor(DeviceArrayView(bins_p + lo * 8, 2 * (hi - lo)), device="cuda").cpu().numpy().view(shf.BIN_DTYPE)
        offs = torch.as_tensor(DeviceArrayView(offs_p + i * per * 4, per), device="cuda").cpu().numpy().view(np.uint32)
        assert_same((raw["Item"].copy(), raw["Weight"].copy(), offs.copy()), want, f"neighbourhood {i}")
    buf.close()
