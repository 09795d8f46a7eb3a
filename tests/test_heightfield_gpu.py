"""GPU parity of the device-side consumer: the multi-biome heightfield computed from the
device-resident histogram, through the C ABI (shf_heightfield_*), against
  * the CPU restatement oracle/shf_heightfield_oracle.c  -- bit for bit (both spell out the same roundings), and
  * the reference's own device code (oracle/_ref/libshf_ref_height.so, built from /root/reference) -- within TOL, because
    the reference leaves multiply-add contraction to its run-time compiler.
"""
import ctypes

import numpy as np
import pytest

from golden.heightfield_cases import HEIGHT_CASES, make_height_case
from helpers import make_generator_tables, split_result

pytestmark = pytest.mark.gpu

# heights are sums of weights (which add up to 1) times values in [Depth, Depth + Variation] <= 2
TOL_REFERENCE = 2e-5


def d2h(ptr, n):
    out = np.zeros(n, dtype=np.float32)
    cudart = ctypes.CDLL("libcudart.so.12")
    assert cudart.cudaDeviceSynchronize() == 0
    assert cudart.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(out.nbytes), 2) == 0
    return out


@pytest.fixture(scope="module")
def filt(shf):
    return shf.STPSingleHistogramFilter()


@pytest.mark.parametrize("index", range(len(HEIGHT_CASES)))
def test_heightfield_cases(shf, filt, oracle_mod, index):
    import torch

    case = HEIGHT_CASES[index]
    m, table, perm, grad = make_height_case(case)
    w, h = case["w"], case["h"]
    info = shf.STPNearestNeighbourInformation((w, h), (3, 3), (3 * w, 3 * h))
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    items, weights, offsets = split_result(filt(m, info, buf, case["r"]))
    gen = shf.STPMultiBiomeHeightfield(filt, table, perm, grad)
    out = torch.empty((h, w), dtype=torch.float32, device="cuda")
    gen(buf, 0, 1, case["offset"], out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    got = d2h(out.data_ptr(), w * h).reshape(h, w)
    want = oracle_mod.heightfield_port(items, weights, offsets, (w, h), table, perm, grad, case["offset"])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
        f"case {index}: max |diff| {np.abs(got - want).max()} vs the CPU restatement"
    assert np.isfinite(got).all() and got.min() >= 0.0 and got.max() <= 2.0
    if oracle_mod.have_height_reference():
        ref = oracle_mod.heightfield_reference(items, weights, offsets, (w, h), table, perm, grad, case["offset"])
        assert np.abs(ref - got).max() <= TOL_REFERENCE, np.abs(ref - got).max()
    gen.close()
    buf.close()


def test_heightfield_batch_offsets_and_unknown_biomes(shf, filt, oracle_mod):
    import torch

    from test_parity_gpu import random_map

    rng = np.random.default_rng(31)
    w, h, r = 40, 28, 10
    maps = [random_map(rng, w, h, 9, k) for k in ("iid", "blocky", "rare")]
    info = shf.STPNearestNeighbourInformation((w, h), (3, 3), (3 * w, 3 * h))
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    hist = filt.runBatch(maps, info, buf, r)
    base = buf.chunkBase()
    table, perm, grad = make_generator_tables(5, 7)  # biomes 7 and 8 have no table entry: they contribute nothing
    gen = shf.STPMultiBiomeHeightfield(filt, table, perm, grad)
    offsets_xy = [(0.0, 0.0), (40.0, -28.0), (12345.0, 54321.0)]
    out = torch.empty((2, h, w), dtype=torch.float32, device="cuda")
    gen(buf, 1, 2, offsets_xy[1:], out.data_ptr())  # chunks 1 and 2 only
    got = d2h(out.data_ptr(), 2 * w * h).reshape(2, h, w)
    per = w * h + 1
    for k, chunk in enumerate((1, 2)):
        items = hist.Bin["Item"][base[chunk]:base[chunk + 1]].copy()
        weights = hist.Bin["Weight"][base[chunk]:base[chunk + 1]].copy()
        offs = hist.HistogramStartOffset[chunk * per:(chunk + 1) * per].copy()
        want = oracle_mod.heightfield_port(items, weights, offs, (w, h), table, perm, grad, offsets_xy[chunk])
        assert np.array_equal(got[k].view(np.uint32), want.view(np.uint32)), f"chunk {chunk}"
    with pytest.raises(ValueError):
        gen(buf, 2, 2, offsets_xy[1:], out.data_ptr())  # chunk 3 does not exist
    gen.close()
    buf.close()


def test_config4_device_resident_chain(shf, filt, oracle_mod):
    """BASELINE.json config 4: a 2048x2048 chunk, filter output fed to the heightfield kernel on the device with no host
    round trip. Histogram parity at this size is covered by properties in test_parity_gpu; here the heights of a band of
    rows are checked against the CPU restatement evaluated on the downloaded histogram."""
    import dataclasses

    import torch

    from superterrainplus_b200 import workloads

    wl = dataclasses.replace(workloads.CONFIGS["C4"], dist="blocky")
    w, h = wl.map_size
    dev = workloads.make_maps_torch(wl, 0, 1, torch.device("cuda"))
    info = shf.STPNearestNeighbourInformation(wl.map_size, wl.nn, wl.total)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    stream = torch.cuda.current_stream().cuda_stream
    filt.runDevice(dev.data_ptr(), wl.total[0] * wl.total[1], 1, info, buf, wl.radius, stream)
    table, perm, grad = make_generator_tables(9, wl.biomes)
    gen = shf.STPMultiBiomeHeightfield(filt, table, perm, grad)
    out = torch.empty((h, w), dtype=torch.float32, device="cuda")
    gen(buf, 0, 1, (4096.0, -2048.0), out.data_ptr(), stream)
    got = d2h(out.data_ptr(), w * h).reshape(h, w)
    n_bins, n_offs = buf.size()
    bins_p, offs_p = buf.readDevice()
    cudart = ctypes.CDLL("libcudart.so.12")
    bins = np.zeros(n_bins, dtype=shf.BIN_DTYPE)
    offs = np.zeros(n_offs, dtype=np.uint32)
    assert cudart.cudaMemcpy(ctypes.c_void_p(bins.ctypes.data), ctypes.c_void_p(bins_p), ctypes.c_size_t(bins.nbytes), 2) == 0
    assert cudart.cudaMemcpy(ctypes.c_void_p(offs.ctypes.data), ctypes.c_void_p(offs_p), ctypes.c_size_t(offs.nbytes), 2) == 0
    rows = (1000, 1012)
    want = oracle_mod.heightfield_port(bins["Item"], bins["Weight"], offs, (w, h), table, perm, grad, (4096.0, -2048.0),
                                       pixels=(rows[0] * w, rows[1] * w))
    assert np.array_equal(got[rows[0]:rows[1]].view(np.uint32), want[rows[0]:rows[1]].view(np.uint32))
    assert np.isfinite(got).all() and got.min() >= 0.0 and got.max() <= 2.0
    gen.close()
    buf.close()
