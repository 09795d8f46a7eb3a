"""GPU parity: the CUDA path, called through the C ABI (ctypes host layer), against the CPU oracle on the same inputs.

Bar: items, offsets AND weight bits identical to the reference filter. The oracle used is oracle/shf_oracle.c (pinned
to the reference's golden vector and to the reference's own compiled filter by tests/test_oracle.py); where the
reference build travelled to this box (oracle/_ref) it is checked too.
"""
import numpy as np
import pytest

from helpers import assert_same, split_result

pytestmark = pytest.mark.gpu


def nn_info(shf, w, h, nn=(3, 3)):
    return shf.STPNearestNeighbourInformation((w, h), nn, (w * nn[0], h * nn[1]))


@pytest.fixture(scope="module")
def filt(shf):
    return shf.STPSingleHistogramFilter()


from golden.reference_vector import EXPECTED, TEXTURE as GOLDEN_TEXTURE, WEIGHT_RTOL, BAD_RADII


def random_map(rng, w, h, biomes, kind, nn=(3, 3)):
    tw, th = w * nn[0], h * nn[1]
    if kind == "iid":
        m = rng.integers(0, biomes, (th, tw))
    elif kind == "blocky":
        bs = int(rng.integers(2, 9))
        m = rng.integers(0, biomes, (th // bs + 1, tw // bs + 1)).repeat(bs, 0).repeat(bs, 1)[:th, :tw]
    elif kind == "rare":
        m = np.where(rng.random((th, tw)) < 0.95, 0, rng.integers(0, biomes, (th, tw)))
    elif kind == "stripes":
        period = int(rng.integers(2, 12))
        m = np.broadcast_to((np.arange(tw) % period)[None, :] % biomes, (th, tw))
    elif kind == "hstripes":
        period = int(rng.integers(2, 12))
        m = np.broadcast_to((np.arange(th) % period)[:, None] % biomes, (th, tw))
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(m, dtype=np.uint16)


def test_golden_all_pixels(shf, filt, oracle_mod):
    info = nn_info(shf, 4, 4)
    for exec_type in (shf.STPSingleHistogramFilter.STPFilterBuffer.STPExecutionType.Serial,
                      shf.STPSingleHistogramFilter.STPFilterBuffer.STPExecutionType.Parallel):
        buf = shf.STPSingleHistogramFilter.STPFilterBuffer(exec_type)
        assert buf.size() == (0, 0)
        assert buf.readHistogram().Bin is None
        hist = filt(GOLDEN_TEXTURE, info, buf, 2)
        got = split_result(hist)
        want = oracle_mod.run_port(GOLDEN_TEXTURE, (4, 4), (3, 3), 2)
        assert_same(got, want, "golden")
        # the three pixels the reference test pins (STPTestHistogram.cpp:77-94)
        for px, bins in EXPECTED.items():
            lo, hi = int(got[2][px]), int(got[2][px + 1])
            assert list(got[0][lo:hi]) == [b[0] for b in bins]
            np.testing.assert_allclose(got[1][lo:hi], [b[1] for b in bins], rtol=WEIGHT_RTOL)
        # STPTestHistogram.cpp:150-161: same buffer again gives the same result; size and type echo
        again = split_result(filt(GOLDEN_TEXTURE, info, buf, 2))
        assert_same(again, want, "golden rerun")
        assert buf.size() == (64, 17)
        assert buf.type() == exec_type
        buf.close()


def test_error_cases(shf, filt):
    info = nn_info(shf, 4, 4)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    for bad in BAD_RADII:  # STPTestHistogram.cpp:125-129
        with pytest.raises(shf.STPNumericDomainError):
            filt(GOLDEN_TEXTURE, info, buf, bad)
    with pytest.raises(shf.STPNumericDomainError):  # 1x1 neighbourhood rejects every radius (SHF.cpp:876-880)
        filt(GOLDEN_TEXTURE, shf.STPNearestNeighbourInformation((12, 12), (1, 1), (12, 12)), buf, 2)
    with pytest.raises(shf.STPInvalidEnum):
        shf.STPSingleHistogramFilter.STPFilterBuffer(0x42)


CASES = []
_rng = np.random.default_rng(20251017)
for _i in range(48):
    _w, _h = int(_rng.integers(4, 40)), int(_rng.integers(4, 40))
    _r = 2 * int(_rng.integers(1, max(2, min(_w, _h) // 2 + 1)))
    _r = min(_r, min(_w, _h) // 2 * 2)
    CASES.append((_w, _h, _r, int(_rng.integers(1, 40)), ["iid", "blocky", "rare", "stripes", "hstripes"][_i % 5], _i))


@pytest.mark.parametrize("w,h,r,biomes,kind,seed", CASES)
def test_small_random(shf, filt, oracle_mod, w, h, r, biomes, kind, seed):
    rng = np.random.default_rng(seed)
    m = random_map(rng, w, h, biomes, kind)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    got = split_result(filt(m, nn_info(shf, w, h), buf, r))
    assert_same(got, oracle_mod.run_port(m, (w, h), (3, 3), r), f"{w}x{h} r={r} B={biomes} {kind}")
    buf.close()


@pytest.mark.parametrize("w,h,r,biomes,kind", [
    (64, 64, 32, 200, "iid"), (130, 96, 64, 256, "iid"), (48, 160, 48, 5, "blocky"), (96, 80, 16, 64, "blocky"),
    (128, 128, 64, 64, "iid"), (100, 70, 8, 33, "rare"), (72, 72, 36, 65, "stripes"), (200, 50, 50, 129, "iid"),
    (64, 200, 126, 16, "blocky"),
])
def test_medium(shf, filt, oracle_mod, w, h, r, biomes, kind):
    rng = np.random.default_rng(w * 1000 + h)
    nn = (3, 3) if r <= min(w, h) else (2 * ((r + min(w, h) - 1) // min(w, h)) + 1,) * 2
    m = random_map(rng, w, h, biomes, kind, nn)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    got = split_result(filt(m, nn_info(shf, w, h, nn), buf, r))
    assert_same(got, oracle_mod.run_port(m, (w, h), nn, r), f"{w}x{h} r={r} B={biomes} {kind}")
    buf.close()


@pytest.mark.parametrize("w,h,r,biomes,kind", [
    (130, 96, 128, 12, "blocky"),    # 2r+1 = 257: vertical window counts no longer fit 8 bits (16-bit ring)
    (192, 192, 162, 5, "iid"),       # the reference benchmark's largest radius (STPTestHistogram.cpp:266-279)
    (64, 48, 8, 1000, "iid"),        # more than 256 distinct values in one neighbourhood: wide path
    (130, 96, 128, 1024, "iid"),     # BASELINE.json C5 corner: radius 128 x 1024 biomes: wide path
    (96, 64, 64, 1024, "blocky"),
    (40, 40, 20, 5000, "iid"),
    (140, 30, 130, 300, "stripes"),
    (36, 150, 128, 400, "hstripes"),
    (64, 64, 32, 700, "rare"),
    (140, 100, 128, 40, "iid"),      # 16-bit ring, K = 2
    (150, 110, 200, 100, "iid"),     # 16-bit ring, K = 4
    (120, 90, 254, 100, "blocky"),   # 2r+1 = 509, the largest window of the 16-bit ring, K = 4
    (120, 90, 254, 200, "blocky"),   # ... with K = 8 one row's ring (557 columns x 512 B) exceeds shared memory: wide path
    (100, 80, 254, 3, "iid"),        # counts near (2r+1)^2 = 259 081 need the 32-bit horizontal sums
    (100, 80, 256, 3, "blocky"),     # 2r+1 = 513: wide path
    (257, 64, 128, 1, "iid"),        # one value: every window count is exactly 257 and 66 049
])
def test_large_radius_and_wide_path(shf, filt, oracle_mod, w, h, r, biomes, kind):
    """Shapes outside the 8-bit ring: radius 128..254 (16-bit vertical counts, 32-bit window sums) and the wide path
    (shf_generic.cuh: radius above 254 or more than 256 distinct values)."""
    rng = np.random.default_rng(w * 7919 + h * 31 + r)
    nn = (2 * ((r + w - 1) // w) + 1, 2 * ((r + h - 1) // h) + 1)
    m = random_map(rng, w, h, biomes, kind, nn)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    got = split_result(filt(m, nn_info(shf, w, h, nn), buf, r))
    assert_same(got, oracle_mod.run_port(m, (w, h), nn, r), f"{w}x{h} r={r} B={biomes} {kind}")
    sx, sy = w * (nn[0] // 2), h * (nn[1] // 2)
    distinct = len(np.unique(m[sy - r:sy + h + r, sx - r:sx + w + r]))
    plan = buf.lastPlan()
    if distinct > 256 or 2 * r + 1 > 511:
        assert plan["k_sets"] == 0   # the wide path ran
    elif distinct <= 128:
        assert plan["k_sets"] > 0    # the event-list path ran (129..256 values: only while one row's ring fits)
    buf.close()


def test_wide_path_batch(shf, filt, oracle_mod):
    rng = np.random.default_rng(77)
    w, h, r = 48, 40, 16
    maps = [random_map(rng, w, h, b, k) for b, k in ((900, "iid"), (3, "blocky"), (300, "rare"))]
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0x00)
    hist = filt.runBatch(maps, nn_info(shf, w, h), buf, r)
    base = buf.chunkBase()
    per = w * h + 1
    for i, m in enumerate(maps):
        got = (hist.Bin["Item"][base[i]:base[i + 1]].copy(), hist.Bin["Weight"][base[i]:base[i + 1]].copy(),
               hist.HistogramStartOffset[i * per:(i + 1) * per].copy())
        assert_same(got, oracle_mod.run_port(m, (w, h), (3, 3), r), f"wide batch chunk {i}")
    buf.close()


def test_sparse_ids_and_non_square_neighbourhood(shf, filt, oracle_mod):
    # sample values are arbitrary uint16 (SHF.cpp:404-410 grows its dictionary to max id + 1)
    rng = np.random.default_rng(5)
    ids = np.array([0, 7, 300, 4095, 65535, 12345, 32768], dtype=np.uint16)
    m = ids[rng.integers(0, len(ids), (5 * 20, 3 * 24))]
    info = shf.STPNearestNeighbourInformation((24, 20), (3, 5), (72, 100))
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0x00)
    got = split_result(filt(m, info, buf, 10))
    assert_same(got, oracle_mod.run_port(m, (24, 20), (3, 5), 10), "sparse ids")
    buf.close()


def test_batch_matches_single(shf, filt, oracle_mod):
    rng = np.random.default_rng(9)
    w, h, r = 40, 24, 12
    maps = [random_map(rng, w, h, 20, k) for k in ("iid", "blocky", "rare", "stripes", "iid")]
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    hist = filt.runBatch(maps, nn_info(shf, w, h), buf, r)
    base = buf.chunkBase()
    assert len(base) == len(maps) + 1 and base[-1] == len(hist.Bin)
    per = w * h + 1
    for i, m in enumerate(maps):
        items = hist.Bin["Item"][base[i]:base[i + 1]].copy()
        weights = hist.Bin["Weight"][base[i]:base[i + 1]].copy()
        offs = hist.HistogramStartOffset[i * per:(i + 1) * per].copy()
        assert_same((items, weights, offs), oracle_mod.run_port(m, (w, h), (3, 3), r), f"batch chunk {i}")
    buf.close()


SWEEP = [(r, b, kind) for r in (8, 16, 32, 64, 128) for b in (4, 16, 64, 256, 1024) for kind in ("iid", "blocky")]


@pytest.mark.parametrize("r,biomes,kind", SWEEP)
def test_config5_radius_biome_sweep(shf, filt, oracle_mod, r, biomes, kind):
    """BASELINE.json config 5: radius 8-128 x biome count 4-1024, dense (iid) and sparse (blocky) histograms, on maps the
    oracle finishes in well under a second; both the register-list path and the wide path are crossed."""
    w, h = 72, 40
    nn = (2 * ((r + w - 1) // w) + 1, 2 * ((r + h - 1) // h) + 1)
    rng = np.random.default_rng(r * 4099 + biomes)
    m = random_map(rng, w, h, biomes, kind, nn)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    got = split_result(filt(m, nn_info(shf, w, h, nn), buf, r))
    assert_same(got, oracle_mod.run_port(m, (w, h), nn, r), f"sweep r={r} B={biomes} {kind}")
    buf.close()


@pytest.mark.parametrize("w,h,r,biomes,kind,segments", [
    (70, 300, 8, 6, "rare", 4), (70, 300, 8, 30, "hstripes", 8), (96, 260, 32, 12, "blocky", 3), (64, 200, 64, 5, "iid", 2),
    (40, 520, 16, 70, "iid", 8), (130, 400, 100, 3, "blocky", 3), (50, 330, 130, 9, "hstripes", 2), (33, 700, 4, 2, "rare", 8),
])
def test_vscan_row_segments(shf, filt, oracle_mod, w, h, r, biomes, kind, segments, monkeypatch):
    """Small calls split the vertical scan into row segments whose chain starts are resolved on demand by events_kernel (resolve_start):
    chains that never break (hstripes), chains that break inside segments (rare / small radius), 16-bit ring, and the
    same map with the segmentation switched off must all agree with the oracle."""
    rng = np.random.default_rng(w * 131 + h * 7 + r)
    nn = (2 * ((r + w - 1) // w) + 1, 2 * ((r + h - 1) // h) + 1)
    m = random_map(rng, w, h, biomes, kind, nn)
    want = oracle_mod.run_port(m, (w, h), nn, r)
    for setting in (str(segments), None, "off"):
        if setting == "off":
            monkeypatch.setenv("SHF_NO_VSEG", "1")
        elif setting is None:
            monkeypatch.delenv("SHF_DEBUG_VSEG", raising=False)
        else:
            monkeypatch.setenv("SHF_DEBUG_VSEG", setting)
        local = shf.STPSingleHistogramFilter()   # the measurement toggles are read when a filter is created
        buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
        got = split_result(local(m, nn_info(shf, w, h, nn), buf, r))
        assert_same(got, want, f"{w}x{h} r={r} B={biomes} {kind} segments={setting}")
        buf.close()
        local.close()


@pytest.mark.parametrize("w,h,r,biomes,kind,segments", [
    (300, 40, 8, 6, "rare", 4), (520, 33, 16, 30, "stripes", 8), (260, 96, 32, 12, "blocky", 3), (200, 64, 64, 5, "iid", 2),
    (640, 20, 16, 70, "iid", 8), (400, 50, 100, 3, "blocky", 2), (530, 30, 130, 9, "stripes", 2), (700, 12, 4, 2, "rare", 8),
    (333, 17, 32, 100, "iid", 5),
])
def test_emit_column_segments(shf, filt, oracle_mod, w, h, r, biomes, kind, segments, monkeypatch):
    """Small calls split every row tile of the emitting kernel into column segments (each slides over the 2r columns before
    its first pixel and starts its offsets after the bins of the pixels to its left): forced segment counts, the
    planner's own choice and no segments must all agree with the oracle."""
    rng = np.random.default_rng(w * 17 + h * 1009 + r)
    nn = (2 * ((r + w - 1) // w) + 1, 2 * ((r + h - 1) // h) + 1)
    m = random_map(rng, w, h, biomes, kind, nn)
    want = oracle_mod.run_port(m, (w, h), nn, r)
    for setting in (str(segments), None, "off"):
        if setting == "off":
            monkeypatch.setenv("SHF_NO_CSEG", "1")
        elif setting is None:
            monkeypatch.delenv("SHF_DEBUG_CSEG", raising=False)
        else:
            monkeypatch.setenv("SHF_DEBUG_CSEG", setting)
        local = shf.STPSingleHistogramFilter()   # the measurement toggles are read when a filter is created
        buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
        got = split_result(local(m, nn_info(shf, w, h, nn), buf, r))
        assert_same(got, want, f"{w}x{h} r={r} B={biomes} {kind} column segments={setting}")
        buf.close()
        local.close()


@pytest.mark.parametrize("w,h,r,biomes,kind,sms", [
    (300, 70, 8, 6, "rare", 1), (520, 33, 16, 30, "stripes", 2), (260, 96, 32, 12, "blocky", 1), (200, 64, 64, 5, "iid", 3),
    (640, 40, 16, 70, "iid", 2), (400, 50, 100, 3, "blocky", 1), (130, 90, 130, 9, "stripes", 2), (48, 100, 4, 2, "rare", 1),
    (333, 57, 32, 100, "iid", 4), (17, 200, 6, 40, "iid", 1), (96, 96, 2, 200, "iid", 2),
])
def test_emit_persistent_ctas(shf, oracle_mod, w, h, r, biomes, kind, sms, monkeypatch):
    """Large calls run the emitting kernel with persistent CTAs that walk the flat tile list, the first tile of every
    CTA cut in two parts (emitted first and last) so that the CTAs run out of phase. Forced here on small batches with
    the CTAs of 1-4 "SMs": several tiles per CTA, cuts at every multiple of 16, ragged last tiles, rings that wrap
    between items; all chunks against the oracle, persistent and not."""
    rng = np.random.default_rng(w * 31 + h * 977 + r)
    nn = (2 * ((r + w - 1) // w) + 1, 2 * ((r + h - 1) // h) + 1)
    maps = [random_map(rng, w, h, biomes, kind, nn) for _ in range(3)]
    want = [oracle_mod.run_port(m, (w, h), nn, r) for m in maps]
    monkeypatch.setenv("SHF_NO_CSEG", "1")
    for setting in (str(sms), None):
        if setting is None:
            monkeypatch.delenv("SHF_DEBUG_PERSIST", raising=False)
        else:
            monkeypatch.setenv("SHF_DEBUG_PERSIST", setting)
        local = shf.STPSingleHistogramFilter()
        buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
        hist = local.runBatch(maps, nn_info(shf, w, h, nn), buf, r)
        base = buf.chunkBase()
        per = w * h + 1
        for i in range(3):
            got = (hist.Bin["Item"][base[i]:base[i + 1]].copy(), hist.Bin["Weight"][base[i]:base[i + 1]].copy(),
                   hist.HistogramStartOffset[i * per:(i + 1) * per].copy())
            assert_same(got, want[i], f"{w}x{h} r={r} B={biomes} {kind} persistent={setting} chunk {i}")
        buf.close()
        local.close()


def split_neighbours(m, w, h, nn):
    """The nn.x * nn.y chunk maps of a merged map, in local-index order (STPChunk::calcLocalChunkCoordinate)."""
    return [np.ascontiguousarray(m[cy * h:(cy + 1) * h, cx * w:(cx + 1) * w]) for cy in range(nn[1]) for cx in range(nn[0])]


@pytest.mark.parametrize("w,h,r,nn,biomes,kind", [
    (16, 12, 4, (3, 3), 9, "iid"), (33, 20, 20, (3, 3), 6, "blocky"), (24, 24, 24, (3, 3), 40, "iid"),
    (10, 14, 10, (5, 3), 7, "rare"), (8, 8, 16, (5, 5), 5, "blocky"), (40, 36, 8, (3, 5), 300, "iid"),
])
def test_unmerged_neighbours_match_merged(shf, filt, oracle_mod, w, h, r, nn, biomes, kind):
    """Row f2: the filter fed by the separate chunk maps (no merged buffer) == the filter on the merged map == oracle.
    Covers halos that span more than one neighbour (r > W with 5 neighbours across) and the wide path (300 values)."""
    rng = np.random.default_rng(w * 1000 + h * 10 + r)
    info = nn_info(shf, w, h, nn)
    maps = [random_map(rng, w, h, biomes, kind, nn) for _ in range(3)]
    want = [oracle_mod.run_port(m, (w, h), nn, r) for m in maps]
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(shf.STPSingleHistogramFilter.STPFilterBuffer.STPExecutionType.Parallel)
    flat = [c for m in maps for c in split_neighbours(m, w, h, nn)]
    hist = filt.runNeighbours(flat, info, buf, r)
    base = buf.chunkBase()
    stride = w * h + 1
    for i, wnt in enumerate(want):
        lo, hi = int(base[i]), int(base[i + 1])
        got = (hist.Bin["Item"][lo:hi].copy(), hist.Bin["Weight"][lo:hi].copy(),
               hist.HistogramStartOffset[i * stride:(i + 1) * stride].copy())
        assert_same(got, wnt, f"unmerged neighbourhood {i}")
    # a single neighbourhood through the same entry point, and the argument checks
    one = filt.runNeighbours(split_neighbours(maps[1], w, h, nn), info, buf, r)
    assert_same(split_result(one), want[1], "single unmerged neighbourhood")
    with pytest.raises(shf.STPNumericDomainError):
        filt.runNeighbours(split_neighbours(maps[0], w, h, nn), info, buf, 3)
    with pytest.raises(ValueError):
        filt.runNeighbours(flat[:-1], info, buf, r)


def test_unmerged_neighbours_from_device_memory(shf, filt, oracle_mod):
    """The same entry point with the chunk maps already in device memory and the result left there."""
    import torch

    w, h, r, nn, biomes = 64, 48, 32, (3, 3), 20
    rng = np.random.default_rng(99)
    info = nn_info(shf, w, h, nn)
    maps = [random_map(rng, w, h, biomes, "blocky", nn) for _ in range(2)]
    chunks = [torch.from_numpy(c.view(np.int16)).cuda() for m in maps for c in split_neighbours(m, w, h, nn)]
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(shf.STPSingleHistogramFilter.STPFilterBuffer.STPExecutionType.Parallel)
    stream = torch.cuda.current_stream().cuda_stream
    filt.runNeighboursDevice([c.data_ptr() for c in chunks], info, buf, r, stream)
    torch.cuda.synchronize()
    bins_p, offs_p = buf.readDevice()
    n_bins, n_offs = buf.size()
    bins = torch.empty(n_bins * 8, dtype=torch.uint8, device="cuda")
    offs = torch.empty(n_offs, dtype=torch.int32, device="cuda")
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12")
    assert cudart.cudaMemcpy(ctypes.c_void_p(bins.data_ptr()), ctypes.c_void_p(bins_p), ctypes.c_size_t(n_bins * 8), 3) == 0
    assert cudart.cudaMemcpy(ctypes.c_void_p(offs.data_ptr()), ctypes.c_void_p(offs_p), ctypes.c_size_t(n_offs * 4), 3) == 0
    rec = bins.cpu().numpy().view(shf.BIN_DTYPE)
    items, weights, offsets = rec["Item"].copy(), rec["Weight"].copy(), offs.cpu().numpy().view(np.uint32)
    base = buf.chunkBase()
    stride = w * h + 1
    for i, m in enumerate(maps):
        lo, hi = int(base[i]), int(base[i + 1])
        assert_same((items[lo:hi], weights[lo:hi], offsets[i * stride:(i + 1) * stride]),
                    oracle_mod.run_port(m, (w, h), nn, r), f"device neighbourhood {i}")


def test_multi_call_pass_fills_every_callers_buffer(shf, filt, oracle_mod):
    """Row f3 at the C ABI: shf_run_multi == one shf_run per (map, buffer) pair, bit for bit, including size()/type() and
    a later ordinary call on one of the buffers."""
    FB = shf.STPSingleHistogramFilter.STPFilterBuffer
    rng = np.random.default_rng(4242)
    w, h, r = 40, 28, 12
    info = nn_info(shf, w, h)
    kinds = ["iid", "blocky", "rare", "stripes", "hstripes"]
    maps = [random_map(rng, w, h, 5 + 9 * i, kinds[i]) for i in range(5)]
    bufs = [FB(FB.STPExecutionType.Parallel if i % 2 else FB.STPExecutionType.Serial) for i in range(5)]
    hists = filt.runMulti(maps, info, bufs, r)
    for i, (m, hist, buf) in enumerate(zip(maps, hists, bufs)):
        want = oracle_mod.run_port(m, (w, h), (3, 3), r)
        assert_same(split_result(hist), want, f"multi call {i}")
        assert buf.size() == (len(want[0]), w * h + 1)
        assert buf.type() == (FB.STPExecutionType.Parallel if i % 2 else FB.STPExecutionType.Serial)
    # buffers stay ordinary buffers: a follower can serve a normal call next, the leader as well
    for i in (3, 0):
        again = split_result(filt(maps[(i + 1) % 5], info, bufs[i], r))
        assert_same(again, oracle_mod.run_port(maps[(i + 1) % 5], (w, h), (3, 3), r), f"reuse of buffer {i}")
    with pytest.raises(ValueError):  # one buffer cannot serve two calls of the same pass
        filt.runMulti(maps[:2], info, [bufs[0], bufs[0]], r)
    with pytest.raises(shf.STPNumericDomainError):
        filt.runMulti(maps[:2], info, bufs[:2], 5)
    for b in bufs:
        b.close()


def test_device_resident(shf, filt, oracle_mod):
    import torch

    rng = np.random.default_rng(11)
    w, h, r = 32, 32, 8
    maps = np.stack([random_map(rng, w, h, 12, k) for k in ("iid", "blocky", "rare")])
    dev = torch.from_numpy(maps.view(np.int16)).cuda()
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    filt.runDevice(dev.data_ptr(), maps.shape[1] * maps.shape[2], 3, nn_info(shf, w, h), buf, r,
                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    bins_p, offs_p = buf.readDevice()
    base = buf.chunkBase()
    n_bins, n_offs = buf.size()
    assert n_bins == base[-1] and n_offs == 3 * (w * h + 1)
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12")
    bins = np.zeros(n_bins, dtype=shf.BIN_DTYPE)
    offs = np.zeros(n_offs, dtype=np.uint32)
    assert cudart.cudaMemcpy(ctypes.c_void_p(bins.ctypes.data), ctypes.c_void_p(bins_p), ctypes.c_size_t(bins.nbytes), 2) == 0
    assert cudart.cudaMemcpy(ctypes.c_void_p(offs.ctypes.data), ctypes.c_void_p(offs_p), ctypes.c_size_t(offs.nbytes), 2) == 0
    per = w * h + 1
    for i in range(3):
        got = (bins["Item"][base[i]:base[i + 1]].copy(), bins["Weight"][base[i]:base[i + 1]].copy(),
               offs[i * per:(i + 1) * per].copy())
        assert_same(got, oracle_mod.run_port(maps[i], (w, h), (3, 3), r), f"device chunk {i}")
    buf.close()


def cpu_filter(oracle_mod, m, wl):
    """The reference's own compiled filter when it travelled to this box (oracle/_ref), else the C restatement."""
    if oracle_mod.have_reference():
        return oracle_mod.run_reference(m, wl.map_size, wl.nn, wl.radius)
    return oracle_mod.run_port(m, wl.map_size, wl.nn, wl.radius)


@pytest.mark.parametrize("name,dist", [("C1", "uniform"), ("C1", "blocky"), ("C2", "uniform"), ("C2", "blocky"),
                                       ("C4", "blocky"), ("C4", "uniform")])
def test_baseline_configs_vs_oracle(shf, filt, oracle_mod, name, dist):
    """BASELINE.json's single-neighbourhood configs at their full sizes (512x512 r=32 B=8, 1024x1024 r=64 B=32, one
    2048x2048 chunk r=64 B=32), dense and clustered ids: every pixel's items, offsets and weight bits."""
    import dataclasses
    from superterrainplus_b200 import workloads

    wl = dataclasses.replace(workloads.CONFIGS[name], dist=dist)
    m = workloads.make_map_np(wl, 0)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    got = split_result(filt(m, shf.STPNearestNeighbourInformation(wl.map_size, wl.nn, wl.total), buf, wl.radius))
    assert_same(got, cpu_filter(oracle_mod, m, wl), f"{name} {dist}")
    buf.close()


@pytest.mark.parametrize("dist", ["uniform", "blocky"])
def test_c3_full_batch_256_chunks(shf, filt, oracle_mod, dist):
    """The configuration every headline number is quoted on, at its full size: 256 neighbourhoods of 512x512, r=64, 64
    biomes, device-resident (the bench's own call sequence: one checked call, then shf_run_device_async on OTHER chunk
    ids with the first call's plan, completed by wait()). Every chunk through size-independent properties evaluated on
    the device (offsets start at 0, grow by 1..64 per pixel and close on the chunk's bin total; every weight is
    count * 1/(2r+1)^2 bit for bit with count >= 1; counts of a pixel sum to (2r+1)^2; items distinct within a pixel);
    four chunks against the reference CPU filter in full."""
    import dataclasses
    import torch
    from superterrainplus_b200 import workloads

    wl = dataclasses.replace(workloads.CONFIGS["C3"], dist=dist)
    n = wl.chunks
    w, h = wl.map_size
    tw, th = wl.total
    dev = torch.device("cuda", 0)
    info = shf.STPNearestNeighbourInformation(wl.map_size, wl.nn, wl.total)
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    stream = torch.cuda.current_stream().cuda_stream
    warm = workloads.make_maps_torch(wl, 1000, n, dev)
    filt.runDevice(warm.data_ptr(), th * tw, n, info, buf, wl.radius, stream)
    torch.cuda.synchronize()
    del warm
    maps = workloads.make_maps_torch(wl, 0, n, dev)
    filt.runDeviceAsync(maps.data_ptr(), th * tw, n, info, buf, wl.radius, stream)
    buf.wait()
    torch.cuda.synchronize()
    bins_p, offs_p = buf.readDevice()
    base = buf.chunkBase()
    n_bins, n_offs = buf.size()
    per = w * h + 1
    assert n_offs == n * per and n_bins == int(base[-1]) and len(base) == n + 1
    total = (2 * wl.radius + 1) ** 2
    inv = torch.tensor(1.0, dtype=torch.float32, device=dev) / torch.tensor(float(total), dtype=torch.float32, device=dev)
    from superterrainplus_b200.api import DeviceArrayView
    for i in range(n):
        lo, hi = int(base[i]), int(base[i + 1])
        raw = torch.as_tensor(DeviceArrayView(bins_p + lo * 8, 2 * (hi - lo)), device=dev).view(hi - lo, 2)
        offs = torch.as_tensor(DeviceArrayView(offs_p + i * per * 4, per), device=dev)
        offs = offs.to(torch.int64) & 0xFFFFFFFF
        assert int(offs[0]) == 0 and int(offs[-1]) == hi - lo, f"chunk {i}: offsets do not span the chunk's bins"
        nb = offs[1:] - offs[:-1]
        assert int(nb.min()) >= 1 and int(nb.max()) <= wl.biomes, f"chunk {i}: bins per pixel out of range"
        items = raw[:, 0].to(torch.int64) & 0xFFFF
        assert int((raw[:, 0].to(torch.int64) & 0xFFFF0000).abs().max()) == 0, f"chunk {i}: padding bytes not zero"
        weights = raw[:, 1].view(torch.float32)
        counts = torch.round(weights.to(torch.float64) * total).to(torch.int64)
        assert int(counts.min()) >= 1
        assert torch.equal((counts.to(torch.float32) * inv).view(torch.int32), raw[:, 1]), f"chunk {i}: weight bits"
        csum = torch.cumsum(counts, 0)
        ends = csum[offs[1:] - 1]
        sums = ends - torch.cat([ends.new_zeros(1), ends[:-1]])
        assert bool((sums == total).all()), f"chunk {i}: window counts do not sum to (2r+1)^2"
        px = torch.repeat_interleave(torch.arange(w * h, device=dev, dtype=torch.int64), nb)
        key = torch.sort(px * 65536 + items).values
        assert bool((key[1:] != key[:-1]).all()), f"chunk {i}: an item appears twice in one pixel"
        if i in (0, 85, 170, 255):
            got = (items.to(torch.int32).cpu().numpy().astype(np.uint16), weights.cpu().numpy().copy(),
                   offs.cpu().numpy().astype(np.uint32))
            assert_same(got, cpu_filter(oracle_mod, maps[i].cpu().numpy(), wl), f"C3 {dist} chunk {i}")
        del raw, offs, items, weights, counts, csum, ends, sums, px, key
    buf.close()


def test_async_call_with_unfitting_plan_is_repeated(shf, filt, oracle_mod):
    """shf_run_device_async runs with the previous call's plan; when the new maps need another one (more distinct values
    -> more register sets, more bins than the buffer holds) the first query repeats the call on the checked path."""
    import torch

    w, h, r = 96, 80, 16
    rng = np.random.default_rng(99)
    info = nn_info(shf, w, h)
    small = [random_map(rng, w, h, 5, "blocky") for _ in range(3)]
    big = [random_map(rng, w, h, 150, "iid") for _ in range(3)]
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    dev = torch.device("cuda", 0)
    d_small = torch.from_numpy(np.stack(small).view(np.int16)).to(dev)
    d_big = torch.from_numpy(np.stack(big).view(np.int16)).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    filt.runDevice(d_small.data_ptr(), 9 * w * h, 3, info, buf, r, stream)
    before = buf.wait()
    filt.runDeviceAsync(d_small.data_ptr(), 9 * w * h, 3, info, buf, r, stream)   # abandoned: replaced by the next call
    filt.runDeviceAsync(d_big.data_ptr(), 9 * w * h, 3, info, buf, r, stream)
    assert buf.wait() == before + 1
    torch.cuda.synchronize()
    bins_p, offs_p = buf.readDevice()
    base = buf.chunkBase()
    n_bins, n_offs = buf.size()
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12")
    bins = np.zeros(n_bins, dtype=shf.BIN_DTYPE)
    offs = np.zeros(n_offs, dtype=np.uint32)
    assert cudart.cudaMemcpy(ctypes.c_void_p(bins.ctypes.data), ctypes.c_void_p(bins_p), ctypes.c_size_t(bins.nbytes), 2) == 0
    assert cudart.cudaMemcpy(ctypes.c_void_p(offs.ctypes.data), ctypes.c_void_p(offs_p), ctypes.c_size_t(offs.nbytes), 2) == 0
    per = w * h + 1
    for i in range(3):
        got = (bins["Item"][base[i]:base[i + 1]].copy(), bins["Weight"][base[i]:base[i + 1]].copy(), offs[i * per:(i + 1) * per].copy())
        assert_same(got, oracle_mod.run_port(big[i], (w, h), (3, 3), r), f"repeated async chunk {i}")
    # and once the plan fits, nothing is repeated
    filt.runDeviceAsync(d_big.data_ptr(), 9 * w * h, 3, info, buf, r, stream)
    assert buf.wait() == before + 1
    buf.close()


def test_wide_maps_leave_the_16_bit_event_records(shf, filt, oracle_mod):
    """W + 2r >= 65536: pixel columns no longer fit the event records' 16 bits; such maps take the wide path."""
    w, h, r = 70000, 4, 2
    rng = np.random.default_rng(5)
    m = random_map(rng, w, h, 9, "rare")
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    got = split_result(filt(m, nn_info(shf, w, h), buf, r))
    assert buf.lastPlan()["k_sets"] == 0
    assert_same(got, oracle_mod.run_port(m, (w, h), (3, 3), r), "W=70000")
    buf.close()


def test_halo_outside_the_neighbourhood_is_rejected(shf, filt):
    """An even neighbour count leaves the right / bottom halo without a source (validate() only guards left / top)."""
    w, h, r = 16, 16, 4
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    maps = [np.zeros((h, w), dtype=np.uint16) for _ in range(4)]
    with pytest.raises(ValueError):
        filt.runNeighbours(maps, shf.STPNearestNeighbourInformation((w, h), (2, 2), (2 * w, 2 * h)), buf, r)
    merged = np.zeros((2 * h, 2 * w + 8), dtype=np.uint16)   # wide enough in x, too short in y
    with pytest.raises(ValueError):
        filt(merged, shf.STPNearestNeighbourInformation((w, h), (2, 2), (2 * w + 8, 2 * h)), buf, r)
    buf.close()


def test_calls_leave_the_current_device_alone(shf, filt):
    """The reference's operator() never touches CUDA device state; the library restores the caller's device."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    torch.cuda.set_device(1)
    try:
        buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
        f0 = shf.STPSingleHistogramFilter(0)
        f0(GOLDEN_TEXTURE, nn_info(shf, 4, 4), buf, 2)
        assert torch.cuda.current_device() == 1
        import ctypes
        dev = ctypes.c_int(-1)
        ctypes.CDLL("libcudart.so.12").cudaGetDevice(ctypes.byref(dev))
        assert dev.value == 1
        buf.close()
        f0.close()
    finally:
        torch.cuda.set_device(0)


def test_c3_batch_properties_and_sampled_chunks(shf, filt, oracle_mod):
    """C3 shape (512x512, r=64, 64 biomes) on a 6-chunk batch: two chunks against the oracle in full, all chunks through
    size-independent properties (offsets monotone and consistent with the bin total, window counts sum to (2r+1)^2,
    items distinct within a pixel, every bin non-empty)."""
    import dataclasses
    from superterrainplus_b200 import workloads

    wl = dataclasses.replace(workloads.CONFIGS["C3"], chunks=6)
    maps = [workloads.make_map_np(dataclasses.replace(wl, dist=("uniform" if i % 2 == 0 else "blocky")), i)
            for i in range(wl.chunks)]
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    hist = filt.runBatch(maps, shf.STPNearestNeighbourInformation(wl.map_size, wl.nn, wl.total), buf, wl.radius)
    base = buf.chunkBase()
    w, h = wl.map_size
    per = w * h + 1
    total = (2 * wl.radius + 1) ** 2
    inv = np.float32(1.0) / np.float32(total)
    for i in range(wl.chunks):
        offs = hist.HistogramStartOffset[i * per:(i + 1) * per].astype(np.int64)
        items = hist.Bin["Item"][base[i]:base[i + 1]]
        weights = hist.Bin["Weight"][base[i]:base[i + 1]]
        assert offs[0] == 0 and offs[-1] == base[i + 1] - base[i]
        nb = np.diff(offs)
        assert (nb >= 1).all() and (nb <= wl.biomes).all()
        counts = np.rint(weights.astype(np.float64) * total).astype(np.int64)
        assert (counts >= 1).all()
        assert np.array_equal((counts.astype(np.float32) * inv).view(np.uint32), weights.view(np.uint32))
        sums = np.add.reduceat(counts, offs[:-1])
        assert (sums == total).all()
        # distinct items per pixel: sort (pixel, item) pairs and look for equal neighbours
        px = np.repeat(np.arange(w * h, dtype=np.int64), nb)
        key = np.sort(px * 65536 + items.astype(np.int64))
        assert (np.diff(key) != 0).all()
        if i < 2:
            want = oracle_mod.run_port(maps[i], wl.map_size, wl.nn, wl.radius)
            assert_same((items.copy(), weights.copy(), offs.astype(np.uint32)), want, f"C3 chunk {i}")
    buf.close()


from golden.cases import CASES as GOLDEN_CASES


@pytest.mark.parametrize("index", range(len(GOLDEN_CASES)))
def test_against_stored_reference_outputs(shf, filt, index):
    """CUDA path vs outputs of the reference's own compiled filter, committed in tests/golden/ref_vectors.npz."""
    import os

    from golden.cases import CASES, make_case

    case = CASES[index]
    stored = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
    want = (stored[f"items_{index}"], stored[f"weights_{index}"], stored[f"offsets_{index}"])
    info = shf.STPNearestNeighbourInformation((case["w"], case["h"]), case["nn"],
                                              (case["w"] * case["nn"][0], case["h"] * case["nn"][1]))
    buf = shf.STPSingleHistogramFilter.STPFilterBuffer(0xFF)
    assert_same(split_result(filt(make_case(case), info, buf, case["r"])), want, f"golden case {index}")
    buf.close()
