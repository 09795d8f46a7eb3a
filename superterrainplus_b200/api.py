"""ctypes host layer over include/shf_b200.h with the reference's class and method names.

Reference interface mirrored here (all under SuperTerrain+/):
  SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPSingleHistogramFilter.h:36-192  (filter, filter buffer, operator())
  SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPSingleHistogram.hpp:15-50        (output view)
  SuperTerrain+/Public/SuperTerrain+/World/Chunk/STPNearestNeighbourInformation.hpp:13-25
  SuperTerrain+/Public/SuperTerrain+/Exception/STPNumericDomainError.h, STPInvalidEnum.h; Utility/STPDeviceErrorHandler.hpp
"""
from __future__ import annotations

import ctypes
import enum
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libshf_b200.so"

# STPSingleHistogram::STPBin: {uint16 Item; float Weight}, 8 bytes with 2 bytes of padding after Item
BIN_DTYPE = np.dtype({"names": ["Item", "Weight"], "formats": ["<u2", "<f4"], "offsets": [0, 4], "itemsize": 8})

# STPDemo::STPBiomeProperty (SuperDemo+/World/Biomes/STPBiomeProperty.hpp:10-27), 24 bytes
BIOME_PROPERTY_DTYPE = np.dtype([("Scale", "<f4"), ("Octave", "<u4"), ("Persistence", "<f4"), ("Lacunarity", "<f4"),
                                 ("Depth", "<f4"), ("Variation", "<f4")])

SHF_OK, SHF_ERR_NUMERIC_DOMAIN, SHF_ERR_INVALID_ENUM, SHF_ERR_CUDA = 0, 1, 2, 3
SHF_ERR_OFFSET_OVERFLOW, SHF_ERR_UNSUPPORTED, SHF_ERR_INVALID_ARGUMENT = 4, 5, 6

# every symbol include/shf_b200.h declares (tests check the library exports exactly these)
C_ABI_SYMBOLS = (
    "shf_filter_create", "shf_filter_destroy", "shf_buffer_create", "shf_buffer_destroy", "shf_buffer_read",
    "shf_buffer_size", "shf_buffer_type", "shf_run", "shf_run_batch", "shf_run_device", "shf_run_device_async", "shf_buffer_wait", "shf_run_neighbours",
    "shf_run_neighbours_device", "shf_run_multi", "shf_buffer_read_device",
    "shf_buffer_chunk_base", "shf_last_error", "shf_stats_reset", "shf_stats_get", "shf_buffer_last_plan",
    "shf_set_profiling", "shf_buffer_phase_ms", "shf_buffer_phase_history", "shf_heightfield_create", "shf_heightfield_destroy", "shf_heightfield_run",
    "shf_biome_factory_create", "shf_biome_factory_destroy", "shf_biome_factory_run",
)
PHASES = ("dictionary", "remap_vscan", "events", "rowscan", "host_gap", "emit")


class STPBasic(Exception):
    """Base of the exceptions this path can raise (STPException::STPFundamentalException::STPBasic)."""


class STPNumericDomainError(STPBasic):
    """Radius not a positive even number, or wider than the neighbour ring (SHF.cpp:874,879-880)."""


class STPInvalidEnum(STPBasic):
    """Execution type is not a STPFilterBuffer::STPExecutionType (SHF.cpp:721)."""


class STPCUDAError(STPBasic):
    """A CUDA call failed, or there is no device (STP_CHECK_CUDA)."""


class STPUnsupportedError(STPBasic):
    """Shape outside what the kernels implement, or more than 2^32 bins in a chunk."""


def library_path() -> str:
    # SHF_LIBRARY: another build of the same library (A/B measurements of kernel variants); never a different implementation
    return os.environ.get("SHF_LIBRARY") or os.path.join(_HERE, _LIB_NAME)


_lib = None
_U32x2 = ctypes.c_uint32 * 2


def library() -> ctypes.CDLL:
    """The CUDA library. Fails loudly when it has not been built: there is no other implementation to fall back on."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise STPCUDAError(f"{path} is missing: build it with `python -m superterrainplus_b200.build` "
                           "(or __graft_entry__.build()); this package has no CPU implementation")
    lib = ctypes.CDLL(path)
    vp, u32, u64, sz = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t
    P = ctypes.POINTER
    lib.shf_filter_create.argtypes = [P(vp), ctypes.c_int]
    lib.shf_filter_destroy.argtypes = [vp]
    lib.shf_filter_destroy.restype = None
    lib.shf_buffer_create.argtypes = [P(vp), ctypes.c_ubyte]
    lib.shf_buffer_destroy.argtypes = [vp]
    lib.shf_buffer_destroy.restype = None
    lib.shf_buffer_read.argtypes = [vp, P(vp), P(vp)]
    lib.shf_buffer_size.argtypes = [vp, P(sz), P(sz)]
    lib.shf_buffer_type.argtypes = [vp]
    lib.shf_buffer_type.restype = ctypes.c_ubyte
    lib.shf_run.argtypes = [vp, vp, _U32x2, _U32x2, _U32x2, vp, u32]
    lib.shf_run_batch.argtypes = [vp, P(vp), u32, _U32x2, _U32x2, _U32x2, vp, u32]
    lib.shf_run_device.argtypes = [vp, vp, u64, u32, _U32x2, _U32x2, _U32x2, vp, u32, vp]
    lib.shf_run_device_async.argtypes = [vp, vp, u64, u32, _U32x2, _U32x2, _U32x2, vp, u32, vp]
    lib.shf_buffer_wait.argtypes = [vp, P(u64)]
    lib.shf_run_neighbours.argtypes = [vp, P(vp), u32, _U32x2, _U32x2, vp, u32]
    lib.shf_run_multi.argtypes = [vp, P(vp), P(vp), u32, _U32x2, _U32x2, _U32x2, u32]
    lib.shf_run_neighbours_device.argtypes = [vp, P(vp), u32, _U32x2, _U32x2, vp, u32, vp]
    lib.shf_buffer_read_device.argtypes = [vp, P(vp), P(vp)]
    lib.shf_buffer_chunk_base.argtypes = [vp, P(vp), P(u32)]
    lib.shf_last_error.restype = ctypes.c_char_p
    lib.shf_stats_reset.restype = None
    lib.shf_stats_get.argtypes = [P(u64), P(u64), P(u64)]
    lib.shf_stats_get.restype = None
    lib.shf_buffer_last_plan.argtypes = [vp, P(u32), P(u32), P(u32), P(u32)]
    lib.shf_set_profiling.argtypes = [ctypes.c_int]
    lib.shf_set_profiling.restype = None
    lib.shf_buffer_phase_ms.argtypes = [vp, P(ctypes.c_float), u32]
    lib.shf_buffer_phase_history.argtypes = [vp, u32, P(ctypes.c_float), u32]
    lib.shf_heightfield_create.argtypes = [P(vp), vp, vp, u32, vp, vp, u32]
    lib.shf_heightfield_destroy.argtypes = [vp]
    lib.shf_heightfield_destroy.restype = None
    lib.shf_heightfield_run.argtypes = [vp, vp, u32, u32, vp, vp, vp]
    lib.shf_biome_factory_create.argtypes = [P(vp), vp, u32, u32, vp, u32, u64, u64, vp]
    lib.shf_biome_factory_destroy.argtypes = [vp]
    lib.shf_biome_factory_destroy.restype = None
    lib.shf_biome_factory_run.argtypes = [vp, vp, u32, u64, u32, vp, vp]
    _lib = lib
    return lib


def _raise(status: int) -> None:
    msg = (library().shf_last_error() or b"").decode("utf-8", "replace")
    if status == SHF_ERR_NUMERIC_DOMAIN:
        raise STPNumericDomainError(msg)
    if status == SHF_ERR_INVALID_ENUM:
        raise STPInvalidEnum(msg)
    if status == SHF_ERR_CUDA:
        raise STPCUDAError(msg)
    if status in (SHF_ERR_UNSUPPORTED, SHF_ERR_OFFSET_OVERFLOW):
        raise STPUnsupportedError(msg)
    raise ValueError(msg)


def _check(status: int) -> None:
    if status != SHF_OK:
        _raise(status)


@dataclass(frozen=True)
class STPNearestNeighbourInformation:
    """MapSize, ChunkNearestNeighbour, TotalMapSize as (x, y) pairs (STPNearestNeighbourInformation.hpp:13-25)."""

    MapSize: Tuple[int, int]
    ChunkNearestNeighbour: Tuple[int, int]
    TotalMapSize: Tuple[int, int]


@dataclass(frozen=True)
class STPSingleHistogram:
    """Non-owning view of a filter result (STPSingleHistogram.hpp:15-50): `Bin` is a structured array with fields
    Item / Weight, `HistogramStartOffset` has W*H+1 entries per chunk. Both alias the filter buffer's page-locked
    memory and are valid until the buffer is reused or destroyed. A fresh buffer reads as (None, None)."""

    Bin: Optional[np.ndarray]
    HistogramStartOffset: Optional[np.ndarray]


class DeviceArrayView:
    """A raw device pointer dressed as a CUDA array (``__cuda_array_interface__``), so that torch / cupy can wrap the
    device-resident result without a copy: ``torch.as_tensor(DeviceArrayView(ptr, n, "<i4"), device="cuda")``."""

    def __init__(self, ptr: int, count: int, typestr: str = "<i4"):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def set_profiling(enabled: bool) -> None:
    library().shf_set_profiling(1 if enabled else 0)


def stats_reset() -> None:
    library().shf_stats_reset()


def stats() -> Tuple[int, int, int]:
    """(kernel launches, host->device bytes, device->host bytes) issued by this thread since the last reset."""
    a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    library().shf_stats_get(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


class STPSingleHistogramFilter:
    """GPU single histogram filter with the reference's call signature (SHF.h:160-188)."""

    class STPFilterBuffer:
        """Opaque scratch + output memory of one filter execution at a time (SHF.h:43-123). Reuse it across calls."""

        class STPExecutionType(enum.IntEnum):
            Serial = 0x00
            Parallel = 0xFF

        def __init__(self, execution_type):
            self._h = ctypes.c_void_p()
            self._keep = None
            _check(library().shf_buffer_create(ctypes.byref(self._h), int(execution_type) & 0xFF))

        def close(self) -> None:
            if getattr(self, "_h", None):
                library().shf_buffer_destroy(self._h)
                self._h = ctypes.c_void_p()

        def __del__(self):
            try:
                self.close()
            except Exception:
                pass

        def readHistogram(self) -> STPSingleHistogram:
            bins_p, offs_p = ctypes.c_void_p(), ctypes.c_void_p()
            _check(library().shf_buffer_read(self._h, ctypes.byref(bins_p), ctypes.byref(offs_p)))
            n_bins, n_offs = self.size()
            if not offs_p.value:
                return STPSingleHistogram(None, None)
            offs = np.ctypeslib.as_array(ctypes.cast(offs_p, ctypes.POINTER(ctypes.c_uint32)), (n_offs,))
            if n_bins:
                raw = (ctypes.c_char * (n_bins * BIN_DTYPE.itemsize)).from_address(bins_p.value)
                bins = np.frombuffer(raw, dtype=BIN_DTYPE)
            else:
                bins = np.zeros(0, dtype=BIN_DTYPE)
            return STPSingleHistogram(bins, offs)

        def size(self) -> Tuple[int, int]:
            a, b = ctypes.c_size_t(), ctypes.c_size_t()
            _check(library().shf_buffer_size(self._h, ctypes.byref(a), ctypes.byref(b)))
            return a.value, b.value

        def type(self) -> "STPSingleHistogramFilter.STPFilterBuffer.STPExecutionType":
            return self.STPExecutionType(library().shf_buffer_type(self._h))

        # ---- additive (device-resident results) ----
        def readDevice(self) -> Tuple[int, int]:
            """(device pointer of the bins, device pointer of the offsets) of the last result."""
            bins_p, offs_p = ctypes.c_void_p(), ctypes.c_void_p()
            _check(library().shf_buffer_read_device(self._h, ctypes.byref(bins_p), ctypes.byref(offs_p)))
            return bins_p.value or 0, offs_p.value or 0

        def chunkBase(self) -> np.ndarray:
            """First-bin index of every chunk of the last (batch) result, n_chunks+1 entries."""
            p, n = ctypes.c_void_p(), ctypes.c_uint32()
            _check(library().shf_buffer_chunk_base(self._h, ctypes.byref(p), ctypes.byref(n)))
            if not p.value:
                return np.zeros(0, dtype=np.uint64)
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), (n.value + 1,)).copy()

        def wait(self) -> int:
            """Completes a pending runDeviceAsync (shf_buffer_wait); returns how many calls on this buffer ran ahead with
            a plan that did not fit and were repeated so far."""
            n = ctypes.c_uint64()
            _check(library().shf_buffer_wait(self._h, ctypes.byref(n)))
            return n.value

        def phaseMs(self, back: int = 0) -> dict:
            """Milliseconds of every kernel phase of the last call, or of the call `back` calls before it (the last 64
            profiled calls are kept; needs set_profiling(True) before the calls)."""
            ms = (ctypes.c_float * len(PHASES))()
            _check(library().shf_buffer_phase_history(self._h, back, ms, len(PHASES)))
            return dict(zip(PHASES, [float(v) for v in ms]))

        def lastPlan(self) -> dict:
            k, ty, nb, sm = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
            _check(library().shf_buffer_last_plan(self._h, ctypes.byref(k), ctypes.byref(ty), ctypes.byref(nb),
                                                  ctypes.byref(sm)))
            return {"k_sets": k.value, "rows_per_cta": ty.value, "biomes": nb.value, "smem_bytes": sm.value}

    def __init__(self, device: int = -1):
        self._h = ctypes.c_void_p()
        _check(library().shf_filter_create(ctypes.byref(self._h), device))

    def close(self) -> None:
        if getattr(self, "_h", None):
            library().shf_filter_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _host_map(samplemap, nn_info: STPNearestNeighbourInformation) -> np.ndarray:
        m = np.ascontiguousarray(samplemap, dtype=np.uint16)
        if m.size < nn_info.TotalMapSize[0] * nn_info.TotalMapSize[1]:
            raise ValueError("sample map smaller than TotalMapSize")
        return m

    def __call__(self, samplemap, nn_info: STPNearestNeighbourInformation, filter_buffer: "STPFilterBuffer",
                 radius: int) -> STPSingleHistogram:
        """operator()(samplemap, nn_info, filter_buffer, radius): host map in, page-locked host histogram out."""
        m = self._host_map(samplemap, nn_info)
        _check(library().shf_run(self._h, m.ctypes.data, _U32x2(*nn_info.MapSize),
                                 _U32x2(*nn_info.ChunkNearestNeighbour), _U32x2(*nn_info.TotalMapSize),
                                 filter_buffer._h, radius))
        return filter_buffer.readHistogram()

    def runBatch(self, samplemaps: Sequence, nn_info: STPNearestNeighbourInformation,
                 filter_buffer: "STPFilterBuffer", radius: int) -> STPSingleHistogram:
        """N independent neighbourhoods of equal geometry in one call (additive; shf_run_batch)."""
        maps = [self._host_map(m, nn_info) for m in samplemaps]
        arr = (ctypes.c_void_p * len(maps))(*[m.ctypes.data for m in maps])
        _check(library().shf_run_batch(self._h, arr, len(maps), _U32x2(*nn_info.MapSize),
                                       _U32x2(*nn_info.ChunkNearestNeighbour), _U32x2(*nn_info.TotalMapSize),
                                       filter_buffer._h, radius))
        return filter_buffer.readHistogram()

    def runMulti(self, samplemaps: Sequence, nn_info: STPNearestNeighbourInformation, filter_buffers: Sequence,
                 radius: int) -> list:
        """len(samplemaps) concurrent operator() calls served by one device pass (additive; shf_run_multi, SURVEY.md
        section 8 row f3): afterwards filter_buffers[i] reads as if self(samplemaps[i], nn_info, filter_buffers[i], radius)
        had been called. Returns the per-call histograms."""
        maps = [self._host_map(m, nn_info) for m in samplemaps]
        if len(maps) != len(filter_buffers):
            raise ValueError("one filter buffer per sample map")
        arr = (ctypes.c_void_p * len(maps))(*[m.ctypes.data for m in maps])
        bufs = (ctypes.c_void_p * len(maps))(*[b._h.value for b in filter_buffers])
        _check(library().shf_run_multi(self._h, arr, bufs, len(maps), _U32x2(*nn_info.MapSize),
                                       _U32x2(*nn_info.ChunkNearestNeighbour), _U32x2(*nn_info.TotalMapSize), radius))
        return [b.readHistogram() for b in filter_buffers]

    def runNeighbours(self, neighbour_maps: Sequence, nn_info: STPNearestNeighbourInformation,
                      filter_buffer: "STPFilterBuffer", radius: int) -> STPSingleHistogram:
        """The filter fed by the UNMERGED chunk maps (additive; shf_run_neighbours).
        neighbour_maps: for every neighbourhood its nn.x * nn.y chunk maps (MapSize.y x MapSize.x uint16 each) in
        local-index order (i % nn.x, i / nn.x), i.e. what STPNearestNeighbourTextureBuffer is constructed from; a flat
        sequence of n * nn.x * nn.y arrays. Only the centre chunk + halo travel to the device, no merged host buffer."""
        w, h = nn_info.MapSize
        per = nn_info.ChunkNearestNeighbour[0] * nn_info.ChunkNearestNeighbour[1]
        maps = [np.ascontiguousarray(m, dtype=np.uint16) for m in neighbour_maps]
        if not maps or len(maps) % per or any(m.size != w * h for m in maps):
            raise ValueError("neighbour_maps must hold n * nn.x * nn.y maps of MapSize samples each")
        arr = (ctypes.c_void_p * len(maps))(*[m.ctypes.data for m in maps])
        _check(library().shf_run_neighbours(self._h, arr, len(maps) // per, _U32x2(*nn_info.MapSize),
                                            _U32x2(*nn_info.ChunkNearestNeighbour), filter_buffer._h, radius))
        return filter_buffer.readHistogram()

    def runNeighboursDevice(self, neighbour_ptrs: Sequence[int], nn_info: STPNearestNeighbourInformation,
                            filter_buffer: "STPFilterBuffer", radius: int, stream: int = 0) -> None:
        """Same from raw pointers (host or device memory), result left in device memory (shf_run_neighbours_device)."""
        per = nn_info.ChunkNearestNeighbour[0] * nn_info.ChunkNearestNeighbour[1]
        if not neighbour_ptrs or len(neighbour_ptrs) % per:
            raise ValueError("neighbour_ptrs must hold n * nn.x * nn.y pointers")
        arr = (ctypes.c_void_p * len(neighbour_ptrs))(*neighbour_ptrs)
        _check(library().shf_run_neighbours_device(self._h, arr, len(neighbour_ptrs) // per, _U32x2(*nn_info.MapSize),
                                                   _U32x2(*nn_info.ChunkNearestNeighbour), filter_buffer._h, radius,
                                                   stream))

    def runDevice(self, device_ptr: int, chunk_stride: int, n_chunks: int, nn_info: STPNearestNeighbourInformation,
                  filter_buffer: "STPFilterBuffer", radius: int, stream: int = 0) -> None:
        """Device-resident variant (additive; shf_run_device): merged maps already in HBM, result stays in HBM."""
        _check(library().shf_run_device(self._h, device_ptr, chunk_stride, n_chunks, _U32x2(*nn_info.MapSize),
                                        _U32x2(*nn_info.ChunkNearestNeighbour), _U32x2(*nn_info.TotalMapSize),
                                        filter_buffer._h, radius, stream))


    def runDeviceAsync(self, device_ptr: int, chunk_stride: int, n_chunks: int, nn_info: STPNearestNeighbourInformation,
                       filter_buffer: "STPFilterBuffer", radius: int, stream: int = 0) -> None:
        """runDevice without host synchronisation when the call can run ahead (shf_run_device_async); the first query of
        the result (filter_buffer.wait(), size(), readDevice(), ...) completes it."""
        _check(library().shf_run_device_async(self._h, device_ptr, chunk_stride, n_chunks, _U32x2(*nn_info.MapSize),
                                              _U32x2(*nn_info.ChunkNearestNeighbour), _U32x2(*nn_info.TotalMapSize),
                                              filter_buffer._h, radius, stream))


STPFilterBuffer = STPSingleHistogramFilter.STPFilterBuffer


class STPMultiBiomeHeightfield:
    """Device-side consumer of a filter result: the multi-biome heightfield kernel of
    SuperDemo+/Script/STPMultiHeightGenerator.cu:35-71 (generateMultiBiomeHeightmap) reading the histogram where the
    filter left it in device memory (additive; shf_heightfield_*).

    table: BIOME_PROPERTY_DTYPE array indexed by sample value (the reference's __constant__ BiomeTable);
    permutation: 512 uint8; gradient: [N, 2] float32 (STPPermutationGenerator's tables, taken as inputs)."""

    def __init__(self, filt: STPSingleHistogramFilter, table, permutation, gradient):
        table = np.ascontiguousarray(table, dtype=BIOME_PROPERTY_DTYPE)
        perm = np.ascontiguousarray(permutation, dtype=np.uint8)
        grad = np.ascontiguousarray(gradient, dtype=np.float32).reshape(-1)
        if perm.size != 512 or grad.size % 2:
            raise ValueError("permutation must hold 512 bytes and gradient N x 2 floats")
        self._h = ctypes.c_void_p()
        _check(library().shf_heightfield_create(ctypes.byref(self._h), filt._h, table.ctypes.data, len(table),
                                                perm.ctypes.data, grad.ctypes.data, grad.size // 2))

    def close(self) -> None:
        if getattr(self, "_h", None):
            library().shf_heightfield_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, filter_buffer: STPFilterBuffer, first_chunk: int, n_chunks: int, offsets_xy,
                 height_device_ptr: int, stream: int = 0) -> None:
        """Heightfields of chunks [first_chunk, first_chunk + n_chunks) of the buffer's last result into device memory
        (n_chunks x H x W float32); enqueued on `stream`."""
        off = np.ascontiguousarray(offsets_xy, dtype=np.float32).reshape(-1)
        if off.size != 2 * n_chunks:
            raise ValueError("offsets_xy must hold one (x, y) pair per chunk")
        _check(library().shf_heightfield_run(self._h, filter_buffer._h, first_chunk, n_chunks, off.ctypes.data,
                                             height_device_ptr, stream))


# shf_biome_layer / shf_biome_ids of include/shf_b200.h
BIOME_LAYER_DTYPE = np.dtype([("kind", "<u4"), ("parent", "<u4"), ("salt", "<u8")])
BIOME_IDS_DTYPE = np.dtype([(n, "<u2") for n in ("ocean", "plains", "forest", "frozen_ocean", "warm_ocean", "lukewarm_ocean",
                                                  "cold_ocean")])


class STPLayerKind(enum.IntEnum):
    """The demo's layer classes this producer implements (SuperDemo+/World/Layers/)."""

    Continent = 0
    ScaleNormal = 1
    ScaleFuzzy = 2
    Land = 3
    Island = 4
    Voronoi = 5


class STPBiomeFactory:
    """Biome-map producer on the device (additive; shf_biome_factory_*) with the call shape of
    STPDiversity::STPBiomeFactory (STPBiomeFactory.h:21-73): constructed with the biome map dimension, called with a map
    and the world offset of its first cell. The layer chain the reference builds in supply() out of virtual STPLayer
    objects is given as data: ``layers`` = sequence of (STPLayerKind, salt) or (STPLayerKind, salt, ascendant index), in
    construction order, the last one being the root; a missing ascendant index means "the layer before"."""

    def __init__(self, filt: "STPSingleHistogramFilter", dimension: Tuple[int, int], layers: Sequence, global_seed: int,
                 ids, voronoi_seed: Optional[int] = None):
        arr = np.zeros(len(layers), dtype=BIOME_LAYER_DTYPE)
        for i, ly in enumerate(layers):
            arr[i] = (int(ly[0]), int(ly[2]) if len(ly) > 2 else max(i - 1, 0), int(ly[1]))
        idv = np.zeros(1, dtype=BIOME_IDS_DTYPE)
        idv[0] = tuple(int(v) for v in ids)
        self.BiomeDimension = (int(dimension[0]), int(dimension[1]))
        self._h = ctypes.c_void_p()
        # std::hash<STPSeed_t> is the identity with libstdc++ (STPVoronoiLayer.h:52)
        vs = global_seed if voronoi_seed is None else voronoi_seed
        _check(library().shf_biome_factory_create(ctypes.byref(self._h), filt._h, self.BiomeDimension[0],
                                                  self.BiomeDimension[1], arr.ctypes.data, len(arr),
                                                  global_seed & 0xFFFFFFFFFFFFFFFF, vs & 0xFFFFFFFFFFFFFFFF,
                                                  idv.ctypes.data))

    def close(self) -> None:
        if getattr(self, "_h", None):
            library().shf_biome_factory_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, biomemap_device_ptr: int, offsets_xz, row_stride: int = 0, map_stride: int = 0,
                 stream: int = 0) -> None:
        """operator()(biomemap, offset) for one or many offsets: map i is written to device memory at
        biomemap_device_ptr + 2 * i * map_stride bytes (default: maps packed back to back), enqueued on `stream`."""
        off = np.ascontiguousarray(offsets_xz, dtype=np.int32).reshape(-1, 2)
        stride = map_stride or (row_stride or self.BiomeDimension[0]) * self.BiomeDimension[1]
        _check(library().shf_biome_factory_run(self._h, biomemap_device_ptr, row_stride, stride, len(off),
                                               off.ctypes.data, stream))


class STPLayerChainBuilder(STPBiomeFactory):
    """The demo's chain (STPDemo::STPLayerChainBuilder, SuperDemo+/World/Layers/STPAllLayers.cpp:61-109): continent ->
    fuzzy scale -> land -> scale -> 3 x land -> island -> 3 x scale -> 3 x Voronoi, with the demo's salts."""

    CHAIN = (
        (STPLayerKind.Continent, 23457829), (STPLayerKind.ScaleFuzzy, 875944), (STPLayerKind.Land, 5748329),
        (STPLayerKind.ScaleNormal, 8947358941), (STPLayerKind.Land, 361249673), (STPLayerKind.Land, 8769575),
        (STPLayerKind.Land, 43562783426564), (STPLayerKind.Island, 74368),
        (STPLayerKind.ScaleNormal, 1), (STPLayerKind.ScaleNormal, 2), (STPLayerKind.ScaleNormal, 3),
        (STPLayerKind.Voronoi, 4), (STPLayerKind.Voronoi, 5), (STPLayerKind.Voronoi, 6),
    )
    # SuperDemo+/Biome.ini: ocean 0, plains 1, forest 3; the other shallow oceans of the registry are never given an id
    IDS = (0, 1, 3, 0, 0, 0, 0)

    def __init__(self, filt: "STPSingleHistogramFilter", dimension: Tuple[int, int], global_seed: int, ids=None):
        super().__init__(filt, dimension, self.CHAIN, global_seed, ids or self.IDS)
        self.GlobalSeed = global_seed
