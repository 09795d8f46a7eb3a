"""TEST INFRASTRUCTURE ONLY: ctypes front-ends of the two CPU checkers of the device biome-map producer.

* ``biome_port``       -- oracle/biome_oracle.c, the committed C restatement of STPBiomeFactory + STPLayer + the demo layers.
* ``biome_reference``  -- oracle/_ref/libbiome_ref.so, the reference's own STPLayerChainBuilder (SuperDemo+/World/Layers/
                          STPAllLayers.cpp:61-109) compiled from /root/reference by oracle/Makefile.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "libbiome_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libbiome_ref.so")

CONTINENT, SCALE_NORMAL, SCALE_FUZZY, LAND, ISLAND, VORONOI = range(6)

LAYER_DTYPE = np.dtype([("kind", "<u4"), ("parent", "<u4"), ("salt", "<u8")])

# The demo's chain, STPAllLayers.cpp:61-109: (kind, salt), every layer's ascendant is the one before it
DEMO_CHAIN = [
    (CONTINENT, 23457829), (SCALE_FUZZY, 875944), (LAND, 5748329),
    (SCALE_NORMAL, 8947358941), (LAND, 361249673), (LAND, 8769575), (LAND, 43562783426564), (ISLAND, 74368),
    (SCALE_NORMAL, 1), (SCALE_NORMAL, 2), (SCALE_NORMAL, 3), (VORONOI, 4), (VORONOI, 5), (VORONOI, 6),
]
# SuperDemo+/Biome.ini: ocean 0, plains 1, forest 3; the registry's other shallow oceans are never given an id and keep
# the zero of their static storage (STPBiomeRegistry.cpp:9-46)
DEMO_IDS = (0, 1, 3, 0, 0, 0, 0)   # Ocean, Plains, Forest, FrozenOcean, WarmOcean, LukewarmOcean, ColdOcean


def chain_array(chain=DEMO_CHAIN) -> np.ndarray:
    a = np.zeros(len(chain), dtype=LAYER_DTYPE)
    for i, (kind, salt) in enumerate(chain):
        a[i] = (kind, max(i - 1, 0), salt)
    return a


_port = None
_ref = None


def biome_port(width, height, offset, seed, chain=None, ids=DEMO_IDS, voronoi_seed=None) -> np.ndarray:
    """Biome map [height, width] uint16 of the region starting at world coordinate `offset` = (x, z)."""
    global _port
    if _port is None:
        if not os.path.exists(_PORT_SO):
            from .pyoracle import build

            build()
        _port = ctypes.CDLL(_PORT_SO)
        _port.biome_oracle_run.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                           ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                           ctypes.c_uint32, ctypes.c_void_p]
    layers = chain_array() if chain is None else np.ascontiguousarray(chain, dtype=LAYER_DTYPE)
    idv = np.asarray(ids, dtype=np.uint16)
    out = np.zeros((height, width), dtype=np.uint16)
    # libstdc++'s std::hash<unsigned long long> is the identity (STPVoronoiLayer.h:52)
    vs = seed if voronoi_seed is None else voronoi_seed
    st = _port.biome_oracle_run(layers.ctypes.data, len(layers), seed, vs, idv.ctypes.data, offset[0], offset[1], width,
                                height, out.ctypes.data)
    if st != 0:
        raise RuntimeError(f"biome_oracle_run: status {st}")
    return out


def have_biome_reference() -> bool:
    return os.path.exists(_REF_SO)


def biome_reference(width, height, offset, seed, ids=DEMO_IDS) -> np.ndarray:
    """The reference's own demo chain (fixed layer list) on the same region."""
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF_SO)
        _ref.ref_biome_create.restype = ctypes.c_void_p
        _ref.ref_biome_create.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
        _ref.ref_biome_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
        _ref.ref_biome_destroy.argtypes = [ctypes.c_void_p]
        _ref.ref_biome_set_ids.argtypes = [ctypes.c_void_p]
    idv = np.asarray(ids, dtype=np.uint16)
    _ref.ref_biome_set_ids(idv.ctypes.data)
    f = _ref.ref_biome_create(width, height, seed)
    if not f:
        raise RuntimeError("ref_biome_create failed")
    out = np.zeros((height, width), dtype=np.uint16)
    try:
        if _ref.ref_biome_run(f, out.ctypes.data, offset[0], offset[1]) != 0:
            raise RuntimeError("ref_biome_run failed")
    finally:
        _ref.ref_biome_destroy(f)
    return out
