"""ctypes front-ends for the two CPU oracles plus an independent closed-form checker.

* ``run_port``       -- oracle/shf_oracle.c, the committed C restatement (always available once built).
* ``run_reference``  -- oracle/_ref/libshf_ref.so, the reference's own STPSingleHistogramFilter compiled from
                        /root/reference by oracle/Makefile (available where it was prebuilt).
* ``closed_form``    -- pure numpy/Python evaluation of the order rule in SURVEY.md Appendix A.5; slow, small cases.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "libshf_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libshf_ref.so")

# STPSingleHistogram::STPBin = {uint16 Item; float Weight} (8 bytes, 2 padding bytes)
BIN_DTYPE = np.dtype({"names": ["item", "weight"], "formats": ["<u2", "<f4"], "offsets": [0, 4], "itemsize": 8})

_U32x2 = ctypes.c_uint32 * 2


class OracleError(RuntimeError):
    """Raised with ``status`` 1 for the reference's STPNumericDomainError, 2 for STPInvalidEnum."""

    def __init__(self, status: int, what: str):
        super().__init__(f"{what}: status {status}")
        self.status = status


def build(verbose: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference itself."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


_port = None
_ref = None


def _port_lib():
    global _port
    if _port is None:
        if not os.path.exists(_PORT_SO):
            build()
        lib = ctypes.CDLL(_PORT_SO)
        lib.shf_oracle_run.restype = ctypes.c_int
        lib.shf_oracle_run.argtypes = [ctypes.c_void_p] + [ctypes.c_uint32] * 6 + [
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]
        lib.shf_oracle_free.argtypes = [ctypes.c_void_p]
        assert lib.shf_oracle_bin_stride() == BIN_DTYPE.itemsize
        _port = lib
    return _port


def have_reference() -> bool:
    return os.path.exists(_REF_SO)


def _ref_lib():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(_REF_SO)
        lib.ref_shf_create.restype = ctypes.c_void_p
        lib.ref_shf_create.argtypes = [ctypes.c_ubyte, ctypes.POINTER(ctypes.c_int)]
        lib.ref_shf_destroy.argtypes = [ctypes.c_void_p]
        lib.ref_shf_run.restype = ctypes.c_int
        lib.ref_shf_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _U32x2, _U32x2, _U32x2, ctypes.c_uint32,
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        lib.ref_shf_type.restype = ctypes.c_ubyte
        lib.ref_shf_type.argtypes = [ctypes.c_void_p]
        lib.ref_shf_size.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        assert lib.ref_shf_bin_stride() == BIN_DTYPE.itemsize
        _ref = lib
    return _ref


def reference_pinned_active() -> bool:
    """True when the reference build's output buffers are genuinely page-locked (a CUDA driver is present)."""
    return bool(_ref_lib().shf_ref_pinned_active())


def _check_map(sample_map: np.ndarray, total: Tuple[int, int]) -> np.ndarray:
    m = np.ascontiguousarray(sample_map, dtype=np.uint16)
    assert m.ndim == 2 and m.shape == (total[1], total[0]), (m.shape, total)
    return m


def _copy_out(bins_p, offs_p, n_bins, n_offs):
    offsets = np.ctypeslib.as_array(ctypes.cast(offs_p, ctypes.POINTER(ctypes.c_uint32)), (n_offs,)).copy()
    if n_bins:
        raw = (ctypes.c_char * (n_bins * BIN_DTYPE.itemsize)).from_address(bins_p)
        bins = np.frombuffer(raw, dtype=BIN_DTYPE).copy()
    else:
        bins = np.zeros(0, dtype=BIN_DTYPE)
    return bins["item"].copy(), bins["weight"].copy(), offsets


def run_port(sample_map, map_size, nn, radius, total=None):
    """C restatement. Returns (items u16[n], weights f32[n], offsets u32[W*H+1])."""
    w, h = map_size
    total = total or (w * nn[0], h * nn[1])
    m = _check_map(sample_map, total)
    lib = _port_lib()
    bins_p, offs_p, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64()
    st = lib.shf_oracle_run(m.ctypes.data, w, h, nn[0], nn[1], total[0], radius,
                            ctypes.byref(bins_p), ctypes.byref(offs_p), ctypes.byref(n))
    if st != 0:
        raise OracleError(st, "shf_oracle_run")
    try:
        return _copy_out(bins_p.value, offs_p.value, n.value, w * h + 1)
    finally:
        lib.shf_oracle_free(bins_p)
        lib.shf_oracle_free(offs_p)


class ReferenceSession:
    """One reference filter object + one STPFilterBuffer (exec_type 0x00 serial / 0xFF parallel), reusable."""

    def __init__(self, exec_type: int = 0xFF):
        lib = _ref_lib()
        st = ctypes.c_int()
        self._h = lib.ref_shf_create(exec_type, ctypes.byref(st))
        if not self._h:
            raise OracleError(st.value, "ref_shf_create")

    def close(self):
        if self._h:
            _ref_lib().ref_shf_destroy(self._h)
            self._h = None

    __del__ = close

    def type(self) -> int:
        return _ref_lib().ref_shf_type(self._h)

    def size(self):
        nb, no = ctypes.c_uint64(), ctypes.c_uint64()
        null_mask = _ref_lib().ref_shf_size(self._h, ctypes.byref(nb), ctypes.byref(no))
        return nb.value, no.value, null_mask

    def run_raw(self, m: np.ndarray, map_size, nn, total, radius):
        """Run without copying the result out (timing loops). Returns (n_bins, n_offsets)."""
        bins_p, offs_p = ctypes.c_void_p(), ctypes.c_void_p()
        nb, no = ctypes.c_uint64(), ctypes.c_uint64()
        st = _ref_lib().ref_shf_run(self._h, m.ctypes.data, _U32x2(*map_size), _U32x2(*nn), _U32x2(*total), radius,
                                    ctypes.byref(bins_p), ctypes.byref(offs_p), ctypes.byref(nb), ctypes.byref(no))
        if st != 0:
            raise OracleError(st, "ref_shf_run")
        self._last = (bins_p.value, offs_p.value, nb.value, no.value)
        return nb.value, no.value

    def run(self, sample_map, map_size, nn, radius, total=None):
        w, h = map_size
        total = total or (w * nn[0], h * nn[1])
        m = _check_map(sample_map, total)
        self.run_raw(m, map_size, nn, total, radius)
        return _copy_out(*self._last)


def run_reference(sample_map, map_size, nn, radius, exec_type: int = 0xFF, total=None):
    """The reference's own filter. Returns (items, weights, offsets)."""
    s = ReferenceSession(exec_type)
    try:
        return s.run(sample_map, map_size, nn, radius, total)
    finally:
        s.close()


def closed_form(sample_map, map_size, nn, radius):
    """Independent evaluation of SURVEY.md Appendix A.5 (no accumulator emulation): per pixel, the present samples
    sorted by (first column of the sample's current horizontal presence chain, first row of the vertical presence
    chain in that column). O(W*H*(W+H)) Python -- small maps only."""
    w, h = map_size
    r = radius
    if r == 0 or r % 2 or r > w * (nn[0] // 2) or r > h * (nn[1] // 2):
        raise OracleError(1, "closed_form")
    span = 2 * r + 1
    sx, sy = w * (nn[0] // 2), h * (nn[1] // 2)
    m = np.asarray(sample_map, dtype=np.int64)
    reg = m[sy - r: sy + h + r, sx - r: sx + w + r]  # halo-extended region, [rho, c]
    hh, ww = reg.shape
    # vertical: cnt_v[c][y] dict sample->count and key_v[c][y] dict sample->chain start row
    cnt_v = [[None] * h for _ in range(ww)]
    key_v = [[None] * h for _ in range(ww)]
    for c in range(ww):
        last, start = {}, {}
        for rho in range(hh):
            s = int(reg[rho, c])
            if s not in last or rho - last[s] > span:
                start[s] = rho
            last[s] = rho
            y = rho - 2 * r
            if y >= 0:
                vals, counts = np.unique(reg[y: y + span, c], return_counts=True)
                cnt_v[c][y] = {int(v): int(k) for v, k in zip(vals, counts)}
                key_v[c][y] = {int(v): start[int(v)] for v in vals}
    items, weights, offsets = [], [], [0]
    inv = np.float32(1.0) / np.float32(span * span)
    for y in range(h):
        last_h, start_h, start_key = {}, {}, {}
        for x in range(w):
            lo = 0 if x == 0 else x + 2 * r
            for c in range(lo, x + 2 * r + 1):
                for s in cnt_v[c][y]:
                    if s not in last_h or c - last_h[s] > span:
                        start_h[s] = c
                        start_key[s] = key_v[c][y][s]
                    last_h[s] = c
            tot = {}
            for c in range(x, x + span):
                for s, k in cnt_v[c][y].items():
                    tot[s] = tot.get(s, 0) + k
            order = sorted(tot, key=lambda s: (start_h[s], start_key[s]))
            for s in order:
                items.append(s)
                weights.append(np.float32(tot[s]) * inv)
            offsets.append(len(items))
    return (np.array(items, dtype=np.uint16), np.array(weights, dtype=np.float32),
            np.array(offsets, dtype=np.uint32))


# ----------------------------------------------------------------------------------------------------------------------
# heightfield consumer
# ----------------------------------------------------------------------------------------------------------------------
_HF_PORT_SO = os.path.join(_HERE, "libshf_heightfield_oracle.so")
_HF_REF_SO = os.path.join(_HERE, "_ref", "libshf_ref_height.so")
# STPDemo::STPBiomeProperty (SuperDemo+/World/Biomes/STPBiomeProperty.hpp:10-27)
BIOME_PROPERTY_DTYPE = np.dtype([("Scale", "<f4"), ("Octave", "<u4"), ("Persistence", "<f4"), ("Lacunarity", "<f4"),
                                 ("Depth", "<f4"), ("Variation", "<f4")])
_hf_port = None
_hf_ref = None


def _pack_bins(items, weights):
    bins = np.zeros(len(items), dtype=BIN_DTYPE)
    bins["item"] = items
    bins["weight"] = weights
    return bins


def _hf_args(items, weights, offsets, map_size, table, permutation, gradient, offset_xy):
    w, h = map_size
    bins = _pack_bins(items, weights)
    offs = np.ascontiguousarray(offsets, dtype=np.uint32)
    table = np.ascontiguousarray(table, dtype=BIOME_PROPERTY_DTYPE)
    perm = np.ascontiguousarray(permutation, dtype=np.uint8)
    grad = np.ascontiguousarray(gradient, dtype=np.float32).reshape(-1)
    assert perm.size == 512 and offs.size == w * h + 1 and grad.size % 2 == 0
    return bins, offs, table, perm, grad, np.zeros(w * h, dtype=np.float32)


def heightfield_port(items, weights, offsets, map_size, table, permutation, gradient, offset_xy, pixels=None) -> np.ndarray:
    """CPU restatement (oracle/shf_heightfield_oracle.c) of the multi-biome heightfield of one chunk, float32 [H, W].
    pixels = (begin, end) restricts the evaluation to that pixel range (the rest stays 0)."""
    global _hf_port
    if _hf_port is None:
        if not os.path.exists(_HF_PORT_SO):
            build()
        _hf_port = ctypes.CDLL(_HF_PORT_SO)
        _hf_port.shf_heightfield_oracle.restype = None
    bins, offs, table, perm, grad, out = _hf_args(items, weights, offsets, map_size, table, permutation, gradient, offset_xy)
    vp = ctypes.c_void_p
    _hf_port.shf_heightfield_oracle(vp(bins.ctypes.data), vp(offs.ctypes.data), ctypes.c_uint32(map_size[0]),
                                    ctypes.c_uint32(map_size[1]), vp(table.ctypes.data), ctypes.c_uint32(len(table)),
                                    vp(perm.ctypes.data), vp(grad.ctypes.data), ctypes.c_uint32(grad.size // 2),
                                    ctypes.c_float(offset_xy[0]), ctypes.c_float(offset_xy[1]), vp(out.ctypes.data),
                                    ctypes.c_uint32(pixels[0] if pixels else 0),
                                    ctypes.c_uint32(pixels[1] if pixels else map_size[0] * map_size[1]))
    return out.reshape(map_size[1], map_size[0])


def have_height_reference() -> bool:
    return os.path.exists(_HF_REF_SO)


def heightfield_reference(items, weights, offsets, map_size, table, permutation, gradient, offset_xy) -> np.ndarray:
    """The reference's own device code (oracle/_ref/libshf_ref_height.so); needs a GPU."""
    global _hf_ref
    if _hf_ref is None:
        _hf_ref = ctypes.CDLL(_HF_REF_SO)
        _hf_ref.ref_heightfield_run.restype = ctypes.c_int
    bins, offs, table, perm, grad, out = _hf_args(items, weights, offsets, map_size, table, permutation, gradient, offset_xy)
    vp = ctypes.c_void_p
    st = _hf_ref.ref_heightfield_run(vp(bins.ctypes.data), vp(offs.ctypes.data), ctypes.c_uint64(len(bins)),
                                     ctypes.c_uint32(map_size[0]), ctypes.c_uint32(map_size[1]), vp(table.ctypes.data),
                                     ctypes.c_uint32(len(table)), vp(perm.ctypes.data), vp(grad.ctypes.data),
                                     ctypes.c_uint32(grad.size // 2), ctypes.c_float(offset_xy[0]),
                                     ctypes.c_float(offset_xy[1]), vp(out.ctypes.data))
    if st != 0:
        raise OracleError(st, "reference heightfield (CUDA)")
    return out.reshape(map_size[1], map_size[0])
