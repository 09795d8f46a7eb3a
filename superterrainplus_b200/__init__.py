"""superterrainplus_b200 -- B200-native single histogram filter behind SuperTerrain+'s own interface.

Python mirror of the reference's C++ API for this path (names, argument meaning and error behaviour follow
SuperTerrain+/SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPSingleHistogramFilter.h:36-192 and
STPSingleHistogram.hpp:15-50), sitting on the C ABI of include/shf_b200.h through ctypes. The C++ mirror of the same
interface is include/SuperAlgorithm+Host/STPSingleHistogramFilter.h.

There is no CPU path here: importing works without a GPU (so that symbols can be checked), every compute call needs
the CUDA library and a device and raises otherwise.
"""
from .api import (  # noqa: F401
    BIN_DTYPE,
    BIOME_PROPERTY_DTYPE,
    STPBiomeFactory,
    STPCUDAError,
    STPInvalidEnum,
    STPLayerChainBuilder,
    STPLayerKind,
    STPMultiBiomeHeightfield,
    STPNearestNeighbourInformation,
    STPNumericDomainError,
    STPSingleHistogram,
    STPSingleHistogramFilter,
    STPUnsupportedError,
    library,
    library_path,
)

__all__ = [
    "BIN_DTYPE", "BIOME_PROPERTY_DTYPE", "STPBiomeFactory", "STPLayerChainBuilder", "STPLayerKind", "STPCUDAError", "STPMultiBiomeHeightfield", "STPInvalidEnum", "STPNearestNeighbourInformation", "STPNumericDomainError",
    "STPSingleHistogram", "STPSingleHistogramFilter", "STPUnsupportedError", "library", "library_path",
]
