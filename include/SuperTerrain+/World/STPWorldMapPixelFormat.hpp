// Standalone stand-in for SuperTerrain+/SuperTerrain+/Public/SuperTerrain+/World/STPWorldMapPixelFormat.hpp:12 --
// only the types this path touches. Inside the reference tree the reference's own header is
// used instead (same include path, same typedef).
#pragma once
#include <cstdint>

namespace SuperTerrainPlus {
	//biomemap sample, 16-bit unsigned (cuda::std::uint16_t in the reference)
	typedef std::uint16_t STPSample_t;
	//seed of the layer random number generators (STPWorldMapPixelFormat.hpp:21)
	typedef std::uint64_t STPSeed_t;
}
