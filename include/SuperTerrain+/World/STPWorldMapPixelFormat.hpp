// Standalone stand-in for SuperTerrain+/SuperTerrain+/Public/SuperTerrain+/World/STPWorldMapPixelFormat.hpp:12 --
// only the pixel type the single histogram filter touches. Inside the reference tree the reference's own header is
// used instead (same include path, same typedef).
#pragma once
#include <cstdint>

namespace SuperTerrainPlus {
	//biomemap sample, 16-bit unsigned (cuda::std::uint16_t in the reference)
	typedef std::uint16_t STPSample_t;
}
