// Same type as SuperTerrain+/SuperTerrain+/Public/SuperTerrain+/World/Chunk/STPNearestNeighbourInformation.hpp:13-25
// (three glm::uvec2, aggregate-initialisable); standalone copy of the interface for builds outside the reference tree.
#pragma once
#include <glm/vec2.hpp>

namespace SuperTerrainPlus {

	struct STPNearestNeighbourInformation {
	public:

		//dimension of the map of one chunk
		glm::uvec2 MapSize;
		//number of chunks in the neighbourhood (centre chunk included), per axis
		glm::uvec2 ChunkNearestNeighbour;
		//dimension of the merged map = MapSize * ChunkNearestNeighbour; .x is the row stride of the sample map
		glm::uvec2 TotalMapSize;

	};

}
