// Biome-map producer on the device with the call shape of
// SuperTerrainPlus::STPDiversity::STPBiomeFactory (SuperTerrain+/SuperTerrain+/Public/SuperTerrain+/World/Diversity/
// STPBiomeFactory.h:21-73): constructed with the biome map dimension, called with a map and the world offset of its first
// cell (STPBiomeFactory.cpp:24-42). Differences, both forced by where the code runs:
//  * `biomemap` is DEVICE memory (the point: the sample map the single histogram filter reads never exists on the host);
//  * the layer tree is not a `supply()` override building virtual STPLayer objects but a table of layer kinds and salts
//    (shf_biome_layer, include/shf_b200.h), in construction order, every layer naming its ascendant by index, the last
//    layer being the root -- for the demo's chain exactly the member list of STPLayerPipeline
//    (SuperDemo+/World/Layers/STPAllLayers.cpp:61-109).
// Layers are pure functions of the world coordinate (the reference's per-layer cache only saves recomputation), so the
// result is bit-identical to the CPU factory's for the same chain, seed and ids.
#pragma once
#ifndef _STP_BIOME_FACTORY_DEVICE_H_
#define _STP_BIOME_FACTORY_DEVICE_H_

#include <SuperAlgorithm+Host/STPSingleHistogramFilter.h>
#include <SuperTerrain+/World/STPWorldMapPixelFormat.hpp>

#include <shf_b200.h>

#include <glm/vec2.hpp>

#include <cstdint>

namespace SuperTerrainPlus::STPDiversity {

	class STP_ALGORITHM_HOST_API STPBiomeFactoryDevice {
	private:

		shf_biome_factory* Factory;

	public:

		//Specify the dimension of the generated biome map
		const glm::uvec2 BiomeDimension;

		//filter: fixes the device; layer / layer_count: the chain; voronoi_seed: std::hash<STPSeed_t>{}(global_seed) of the
		//toolchain the CPU factory is built with (STPVoronoiLayer.h:52; the identity with libstdc++)
		STPBiomeFactoryDevice(STPAlgorithm::STPSingleHistogramFilter& filter, glm::uvec2 dimension, const shf_biome_layer* layer,
			unsigned int layer_count, STPSeed_t global_seed, const shf_biome_ids& ids, STPSeed_t voronoi_seed);

		STPBiomeFactoryDevice(const STPBiomeFactoryDevice&) = delete;

		STPBiomeFactoryDevice(STPBiomeFactoryDevice&&) = delete;

		STPBiomeFactoryDevice& operator=(const STPBiomeFactoryDevice&) = delete;

		STPBiomeFactoryDevice& operator=(STPBiomeFactoryDevice&&) = delete;

		~STPBiomeFactoryDevice();

		//STPBiomeFactory::operator()(biomemap, offset) for `map_count` maps at once: map i starts at
		//biomemap_device + i * map_stride, its row z at + z * row_stride (0 = BiomeDimension.x), its first cell is world
		//coordinate (offset_xz[2i], offset_xz[2i+1]). Enqueued on `stream` (cudaStream_t); returns without waiting.
		void operator()(STPSample_t* biomemap_device, const int* offset_xz, unsigned int map_count = 1u,
			unsigned int row_stride = 0u, std::uint64_t map_stride = 0u, void* stream = nullptr);

	};

}
#endif//_STP_BIOME_FACTORY_DEVICE_H_
