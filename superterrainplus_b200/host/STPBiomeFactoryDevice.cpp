// Host side of STPBiomeFactoryDevice: forwards to the C ABI (shf_biome_factory_*, include/shf_b200.h) and turns status
// codes into the exception types of the reference (STPBiomeFactory.cpp:20-22 throws STPNumericDomainError for an empty
// dimension).
#include <SuperTerrain+/World/Diversity/STPBiomeFactoryDevice.h>
#include <SuperTerrain+/Exception/STPFundamentalException.h>

#include <string>

using namespace SuperTerrainPlus::STPDiversity;
namespace STPException = SuperTerrainPlus::STPException;

namespace SuperTerrainPlus::STPAlgorithm {
	//the filter's C handle (defined next to the filter class)
	shf_filter* STPFilterHandle(STPSingleHistogramFilter&);
}

static void check(const int status) {
	if (status == SHF_OK) {
		return;
	}
	const std::string message = shf_last_error();
	const size_t colon = message.find(": ");
	switch (status) {
	case SHF_ERR_NUMERIC_DOMAIN:
		throw STPException::STPNumericDomainError(colon == std::string::npos ? "" : message.substr(0u, colon).c_str(),
			colon == std::string::npos ? message : message.substr(colon + 2u));
	case SHF_ERR_INVALID_ENUM: throw STPException::STPInvalidEnum(message);
	case SHF_ERR_CUDA: throw STPException::STPCUDAError(message);
	default: throw STPException::STPUnsupportedOperation(message);
	}
}

STPBiomeFactoryDevice::STPBiomeFactoryDevice(STPAlgorithm::STPSingleHistogramFilter& filter, const glm::uvec2 dimension,
	const shf_biome_layer* const layer, const unsigned int layer_count, const STPSeed_t global_seed, const shf_biome_ids& ids,
	const STPSeed_t voronoi_seed) : Factory(nullptr), BiomeDimension(dimension) {
	check(shf_biome_factory_create(&this->Factory, STPAlgorithm::STPFilterHandle(filter), dimension.x, dimension.y, layer,
		layer_count, global_seed, voronoi_seed, &ids));
}

STPBiomeFactoryDevice::~STPBiomeFactoryDevice() {
	shf_biome_factory_destroy(this->Factory);
}

void STPBiomeFactoryDevice::operator()(STPSample_t* const biomemap_device, const int* const offset_xz, const unsigned int map_count,
	const unsigned int row_stride, const std::uint64_t map_stride, void* const stream) {
	const std::uint64_t stride = map_stride ? map_stride
		: static_cast<std::uint64_t>(row_stride ? row_stride : this->BiomeDimension.x) * this->BiomeDimension.y;
	check(shf_biome_factory_run(this->Factory, biomemap_device, row_stride, stride, map_count, offset_xz, stream));
}
