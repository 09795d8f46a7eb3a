// Host side of the drop-in class: forwards every member to the C ABI (include/shf_b200.h) and turns status codes back
// into the exception types the reference throws on this path (SuperAlgorithm+/Host/Private/STPSingleHistogramFilter.cpp
// :721 STPInvalidEnum, :874,879-880 STPNumericDomainError; STP_CHECK_CUDA -> STPCUDAError).
#include <SuperAlgorithm+Host/STPSingleHistogramFilter.h>
#include <SuperTerrain+/Exception/STPFundamentalException.h>

#include <shf_b200.h>

#include <string>
#include <vector>

using namespace SuperTerrainPlus::STPAlgorithm;
namespace STPException = SuperTerrainPlus::STPException;

using STPFilterBuffer = STPSingleHistogramFilter::STPFilterBuffer;

static_assert(sizeof(STPSingleHistogram::STPBin) == sizeof(shf_bin) && alignof(STPSingleHistogram::STPBin) == alignof(shf_bin),
	"STPBin and shf_bin must be layout-identical");

namespace {

	[[noreturn]] void raise(const int status) {
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

	inline void check(const int status) {
		if (status != SHF_OK) {
			raise(status);
		}
	}

	struct STPGeometry {
		uint32_t MapSize[2], Neighbour[2], Total[2];

		STPGeometry(const SuperTerrainPlus::STPNearestNeighbourInformation& info) :
			MapSize { info.MapSize.x, info.MapSize.y }, Neighbour { info.ChunkNearestNeighbour.x, info.ChunkNearestNeighbour.y },
			Total { info.TotalMapSize.x, info.TotalMapSize.y } { }
	};

}

STPFilterBuffer::STPFilterBuffer(const STPExecutionType execution_type) : Memory(nullptr) {
	check(shf_buffer_create(&this->Memory, static_cast<unsigned char>(execution_type)));
}

STPFilterBuffer::STPFilterBuffer(STPFilterBuffer&& other) noexcept : Memory(other.Memory) {
	other.Memory = nullptr;
}

STPFilterBuffer& STPFilterBuffer::operator=(STPFilterBuffer&& other) noexcept {
	if (this != &other) {
		shf_buffer_destroy(this->Memory);
		this->Memory = other.Memory;
		other.Memory = nullptr;
	}
	return *this;
}

STPFilterBuffer::~STPFilterBuffer() {
	shf_buffer_destroy(this->Memory);
}

STPSingleHistogram STPFilterBuffer::readHistogram() const {
	const shf_bin* bin;
	const uint32_t* offset;
	check(shf_buffer_read(this->Memory, &bin, &offset));
	return STPSingleHistogram { reinterpret_cast<const STPSingleHistogram::STPBin*>(bin), offset };
}

STPFilterBuffer::STPHistogramSize STPFilterBuffer::size() const {
	size_t bin, offset;
	check(shf_buffer_size(this->Memory, &bin, &offset));
	return STPHistogramSize(bin, offset);
}

STPFilterBuffer::STPExecutionType STPFilterBuffer::type() const noexcept {
	return static_cast<STPExecutionType>(shf_buffer_type(this->Memory));
}

STPSingleHistogram STPFilterBuffer::readDeviceHistogram() const {
	const shf_bin* bin;
	const uint32_t* offset;
	check(shf_buffer_read_device(this->Memory, &bin, &offset));
	return STPSingleHistogram { reinterpret_cast<const STPSingleHistogram::STPBin*>(bin), offset };
}

std::uint64_t STPFilterBuffer::wait() {
	uint64_t repeated;
	check(shf_buffer_wait(this->Memory, &repeated));
	return repeated;
}

std::uint64_t STPFilterBuffer::chunkOffset(const unsigned int chunk) const {
	const uint64_t* base;
	uint32_t count;
	check(shf_buffer_chunk_base(this->Memory, &base, &count));
	if (!base || chunk > count) {
		throw STPException::STPNumericDomainError("chunk <= chunk count", "no such chunk in the last result");
	}
	return base[chunk];
}

STPSingleHistogramFilter::STPSingleHistogramFilter() : Filter(nullptr) {
	check(shf_filter_create(&this->Filter, -1));
}

STPSingleHistogramFilter::~STPSingleHistogramFilter() {
	shf_filter_destroy(this->Filter);
}

STPSingleHistogram STPSingleHistogramFilter::operator()(const STPSample_t* const samplemap,
	const STPNearestNeighbourInformation& nn_info, STPFilterBuffer& filter_buffer, const unsigned int radius) {
	const STPGeometry geo(nn_info);
	check(shf_run(this->Filter, samplemap, geo.MapSize, geo.Neighbour, geo.Total, filter_buffer.Memory, radius));
	return filter_buffer.readHistogram();
}

STPSingleHistogram STPSingleHistogramFilter::filterBatch(const STPSample_t* const* const samplemap, const unsigned int chunk_count,
	const STPNearestNeighbourInformation& nn_info, STPFilterBuffer& filter_buffer, const unsigned int radius) {
	const STPGeometry geo(nn_info);
	check(shf_run_batch(this->Filter, samplemap, chunk_count, geo.MapSize, geo.Neighbour, geo.Total, filter_buffer.Memory, radius));
	return filter_buffer.readHistogram();
}

void STPSingleHistogramFilter::filterMulti(const STPSample_t* const* const samplemap, STPFilterBuffer* const* const filter_buffer,
	const unsigned int call_count, const STPNearestNeighbourInformation& nn_info, const unsigned int radius) {
	const STPGeometry geo(nn_info);
	std::vector<shf_buffer*> memory(call_count);
	for (unsigned int i = 0u; i < call_count; i++) {
		memory[i] = filter_buffer[i] ? filter_buffer[i]->Memory : nullptr;
	}
	check(shf_run_multi(this->Filter, samplemap, memory.data(), call_count, geo.MapSize, geo.Neighbour, geo.Total, radius));
}

STPSingleHistogram STPSingleHistogramFilter::filterNeighbours(const STPSample_t* const* const neighbour_map,
	const unsigned int chunk_count, const STPNearestNeighbourInformation& nn_info, STPFilterBuffer& filter_buffer,
	const unsigned int radius) {
	const STPGeometry geo(nn_info);
	check(shf_run_neighbours(this->Filter, neighbour_map, chunk_count, geo.MapSize, geo.Neighbour, filter_buffer.Memory, radius));
	return filter_buffer.readHistogram();
}

void STPSingleHistogramFilter::filterNeighboursDevice(const STPSample_t* const* const neighbour_map, const unsigned int chunk_count,
	const STPNearestNeighbourInformation& nn_info, STPFilterBuffer& filter_buffer, const unsigned int radius, void* const stream) {
	const STPGeometry geo(nn_info);
	check(shf_run_neighbours_device(this->Filter, neighbour_map, chunk_count, geo.MapSize, geo.Neighbour, filter_buffer.Memory,
		radius, stream));
}

void STPSingleHistogramFilter::filterDevice(const STPSample_t* const samplemap_device, const std::uint64_t chunk_stride,
	const unsigned int chunk_count, const STPNearestNeighbourInformation& nn_info, STPFilterBuffer& filter_buffer,
	const unsigned int radius, void* const stream) {
	const STPGeometry geo(nn_info);
	check(shf_run_device(this->Filter, samplemap_device, chunk_stride, chunk_count, geo.MapSize, geo.Neighbour, geo.Total,
		filter_buffer.Memory, radius, stream));
}

void STPSingleHistogramFilter::filterDeviceAsync(const STPSample_t* const samplemap_device, const std::uint64_t chunk_stride,
	const unsigned int chunk_count, const STPNearestNeighbourInformation& nn_info, STPFilterBuffer& filter_buffer,
	const unsigned int radius, void* const stream) {
	const STPGeometry geo(nn_info);
	check(shf_run_device_async(this->Filter, samplemap_device, chunk_stride, chunk_count, geo.MapSize, geo.Neighbour, geo.Total,
		filter_buffer.Memory, radius, stream));
}

namespace SuperTerrainPlus::STPAlgorithm {
	//the C handle behind a filter object, for the sibling classes that live on the same device (STPBiomeFactoryDevice)
	shf_filter* STPFilterHandle(STPSingleHistogramFilter& filter) {
		return filter.handle();
	}
}
