// B200-native STPSingleHistogramFilter: the public interface of
// SuperTerrain+/SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPSingleHistogramFilter.h:36-192, with the CPU thread
// pool and the CPU scratch memory behind it replaced by handles of the CUDA library (include/shf_b200.h).
//
// Kept: namespace, class and nested class names, STPExecutionType values, constructors / deleted copies / moves,
// readHistogram(), size(), type(), operator()(samplemap, nn_info, filter_buffer, radius) and its error behaviour
// (STPNumericDomainError for a bad radius, STPInvalidEnum for a bad execution type, STPCUDAError for CUDA failures).
// The sample map is a HOST pointer and the returned histogram lives in PAGE-LOCKED HOST memory owned by the filter
// buffer, exactly as with the reference, so STPBiomefieldGenerator (SuperDemo+/World/Biomes/STPBiomefieldGenerator.cpp:
// 102-123) compiles and runs unchanged. Both execution types run the same GPU path.
//
// Added (not in the reference): filterBatch(), filterDevice(), readDeviceHistogram(), chunkOffset() for batches of
// neighbourhoods and for consumers that read the histogram on the device without a host round trip.
#pragma once
#ifndef _STP_SINGLE_HISTOGRAM_FILTER_H_
#define _STP_SINGLE_HISTOGRAM_FILTER_H_

#include <SuperAlgorithm+Host/STPAlgorithmDefine.h>
#include <SuperTerrain+/World/STPWorldMapPixelFormat.hpp>
#include <SuperTerrain+/World/Chunk/STPNearestNeighbourInformation.hpp>
#include "STPSingleHistogram.hpp"

#include <cstddef>
#include <cstdint>
#include <utility>

struct shf_filter;
struct shf_buffer;

namespace SuperTerrainPlus::STPAlgorithm {

	class STP_ALGORITHM_HOST_API STPSingleHistogramFilter {
	public:

		class STP_ALGORITHM_HOST_API STPFilterBuffer {
		public:

			enum class STPExecutionType : unsigned char {
				Serial = 0x00u,
				Parallel = 0xFFu
			};

		private:

			friend class STPSingleHistogramFilter;

			//device scratch, device output and page-locked host output of one execution at a time
			shf_buffer* Memory;

		public:

			//number of bins and number of offsets of the histogram currently held
			typedef std::pair<size_t, size_t> STPHistogramSize;

			STPFilterBuffer(STPExecutionType);

			STPFilterBuffer(const STPFilterBuffer&) = delete;

			STPFilterBuffer(STPFilterBuffer&&) noexcept;

			STPFilterBuffer& operator=(const STPFilterBuffer&) = delete;

			STPFilterBuffer& operator=(STPFilterBuffer&&) noexcept;

			~STPFilterBuffer();

			//pointers into the buffer's page-locked host memory; {nullptr, nullptr} before the first execution
			STPSingleHistogram readHistogram() const;

			STPHistogramSize size() const;

			STPExecutionType type() const noexcept;

			/* ---- additive ---- */

			//device pointers of the last result (valid after filterDevice / filterBatch / operator())
			STPSingleHistogram readDeviceHistogram() const;

			//index of the first bin of chunk `chunk` of the last batch result; chunkOffset(chunk count) = total bins
			std::uint64_t chunkOffset(unsigned int chunk) const;

			//completes a pending filterDeviceAsync (every other query of the result does so as well); returns how many
			//calls on this buffer ran ahead with a plan that did not fit their input and were repeated so far
			std::uint64_t wait();

		};

	private:

		shf_filter* Filter;

	public:

		STPSingleHistogramFilter();

		STPSingleHistogramFilter(const STPSingleHistogramFilter&) = delete;

		STPSingleHistogramFilter(STPSingleHistogramFilter&&) = delete;

		STPSingleHistogramFilter& operator=(const STPSingleHistogramFilter&) = delete;

		STPSingleHistogramFilter& operator=(STPSingleHistogramFilter&&) = delete;

		~STPSingleHistogramFilter();

		//the C handle of include/shf_b200.h behind this object (additive)
		shf_filter* handle() noexcept { return this->Filter; }

		//samplemap: host pointer, row-major, row stride nn_info.TotalMapSize.x, not retained.
		//Synchronous; the result is also retrievable later with filter_buffer.readHistogram().
		STPSingleHistogram operator()(const STPSample_t*, const STPNearestNeighbourInformation&, STPFilterBuffer&, unsigned int);

		/* ---- additive ---- */

		//`chunk_count` independent neighbourhoods of equal geometry, host maps in, page-locked host histograms out:
		//bins concatenated, offsets as chunk_count blocks of MapSize.x * MapSize.y + 1 entries relative to each chunk.
		STPSingleHistogram filterBatch(const STPSample_t* const*, unsigned int chunk_count, const STPNearestNeighbourInformation&,
			STPFilterBuffer&, unsigned int);

		//`call_count` concurrent operator() calls served by one pass over the device: samplemap[i] / filter_buffer[i] are
		//what call i would hand to operator(); all share nn_info and radius, the buffers are distinct. Afterwards every
		//filter_buffer[i] reads exactly as after operator()(samplemap[i], nn_info, *filter_buffer[i], radius).
		//(STPSingleHistogramBatcher turns concurrent operator() calls into this.)
		void filterMulti(const STPSample_t* const* samplemap, STPFilterBuffer* const* filter_buffer, unsigned int call_count,
			const STPNearestNeighbourInformation&, unsigned int radius);

		//the filter fed by the UNMERGED neighbour chunk maps: `neighbour_map` holds chunk_count * nn.x * nn.y pointers
		//(host or device memory), neighbour i of a neighbourhood at local coordinate (i % nn.x, i / nn.x) exactly as
		//STPNearestNeighbourTextureBuffer takes them, every map MapSize.x * MapSize.y samples. Replaces the merged
		//page-locked copy of STPNearestNeighbourTextureBuffer::STPMergedBuffer in front of operator(): only the centre
		//chunk and its halo of `radius` samples are moved. nn_info.TotalMapSize is not used.
		STPSingleHistogram filterNeighbours(const STPSample_t* const* neighbour_map, unsigned int chunk_count,
			const STPNearestNeighbourInformation&, STPFilterBuffer&, unsigned int radius);

		//same, result left in device memory (readDeviceHistogram); copies and kernels are enqueued on `stream`
		void filterNeighboursDevice(const STPSample_t* const* neighbour_map, unsigned int chunk_count,
			const STPNearestNeighbourInformation&, STPFilterBuffer&, unsigned int radius, void* stream = nullptr);

		//merged maps already in device memory (chunk i at samplemap_device + i * chunk_stride samples); the result stays
		//in device memory (readDeviceHistogram). `stream` is a cudaStream_t; the call returns once the work is enqueued.
		void filterDevice(const STPSample_t* samplemap_device, std::uint64_t chunk_stride, unsigned int chunk_count,
			const STPNearestNeighbourInformation&, STPFilterBuffer&, unsigned int radius, void* stream = nullptr);

		//filterDevice without any host synchronisation when the shape equals that of the last completed call on the buffer:
		//the work is enqueued on `stream` with the previous call's plan, the first query of the result (wait(), size(),
		//readDeviceHistogram(), ...) checks it and repeats the call when the plan did not fit -- `samplemap_device` must
		//stay unchanged until then
		void filterDeviceAsync(const STPSample_t* samplemap_device, std::uint64_t chunk_stride, unsigned int chunk_count,
			const STPNearestNeighbourInformation&, STPFilterBuffer&, unsigned int radius, void* stream = nullptr);

	};

}
#endif//_STP_SINGLE_HISTOGRAM_FILTER_H_
