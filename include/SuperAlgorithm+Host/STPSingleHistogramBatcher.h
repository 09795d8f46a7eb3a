#pragma once
#ifndef _STP_SINGLE_HISTOGRAM_BATCHER_H_
#define _STP_SINGLE_HISTOGRAM_BATCHER_H_

#include "STPSingleHistogramFilter.h"

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <exception>
#include <mutex>
#include <vector>

namespace SuperTerrainPlus::STPAlgorithm {

	/**
	 * @brief STPSingleHistogramBatcher serves concurrent STPSingleHistogramFilter::operator() calls with one pass over
	 * the device (additive, not in the reference).
	 * The world pipeline generates up to five chunks at once, every worker thread calling the filter with its own pooled
	 * STPFilterBuffer (SuperDemo+/World/Biomes/STPBiomefieldGenerator.cpp:79-104, .h:32,64-65;
	 * SuperTerrain+/Private/World/STPWorldPipeline.cpp:448-474). A single 512x512 chunk keeps a B200 busy for a few percent
	 * of the time its kernels take, so the calls that arrive together are run as ONE batch: the first caller becomes the
	 * leader, lingers a few hundred microseconds for the others (or until `max_batch` are waiting), runs
	 * STPSingleHistogramFilter::filterMulti for all of them and wakes them up; every caller gets its own histogram in its
	 * own buffer. The call keeps operator()'s contract: thread-safe for distinct buffers, synchronous, same exceptions
	 * (an exception raised by the batch is rethrown in every caller of that batch).
	 * A STPBiomefieldGenerator switches over by calling the batcher where it called the filter.
	*/
	class STP_ALGORITHM_HOST_API STPSingleHistogramBatcher {
	public:

		struct STPStatistics {
			std::uint64_t Call, Batch, LargestBatch;
		};

	private:

		struct STPRequest {
			const STPSample_t* Map;
			STPNearestNeighbourInformation Info;
			STPSingleHistogramFilter::STPFilterBuffer* Buffer;
			unsigned int Radius;
			bool Done;
			std::exception_ptr Error;
		};

		STPSingleHistogramFilter& Filter;
		const unsigned int MaxBatch;
		const std::chrono::microseconds Linger;

		std::mutex Lock;
		std::condition_variable Arrival, Completion;
		std::vector<STPRequest*> Pending;
		bool LeaderActive;
		STPStatistics Statistics;

	public:

		//filter: must outlive the batcher. max_batch: calls per pass (the number of pipeline workers is a good value).
		//linger: how long a leader waits for company before it runs alone.
		STPSingleHistogramBatcher(STPSingleHistogramFilter& filter, unsigned int max_batch = 5u,
			std::chrono::microseconds linger = std::chrono::microseconds(300));

		STPSingleHistogramBatcher(const STPSingleHistogramBatcher&) = delete;

		STPSingleHistogramBatcher& operator=(const STPSingleHistogramBatcher&) = delete;

		~STPSingleHistogramBatcher() = default;

		//same arguments, result and errors as STPSingleHistogramFilter::operator()
		STPSingleHistogram operator()(const STPSample_t*, const STPNearestNeighbourInformation&,
			STPSingleHistogramFilter::STPFilterBuffer&, unsigned int radius);

		STPStatistics statistics();

	};

}
#endif//_STP_SINGLE_HISTOGRAM_BATCHER_H_
