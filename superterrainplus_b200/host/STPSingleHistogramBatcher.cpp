// Leader / follower coalescing of concurrent filter calls (include/SuperAlgorithm+Host/STPSingleHistogramBatcher.h).
#include <SuperAlgorithm+Host/STPSingleHistogramBatcher.h>

#include <algorithm>

using namespace SuperTerrainPlus::STPAlgorithm;

namespace {

	inline bool sameGeometry(const SuperTerrainPlus::STPNearestNeighbourInformation& a, const SuperTerrainPlus::STPNearestNeighbourInformation& b) {
		return a.MapSize.x == b.MapSize.x && a.MapSize.y == b.MapSize.y
			&& a.ChunkNearestNeighbour.x == b.ChunkNearestNeighbour.x && a.ChunkNearestNeighbour.y == b.ChunkNearestNeighbour.y
			&& a.TotalMapSize.x == b.TotalMapSize.x && a.TotalMapSize.y == b.TotalMapSize.y;
	}

}

STPSingleHistogramBatcher::STPSingleHistogramBatcher(STPSingleHistogramFilter& filter, const unsigned int max_batch,
	const std::chrono::microseconds linger) : Filter(filter), MaxBatch(std::max(1u, max_batch)), Linger(linger),
	LeaderActive(false), Statistics { } { }

STPSingleHistogram STPSingleHistogramBatcher::operator()(const STPSample_t* const samplemap,
	const STPNearestNeighbourInformation& nn_info, STPSingleHistogramFilter::STPFilterBuffer& filter_buffer, const unsigned int radius) {
	STPRequest self { samplemap, nn_info, &filter_buffer, radius, false, nullptr };
	std::unique_lock<std::mutex> guard(this->Lock);
	this->Pending.push_back(&self);
	this->Statistics.Call++;
	this->Arrival.notify_all();

	while (!self.Done) {
		if (this->LeaderActive) {
			//somebody else is collecting or running a batch; it may or may not contain this request
			this->Completion.wait(guard);
			continue;
		}
		//become the leader: wait a little for company, then take every pending request of this geometry
		this->LeaderActive = true;
		this->Arrival.wait_for(guard, this->Linger, [this]() { return this->Pending.size() >= this->MaxBatch; });
		std::vector<STPRequest*> batch;
		for (auto it = this->Pending.begin(); it != this->Pending.end() && batch.size() < this->MaxBatch;) {
			STPRequest* const candidate = *it;
			if (candidate == &self || (sameGeometry(candidate->Info, self.Info) && candidate->Radius == self.Radius)) {
				batch.push_back(candidate);
				it = this->Pending.erase(it);
			} else {
				++it;
			}
		}
		this->Statistics.Batch++;
		this->Statistics.LargestBatch = std::max<std::uint64_t>(this->Statistics.LargestBatch, batch.size());
		guard.unlock();

		std::exception_ptr error = nullptr;
		try {
			std::vector<const STPSample_t*> map(batch.size());
			std::vector<STPSingleHistogramFilter::STPFilterBuffer*> buffer(batch.size());
			for (size_t i = 0u; i < batch.size(); i++) {
				map[i] = batch[i]->Map;
				buffer[i] = batch[i]->Buffer;
			}
			this->Filter.filterMulti(map.data(), buffer.data(), static_cast<unsigned int>(batch.size()), self.Info, self.Radius);
		} catch (...) {
			error = std::current_exception();
		}

		guard.lock();
		for (STPRequest* const request : batch) {
			request->Error = error;
			request->Done = true;
		}
		this->LeaderActive = false;
		this->Completion.notify_all();
	}
	guard.unlock();
	if (self.Error) {
		std::rethrow_exception(self.Error);
	}
	return filter_buffer.readHistogram();
}

STPSingleHistogramBatcher::STPStatistics STPSingleHistogramBatcher::statistics() {
	const std::lock_guard<std::mutex> guard(this->Lock);
	return this->Statistics;
}
