// Concurrent callers through STPSingleHistogramBatcher: the world pipeline's pattern -- several
// worker threads, one pooled STPFilterBuffer each, one shared filter (SuperDemo+/World/Biomes/STPBiomefieldGenerator.cpp
// :79-104) -- must give every caller exactly the histogram a direct operator() call gives, which in turn is compared with
// the CPU oracle (oracle/shf_oracle.c, test infrastructure). Also prints the wall time of the same calls made directly.
#include <SuperAlgorithm+Host/STPSingleHistogramBatcher.h>
#include <SuperTerrain+/Exception/STPFundamentalException.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

using namespace SuperTerrainPlus;
using namespace SuperTerrainPlus::STPAlgorithm;
typedef STPSingleHistogramFilter::STPFilterBuffer FiltBuf;

extern "C" {
struct shf_oracle_bin {
	uint16_t item;
	float weight;
};
int shf_oracle_run(const uint16_t* map, uint32_t map_w, uint32_t map_h, uint32_t nn_x, uint32_t nn_y, uint32_t total_x,
	uint32_t radius, shf_oracle_bin** bins_out, uint32_t** offsets_out, uint64_t* n_bins_out);
void shf_oracle_free(void* p);
}

static std::atomic<int> Failures { 0 };
#define REQUIRE(COND) do { if (!(COND)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #COND); Failures++; } } while (0)

static std::vector<STPSample_t> makeMap(const unsigned int w, const unsigned int h, const unsigned int biomes, const unsigned int seed) {
	std::mt19937 rng(seed);
	std::vector<STPSample_t> map(size_t(3u * w) * 3u * h);
	const unsigned int block = 1u + seed % 9u;
	std::vector<STPSample_t> coarse(size_t(3u * w / block + 1u) * (3u * h / block + 1u));
	for (STPSample_t& v : coarse) {
		v = static_cast<STPSample_t>(rng() % biomes);
	}
	for (unsigned int y = 0u; y < 3u * h; y++) {
		for (unsigned int x = 0u; x < 3u * w; x++) {
			map[size_t(y) * 3u * w + x] = coarse[size_t(y / block) * (3u * w / block + 1u) + x / block];
		}
	}
	return map;
}

static bool sameAsOracle(const std::vector<STPSample_t>& map, const unsigned int w, const unsigned int h, const unsigned int radius,
	const STPSingleHistogram& result, const FiltBuf& buffer) {
	shf_oracle_bin* bins = nullptr;
	uint32_t* offsets = nullptr;
	uint64_t n_bins = 0u;
	if (shf_oracle_run(map.data(), w, h, 3u, 3u, 3u * w, radius, &bins, &offsets, &n_bins) != 0) {
		return false;
	}
	const auto [bin_count, offset_count] = buffer.size();
	bool same = bin_count == n_bins && offset_count == size_t(w) * h + 1u
		&& std::memcmp(offsets, result.HistogramStartOffset, offset_count * sizeof(uint32_t)) == 0;
	for (uint64_t i = 0u; same && i < n_bins; i++) {
		same = bins[i].item == result.Bin[i].Item && std::memcmp(&bins[i].weight, &result.Bin[i].Weight, sizeof(float)) == 0;
	}
	shf_oracle_free(bins);
	shf_oracle_free(offsets);
	return same;
}

int main() {
	constexpr unsigned int Worker = 5u, Round = 4u;
	STPSingleHistogramFilter filter;
	STPSingleHistogramBatcher batcher(filter, Worker, std::chrono::microseconds(20000));

	//two geometries in flight at once: workers 0-2 filter 96x80 chunks with radius 16, workers 3-4 64x64 chunks with radius 32
	struct Job {
		unsigned int W, H, Radius, Biomes;
	};
	const Job job[Worker] = { { 96u, 80u, 16u, 12u }, { 96u, 80u, 16u, 12u }, { 96u, 80u, 16u, 40u }, { 64u, 64u, 32u, 7u }, { 64u, 64u, 32u, 300u } };
	std::vector<std::thread> pool;
	for (unsigned int t = 0u; t < Worker; t++) {
		pool.emplace_back([&, t]() {
			FiltBuf buffer(FiltBuf::STPExecutionType::Parallel);
			const Job& j = job[t];
			const STPNearestNeighbourInformation info = { glm::uvec2(j.W, j.H), glm::uvec2(3u, 3u), glm::uvec2(3u * j.W, 3u * j.H) };
			for (unsigned int round = 0u; round < Round; round++) {
				const std::vector<STPSample_t> map = makeMap(j.W, j.H, j.Biomes, 1000u * t + round);
				const STPSingleHistogram result = batcher(map.data(), info, buffer, j.Radius);
				REQUIRE(sameAsOracle(map, j.W, j.H, j.Radius, result, buffer));
				REQUIRE(sameAsOracle(map, j.W, j.H, j.Radius, buffer.readHistogram(), buffer));
			}
			//errors reach the caller that made them (SHF.cpp:874)
			bool caught = false;
			try {
				const std::vector<STPSample_t> map = makeMap(j.W, j.H, j.Biomes, 7u);
				batcher(map.data(), info, buffer, 3u);
			} catch (const STPException::STPNumericDomainError&) {
				caught = true;
			} catch (...) { }
			REQUIRE(caught);
		});
	}
	for (std::thread& th : pool) {
		th.join();
	}
	const STPSingleHistogramBatcher::STPStatistics stat = batcher.statistics();
	std::printf("batcher: %llu calls in %llu passes, largest pass %llu calls\n", static_cast<unsigned long long>(stat.Call),
		static_cast<unsigned long long>(stat.Batch), static_cast<unsigned long long>(stat.LargestBatch));
	REQUIRE(stat.Call == Worker * (Round + 1u));
	REQUIRE(stat.LargestBatch >= 2u);
	REQUIRE(stat.Batch < stat.Call);

	//the pipeline's shape: 5 workers x 512x512 chunks, radius 32, 8 biomes (BASELINE.json config 1), direct vs batched
	{
		constexpr unsigned int W = 512u, R = 32u, Calls = 8u;
		const STPNearestNeighbourInformation info = { glm::uvec2(W, W), glm::uvec2(3u, 3u), glm::uvec2(3u * W, 3u * W) };
		std::vector<std::vector<STPSample_t>> map;
		for (unsigned int t = 0u; t < Worker; t++) {
			map.push_back(makeMap(W, W, 8u, 77u + t));
		}
		STPSingleHistogramBatcher quick(filter, Worker, std::chrono::microseconds(300));
		//pooled buffers, grown by a first call each (the reference's buffers are an adaptive pool as well, SHF.h:33-34)
		std::vector<FiltBuf> buffer;
		for (unsigned int t = 0u; t < Worker; t++) {
			buffer.emplace_back(FiltBuf::STPExecutionType::Parallel);
			filter(map[t].data(), info, buffer[t], R);
		}
		for (int mode = 0; mode < 2; mode++) {
			double best = 1e30;
			for (int rep = 0; rep < 3; rep++) {
				std::vector<std::thread> workers;
				const auto t0 = std::chrono::steady_clock::now();
				for (unsigned int t = 0u; t < Worker; t++) {
					workers.emplace_back([&, t]() {
						for (unsigned int c = 0u; c < Calls; c++) {
							if (mode == 0) {
								filter(map[t].data(), info, buffer[t], R);
							} else {
								quick(map[t].data(), info, buffer[t], R);
							}
						}
					});
				}
				for (std::thread& th : workers) {
					th.join();
				}
				best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
			}
			std::printf("%s: %u workers x %u calls of 512x512 r=32, 8 biomes, host maps in, page-locked histograms out: %.2f ms (%.3f ms per chunk)\n",
				mode == 0 ? "direct " : "batched", Worker, Calls, best, best / (Worker * Calls));
		}
	}
	if (Failures == 0) {
		std::printf("all batcher checks passed\n");
	}
	return Failures == 0 ? 0 : 1;
}
