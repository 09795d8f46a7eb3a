// C++ parity test of the drop-in class, written after the reference's own scenario
// (SuperTest+/SuperAlgorithm+/STPTestHistogram.cpp:116-172) without Catch2: same golden input, same expectations for
// pixels 0 / 8 / 15, same error cases, both execution types, re-run on the same buffer, size() / type() invariants --
// plus a full comparison of a larger random map against the CPU oracle (oracle/shf_oracle.c, test infrastructure).
#include <SuperAlgorithm+Host/STPSingleHistogramFilter.h>
#include <SuperTerrain+/Exception/STPFundamentalException.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <utility>
#include <vector>

using namespace SuperTerrainPlus;
using namespace SuperTerrainPlus::STPAlgorithm;
typedef STPSingleHistogramFilter::STPFilterBuffer FiltBuf;
typedef FiltBuf::STPExecutionType Exec;

extern "C" {
struct shf_oracle_bin {
	uint16_t item;
	float weight;
};
int shf_oracle_run(const uint16_t* map, uint32_t map_w, uint32_t map_h, uint32_t nn_x, uint32_t nn_y, uint32_t total_x,
	uint32_t radius, shf_oracle_bin** bins_out, uint32_t** offsets_out, uint64_t* n_bins_out);
void shf_oracle_free(void* p);
}

static int Failures = 0;
#define REQUIRE(COND) do { if (!(COND)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #COND); Failures++; } } while (0)
#define REQUIRE_THROWS_AS(EXPR, TYPE) do { bool caught_ = false; try { (void)(EXPR); } catch (const TYPE&) { caught_ = true; } \
	catch (...) { } if (!caught_) { std::printf("FAILED %s:%d: %s did not throw %s\n", __FILE__, __LINE__, #EXPR, #TYPE); Failures++; } } while (0)

//STPTestHistogram.cpp:44-57
static constexpr STPSample_t Texture[] = {
	2, 0, 2, 0, 1, 1, 0, 3, 3, 2, 1, 2,
	0, 1, 2, 1, 1, 1, 3, 2, 2, 0, 0, 1,
	0, 1, 0, 2, 1, 3, 1, 1, 1, 2, 1, 0,
	3, 3, 0, 1, 1, 2, 2, 2, 2, 0, 0, 3,
	3, 2, 3, 3, 0, 3, 2, 2, 1, 0, 0, 3,
	2, 0, 0, 1, 2, 0, 2, 2, 0, 2, 0, 3,
	1, 2, 3, 0, 3, 2, 1, 2, 2, 3, 0, 2,
	1, 1, 0, 3, 0, 2, 1, 0, 3, 2, 2, 1,
	2, 2, 0, 1, 0, 2, 0, 0, 0, 0, 1, 1,
	0, 1, 3, 3, 0, 3, 3, 3, 1, 0, 1, 1,
	3, 1, 3, 0, 2, 1, 1, 0, 2, 2, 2, 0,
	2, 0, 1, 1, 2, 3, 1, 2, 3, 2, 0, 1
};
static const STPNearestNeighbourInformation Data = { glm::uvec2(4u, 4u), glm::uvec2(3u, 3u), glm::uvec2(12u, 12u) };

static bool withinRel(const float value, const float target, const float eps) {
	return std::fabs(value - target) <= eps * std::fmax(std::fabs(value), std::fabs(target));
}

//STPTestHistogram.cpp:74-107
static void verifyHistogram(const STPSingleHistogram& result) {
	static constexpr unsigned int TrialOffset[] = { 0u, 8u, 15u };
	static constexpr std::pair<STPSample_t, float> Expected[] = {
		{ 0u, 0.24f }, { 3u, 0.24f }, { 2u, 0.28f }, { 1u, 0.24f },
		{ 0u, 0.36f }, { 3u, 0.24f }, { 1u, 0.16f }, { 2u, 0.24f },
		{ 0u, 0.32f }, { 3u, 0.2f }, { 1u, 0.12f }, { 2u, 0.36f }
	};
	for (int trial = 0; trial < 3; trial++) {
		const unsigned int offset = TrialOffset[trial];
		REQUIRE(result.HistogramStartOffset[offset + 1u] - result.HistogramStartOffset[offset] == 4u);
		int counter = 0;
		for (unsigned int i = result.HistogramStartOffset[offset]; i < result.HistogramStartOffset[offset + 1u]; i++, counter++) {
			const auto [item, weight] = Expected[counter + trial * 4];
			REQUIRE(result.Bin[i].Item == item);
			REQUIRE(withinRel(result.Bin[i].Weight, weight, std::numeric_limits<float>::epsilon() * 5.0f));
		}
	}
}

static void compareWithOracle(STPSingleHistogramFilter& filter, const unsigned int w, const unsigned int h, const unsigned int radius,
	const unsigned int biomes, const unsigned int seed) {
	std::mt19937 rng(seed);
	std::vector<STPSample_t> map(size_t(3u * w) * 3u * h);
	const unsigned int block = 1u + seed % 7u;
	for (unsigned int y = 0u; y < 3u * h; y++) {
		for (unsigned int x = 0u; x < 3u * w; x++) {
			std::seed_seq cell { seed, y / block, x / block };
			std::mt19937 local(cell);
			map[size_t(y) * 3u * w + x] = static_cast<STPSample_t>((seed & 1u) ? local() % biomes : rng() % biomes);
		}
	}
	const STPNearestNeighbourInformation info = { glm::uvec2(w, h), glm::uvec2(3u, 3u), glm::uvec2(3u * w, 3u * h) };
	FiltBuf buffer(Exec::Parallel);
	const STPSingleHistogram result = filter(map.data(), info, buffer, radius);
	shf_oracle_bin* bins = nullptr;
	uint32_t* offsets = nullptr;
	uint64_t n_bins = 0u;
	REQUIRE(shf_oracle_run(map.data(), w, h, 3u, 3u, 3u * w, radius, &bins, &offsets, &n_bins) == 0);
	const auto [bin_count, offset_count] = buffer.size();
	REQUIRE(bin_count == n_bins);
	REQUIRE(offset_count == size_t(w) * h + 1u);
	bool same = bin_count == n_bins && std::memcmp(offsets, result.HistogramStartOffset, offset_count * sizeof(uint32_t)) == 0;
	for (uint64_t i = 0u; same && i < n_bins; i++) {
		same = bins[i].item == result.Bin[i].Item && std::memcmp(&bins[i].weight, &result.Bin[i].Weight, sizeof(float)) == 0;
	}
	REQUIRE(same);
	//the same neighbourhood handed over as its nine separate chunk maps (what STPNearestNeighbourTextureBuffer is built
	//from, STPNearestNeighbourTextureBuffer.cpp:70-113) must give the very same histogram
	std::vector<std::vector<STPSample_t>> chunk(9u, std::vector<STPSample_t>(size_t(w) * h));
	const STPSample_t* chunk_ptr[9];
	for (unsigned int i = 0u; i < 9u; i++) {
		const unsigned int cx = i % 3u, cy = i / 3u;
		for (unsigned int y = 0u; y < h; y++) {
			std::memcpy(chunk[i].data() + size_t(y) * w, map.data() + (size_t(cy) * h + y) * 3u * w + size_t(cx) * w, w * sizeof(STPSample_t));
		}
		chunk_ptr[i] = chunk[i].data();
	}
	FiltBuf unmerged(Exec::Serial);
	const STPSingleHistogram again = filter.filterNeighbours(chunk_ptr, 1u, info, unmerged, radius);
	bool same_unmerged = unmerged.size().first == n_bins
		&& std::memcmp(offsets, again.HistogramStartOffset, offset_count * sizeof(uint32_t)) == 0;
	for (uint64_t i = 0u; same_unmerged && i < n_bins; i++) {
		same_unmerged = bins[i].item == again.Bin[i].Item && std::memcmp(&bins[i].weight, &again.Bin[i].Weight, sizeof(float)) == 0;
	}
	REQUIRE(same_unmerged);
	shf_oracle_free(bins);
	shf_oracle_free(offsets);
}

int main() {
	STPSingleHistogramFilter filter;
	//GIVEN a fresh buffer THEN it reads as empty (SHF.cpp:735-750)
	{
		FiltBuf fresh(Exec::Serial);
		const STPSingleHistogram empty = fresh.readHistogram();
		REQUIRE(empty.Bin == nullptr);
		REQUIRE(empty.HistogramStartOffset == nullptr);
		REQUIRE(fresh.size().first == 0u);
		REQUIRE(fresh.size().second == 0u);
	}
	//WHEN launching with wrong arguments THEN error is thrown (STPTestHistogram.cpp:121-130)
	{
		FiltBuf buffer(Exec::Parallel);
		REQUIRE_THROWS_AS(filter(Texture, Data, buffer, 0u), STPException::STPNumericDomainError);
		REQUIRE_THROWS_AS(filter(Texture, Data, buffer, 128u), STPException::STPNumericDomainError);
		REQUIRE_THROWS_AS(filter(Texture, Data, buffer, 3u), STPException::STPNumericDomainError);
		REQUIRE_THROWS_AS(FiltBuf(static_cast<Exec>(0x42u)), STPException::STPInvalidEnum);
	}
	//WHEN launching with correct arguments, for both execution types (STPTestHistogram.cpp:133-166)
	for (const Exec type : { Exec::Serial, Exec::Parallel }) {
		FiltBuf buffer(type);
		const STPSingleHistogram first = filter(Texture, Data, buffer, 2u);
		verifyHistogram(first);
		//the same buffer can be reused, and the stored output can be retrieved later
		filter(Texture, Data, buffer, 2u);
		verifyHistogram(buffer.readHistogram());
		REQUIRE(buffer.type() == type);
		const auto [bin_size, offset_size] = buffer.size();
		REQUIRE(bin_size == buffer.readHistogram().HistogramStartOffset[16u]);
		REQUIRE(bin_size == 64u);
		REQUIRE(offset_size == 17u);
		//moving a buffer keeps its content
		FiltBuf moved(std::move(buffer));
		verifyHistogram(moved.readHistogram());
	}
	//larger maps against the oracle, bit for bit
	compareWithOracle(filter, 48u, 40u, 16u, 24u, 3u);
	compareWithOracle(filter, 96u, 64u, 32u, 70u, 4u);
	compareWithOracle(filter, 33u, 57u, 8u, 5u, 9u);
	if (Failures == 0) {
		std::printf("all C++ histogram checks passed\n");
	}
	return Failures == 0 ? 0 : 1;
}
