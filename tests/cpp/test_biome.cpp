// The C++ device biome factory (include/SuperTerrain+/World/Diversity/STPBiomeFactoryDevice.h) feeding the drop-in filter
// class with no host copy of the map: the demo's chain (SuperDemo+/World/Layers/STPAllLayers.cpp:61-109) for two chunk
// neighbourhoods, produced in device memory, filtered in place, checked against the CPU restatements (test infrastructure:
// oracle/biome_oracle.c for the maps, oracle/shf_oracle.c for the histograms).
#include <SuperTerrain+/World/Diversity/STPBiomeFactoryDevice.h>
#include <SuperAlgorithm+Host/STPSingleHistogramFilter.h>

#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace SuperTerrainPlus;
using SuperTerrainPlus::STPAlgorithm::STPSingleHistogram;
using SuperTerrainPlus::STPAlgorithm::STPSingleHistogramFilter;
using SuperTerrainPlus::STPDiversity::STPBiomeFactoryDevice;

extern "C" {
struct oracle_bin { uint16_t item; float weight; };
int shf_oracle_run(const uint16_t* map, uint32_t w, uint32_t h, uint32_t nnx, uint32_t nny, uint32_t stride, uint32_t radius,
                   oracle_bin** bins, uint32_t** offsets, uint64_t* n_bins);
void shf_oracle_free(void* p);
int biome_oracle_run(const shf_biome_layer* layers, uint32_t n_layers, uint64_t global_seed, uint64_t voronoi_seed,
                     const uint16_t ids[7], int32_t offset_x, int32_t offset_z, uint32_t width, uint32_t height, uint16_t* out);
}

#define REQUIRE(cond)                                                                  \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            std::fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                              \
        }                                                                              \
    } while (0)

int main() {
    const unsigned W = 96, H = 64, r = 16, TW = 3 * W, TH = 3 * H;
    const uint64_t seed = 20261017ull;
    const shf_biome_layer chain[] = {
        {SHF_LAYER_CONTINENT, 0, 23457829ull}, {SHF_LAYER_SCALE_FUZZY, 0, 875944ull}, {SHF_LAYER_LAND, 1, 5748329ull},
        {SHF_LAYER_SCALE_NORMAL, 2, 8947358941ull}, {SHF_LAYER_LAND, 3, 361249673ull}, {SHF_LAYER_LAND, 4, 8769575ull},
        {SHF_LAYER_LAND, 5, 43562783426564ull}, {SHF_LAYER_ISLAND, 6, 74368ull}, {SHF_LAYER_SCALE_NORMAL, 7, 1ull},
        {SHF_LAYER_SCALE_NORMAL, 8, 2ull}, {SHF_LAYER_SCALE_NORMAL, 9, 3ull}, {SHF_LAYER_VORONOI, 10, 4ull},
        {SHF_LAYER_VORONOI, 11, 5ull}, {SHF_LAYER_VORONOI, 12, 6ull}};
    const shf_biome_ids ids = {0, 1, 3, 0, 0, 0, 0};
    const uint16_t id_array[7] = {0, 1, 3, 0, 0, 0, 0};
    const int origin[2][2] = {{-7 * (int)W, 3 * (int)H}, {1234 * (int)W, -77 * (int)H}};   // world coordinate of the centre chunks

    STPSingleHistogramFilter filter;
    STPSingleHistogramFilter::STPFilterBuffer buffer(STPSingleHistogramFilter::STPFilterBuffer::STPExecutionType::Parallel);
    STPBiomeFactoryDevice factory(filter, glm::uvec2(W + 2 * r, H + 2 * r), chain, 14u, seed, ids, seed);

    STPSample_t* merged = nullptr;
    REQUIRE(cudaMalloc(reinterpret_cast<void**>(&merged), sizeof(STPSample_t) * 2 * TW * TH) == cudaSuccess);
    REQUIRE(cudaMemset(merged, 0, sizeof(STPSample_t) * 2 * TW * TH) == cudaSuccess);
    // only the cells the filter reads of every merged map
    const int offsets[4] = {origin[0][0] - (int)r, origin[0][1] - (int)r, origin[1][0] - (int)r, origin[1][1] - (int)r};
    factory(merged + (H - r) * TW + (W - r), offsets, 2u, TW, (uint64_t)TW * TH, nullptr);
    const STPNearestNeighbourInformation info{glm::uvec2(W, H), glm::uvec2(3u, 3u), glm::uvec2(TW, TH)};
    filter.filterDevice(merged, (uint64_t)TW * TH, 2u, info, buffer, r, nullptr);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    const STPSingleHistogram dev = buffer.readDeviceHistogram();
    const auto size = buffer.size();
    std::vector<STPSingleHistogram::STPBin> bins(size.first);
    std::vector<unsigned> offs(size.second);
    REQUIRE(cudaMemcpy(bins.data(), dev.Bin, bins.size() * sizeof(bins[0]), cudaMemcpyDeviceToHost) == cudaSuccess);
    REQUIRE(cudaMemcpy(offs.data(), dev.HistogramStartOffset, offs.size() * sizeof(unsigned), cudaMemcpyDeviceToHost) == cudaSuccess);
    for (unsigned i = 0; i < 2u; i++) {
        std::vector<uint16_t> cpu_map((size_t)TW * TH);
        REQUIRE(biome_oracle_run(chain, 14u, seed, seed, id_array, origin[i][0] - (int)W, origin[i][1] - (int)H, TW, TH, cpu_map.data()) == 0);
        oracle_bin* want_bins;
        uint32_t* want_offs;
        uint64_t n;
        REQUIRE(shf_oracle_run(cpu_map.data(), W, H, 3, 3, TW, r, &want_bins, &want_offs, &n) == 0);
        const uint64_t base = buffer.chunkOffset(i);
        REQUIRE(buffer.chunkOffset(i + 1u) - base == n);
        for (unsigned p = 0; p <= W * H; p++) REQUIRE(offs[(size_t)i * (W * H + 1u) + p] == want_offs[p]);
        for (uint64_t k = 0; k < n; k++) {
            REQUIRE(bins[base + k].Item == want_bins[k].item);
            REQUIRE(std::memcmp(&bins[base + k].Weight, &want_bins[k].weight, 4) == 0);
        }
        shf_oracle_free(want_bins);
        shf_oracle_free(want_offs);
    }
    cudaFree(merged);
    std::printf("all C++ biome factory checks passed\n");
    return 0;
}
