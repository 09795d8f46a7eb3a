// TEST INFRASTRUCTURE ONLY (oracle/): plain-C driver around the UNMODIFIED reference filter.
//
// Links the reference's own STPSingleHistogramFilter.cpp (compiled from /root/reference by oracle/Makefile) and
// exposes it through a C ABI so pytest / bench.py (cpu_baseline, --impl reference) can call it with ctypes.
// Nothing in the product path links or loads this.
#include <SuperAlgorithm+Host/STPSingleHistogramFilter.h>
#include <SuperTerrain+/Exception/STPNumericDomainError.h>
#include <SuperTerrain+/Exception/STPInvalidEnum.h>

#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <new>

using SuperTerrainPlus::STPNearestNeighbourInformation;
using SuperTerrainPlus::STPAlgorithm::STPSingleHistogram;
using SuperTerrainPlus::STPAlgorithm::STPSingleHistogramFilter;
typedef STPSingleHistogramFilter::STPFilterBuffer FilterBuffer;

namespace {
struct RefSession {
    STPSingleHistogramFilter Filter;
    std::unique_ptr<FilterBuffer> Buffer;
};
}

extern "C" {

// status: 0 ok, 1 STPNumericDomainError, 2 STPInvalidEnum, 3 other reference exception, 4 other
void* ref_shf_create(unsigned char exec_type, int* status) {
    try {
        auto s = std::make_unique<RefSession>();
        s->Buffer = std::make_unique<FilterBuffer>(static_cast<FilterBuffer::STPExecutionType>(exec_type));
        *status = 0;
        return s.release();
    } catch (const SuperTerrainPlus::STPException::STPInvalidEnum&) {
        *status = 2;
    } catch (const std::exception&) {
        *status = 4;
    }
    return nullptr;
}

void ref_shf_destroy(void* session) { delete static_cast<RefSession*>(session); }

int ref_shf_run(void* session, const uint16_t* map, const uint32_t map_size[2], const uint32_t nn[2],
                const uint32_t total[2], uint32_t radius, const void** bins, const uint32_t** offsets,
                uint64_t* n_bins, uint64_t* n_offsets) {
    RefSession* s = static_cast<RefSession*>(session);
    const STPNearestNeighbourInformation info{glm::uvec2(map_size[0], map_size[1]), glm::uvec2(nn[0], nn[1]),
                                              glm::uvec2(total[0], total[1])};
    try {
        const STPSingleHistogram h = s->Filter(map, info, *s->Buffer, radius);
        const auto sz = s->Buffer->size();
        *bins = h.Bin;
        *offsets = h.HistogramStartOffset;
        *n_bins = sz.first;
        *n_offsets = sz.second;
        return 0;
    } catch (const SuperTerrainPlus::STPException::STPNumericDomainError&) {
        return 1;
    } catch (const SuperTerrainPlus::STPException::STPInvalidEnum&) {
        return 2;
    } catch (const SuperTerrainPlus::STPException::STPFundamentalException::STPBasic&) {
        return 3;
    } catch (const std::exception&) {
        return 4;
    }
}

unsigned char ref_shf_type(void* session) {
    return static_cast<unsigned char>(static_cast<RefSession*>(session)->Buffer->type());
}

int ref_shf_size(void* session, uint64_t* n_bins, uint64_t* n_offsets) {
    const auto sz = static_cast<RefSession*>(session)->Buffer->size();
    *n_bins = sz.first;
    *n_offsets = sz.second;
    const STPSingleHistogram h = static_cast<RefSession*>(session)->Buffer->readHistogram();
    return (h.Bin == nullptr ? 1 : 0) | (h.HistogramStartOffset == nullptr ? 2 : 0);
}

unsigned ref_shf_bin_stride(void) { return static_cast<unsigned>(sizeof(STPSingleHistogram::STPBin)); }
}
