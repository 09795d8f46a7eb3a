// The device-free part of the drop-in C++ surface: what the reference's class promises a caller
// before any filtering happens -- SuperAlgorithm+Host/STPSingleHistogramFilter.h:36-192, STPSingleHistogram.hpp:15-50,
// STPTestHistogram.cpp:157-161 (type / size echo). Runs on hosts without a GPU; there the filter constructor must throw
// STPCUDAError (no CPU path), on a GPU host it must succeed. (test infrastructure)
#include <SuperAlgorithm+Host/STPSingleHistogramFilter.h>
#include <SuperTerrain+/Exception/STPFundamentalException.h>

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <utility>

using namespace SuperTerrainPlus;
using namespace SuperTerrainPlus::STPAlgorithm;
using FiltBuf = STPSingleHistogramFilter::STPFilterBuffer;
using Exec = FiltBuf::STPExecutionType;

// ---- compile-time surface: the reference deletes the copies of both classes and the moves of the filter ----
static_assert(!std::is_copy_constructible_v<STPSingleHistogramFilter> && !std::is_copy_assignable_v<STPSingleHistogramFilter>);
static_assert(!std::is_move_constructible_v<STPSingleHistogramFilter> && !std::is_move_assignable_v<STPSingleHistogramFilter>);
static_assert(std::is_default_constructible_v<STPSingleHistogramFilter>);
static_assert(!std::is_copy_constructible_v<FiltBuf> && !std::is_copy_assignable_v<FiltBuf>);
static_assert(std::is_nothrow_move_constructible_v<FiltBuf> && std::is_nothrow_move_assignable_v<FiltBuf>);
static_assert(!std::is_default_constructible_v<FiltBuf>, "a buffer is made for an execution type");
static_assert(static_cast<unsigned char>(Exec::Serial) == 0x00u && static_cast<unsigned char>(Exec::Parallel) == 0xFFu);
static_assert(std::is_same_v<FiltBuf::STPHistogramSize, std::pair<size_t, size_t>>);
// STPBin: {uint16 Item; float Weight}, 8 bytes, weight at byte 4 (SH.hpp:23-37; SHF.cpp:326 asserts the size)
static_assert(sizeof(STPSingleHistogram::STPBin) == 8u && alignof(STPSingleHistogram::STPBin) == 4u);
static_assert(offsetof(STPSingleHistogram::STPBin, Item) == 0u && offsetof(STPSingleHistogram::STPBin, Weight) == 4u);
static_assert(std::is_same_v<STPSample_t, std::uint16_t> || sizeof(STPSample_t) == 2u);
static_assert(std::is_same_v<decltype(std::declval<STPSingleHistogramFilter&>()(
	std::declval<const STPSample_t*>(), std::declval<const STPNearestNeighbourInformation&>(), std::declval<FiltBuf&>(), 0u)),
	STPSingleHistogram>, "operator()(samplemap, nn_info, filter_buffer, radius) -> STPSingleHistogram");
// the exception types the path throws derive from STPBasic : std::exception (STPFundamentalException.h:38-82)
static_assert(std::is_base_of_v<std::exception, STPException::STPFundamentalException::STPBasic>);
static_assert(std::is_base_of_v<STPException::STPFundamentalException::STPBasic, STPException::STPNumericDomainError>);
static_assert(std::is_base_of_v<STPException::STPFundamentalException::STPBasic, STPException::STPInvalidEnum>);
static_assert(std::is_base_of_v<STPException::STPFundamentalException::STPBasic, STPException::STPCUDAError>);

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

int main() {
	for (const Exec type : { Exec::Serial, Exec::Parallel }) {
		FiltBuf buffer(type);
		CHECK(buffer.type() == type);
		CHECK(buffer.size() == FiltBuf::STPHistogramSize(0u, 0u));
		const STPSingleHistogram fresh = buffer.readHistogram();
		CHECK(fresh.Bin == nullptr && fresh.HistogramStartOffset == nullptr);
		const STPSingleHistogram fresh_device = buffer.readDeviceHistogram();
		CHECK(fresh_device.Bin == nullptr && fresh_device.HistogramStartOffset == nullptr);
		CHECK(buffer.wait() == 0u);   // nothing pending
		// moves carry the handle along; the moved-from object stays destructible and assignable
		FiltBuf moved(std::move(buffer));
		CHECK(moved.type() == type && moved.size() == FiltBuf::STPHistogramSize(0u, 0u));
		FiltBuf other(type == Exec::Serial ? Exec::Parallel : Exec::Serial);
		other = std::move(moved);
		CHECK(other.type() == type);
		moved = FiltBuf(Exec::Parallel);
		CHECK(moved.type() == Exec::Parallel);
		FiltBuf& self = other;
		other = std::move(self);   // self-move leaves the buffer intact
		CHECK(other.type() == type);
	}
	// SHF.cpp:721: a value outside the enumeration -> STPInvalidEnum
	bool thrown = false;
	try {
		FiltBuf bad(static_cast<Exec>(0x42u));
	} catch (const STPException::STPInvalidEnum& e) {
		thrown = std::strstr(e.what(), "STPExecutionType") != nullptr;
	}
	CHECK(thrown);
	// no CPU path: without a device the filter cannot be constructed, and says why
	bool constructed = false, cuda_error = false;
	try {
		STPSingleHistogramFilter filter;
		constructed = filter.handle() != nullptr;
	} catch (const STPException::STPCUDAError& e) {
		cuda_error = std::strlen(e.what()) > 0u;
		std::printf("filter constructor: %s\n", e.what());
	}
	CHECK(constructed != cuda_error);
	std::printf("device %s\n", constructed ? "present" : "absent");
	if (failures == 0) std::printf("all C++ surface checks passed\n");
	return failures == 0 ? 0 : 1;
}
