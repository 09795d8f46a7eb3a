// shf_capi.cu -- the C ABI declared in include/shf_b200.h: handles, growable device / page-locked buffers and the launch
// sequence of the kernels in shf_kernels.cuh. No torch, no CPU compute path: every result comes from the kernels.
#include "../../include/shf_b200.h"
#include "shf_kernels.cuh"
#include "shf_events.cuh"
#include "shf_generic.cuh"
#include "shf_heightfield.cuh"
#include "shf_biome.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local std::string tls_error;
thread_local uint64_t tls_launches = 0, tls_h2d = 0, tls_d2h = 0;
volatile int g_profiling = 0;  // shf_set_profiling: record CUDA events between the kernel phases
constexpr int kEvRing = 64;    // profiled calls whose phase events are kept per buffer (read back without a sync per call)
constexpr int kPhases = 6;     // dictionary | remap + vscan | event lists | row scan | host gap | emit

// The reference's operator() never touches the calling thread's CUDA device; every entry point that binds the filter's
// device restores the caller's on the way out.
struct DeviceGuard {
    int prev = -1, want = -1;
    explicit DeviceGuard(int device) : want(device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != want) cudaSetDevice(want);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int fail(int status, const char* expr, const std::string& what) {
    tls_error = std::string(expr) + ": " + what;
    return status;
}

#define SHF_CUDA(call)                                                                              \
    do {                                                                                            \
        const cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return fail(SHF_ERR_CUDA, #call, std::string(cudaGetErrorString(e_))); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        size_t want = std::max(bytes, std::min(cap * 2, cap + (size_t(1) << 30)));
        want = (want + 511) & ~size_t(511);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && want != bytes) {
            (void)cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        cap = (e == cudaSuccess) ? want : 0;
        return e;
    }
    // for buffers whose users leave them zeroed: zero-filled whenever (re)allocated
    cudaError_t ensure_zeroed(size_t bytes, cudaStream_t s) {
        if (bytes <= cap) return cudaSuccess;
        const cudaError_t e = ensure(bytes);
        return e == cudaSuccess ? cudaMemsetAsync(p, 0, cap, s) : e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// page-locked host memory: the reference keeps its output in cudaMallocHost memory as well
// (STPSmartDeviceMemory::makeHost, SHF.cpp:80,105) so that callers can cudaMemcpyAsync from it
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        size_t want = std::max(bytes, std::min(cap * 2, cap + (size_t(1) << 30)));
        // (portable: page-locked for every CUDA context of the process, so that a caller may copy from it on any device)
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e != cudaSuccess && want != bytes) {
            (void)cudaGetLastError();
            want = bytes;
            e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        }
        cap = (e == cudaSuccess) ? want : 0;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

// measurement toggles, read once when the filter is created (never needed for correctness)
struct shf_toggles {
    bool no_speculation = false, no_cseg = false, no_vseg = false;
    uint32_t persist = 0;   // 1: persistent emit CTAs with the phase shift, 2: without
    bool no_split = false;  // large calls: one emit launch instead of bulk + low-priority tail
    bool no_small_ids = false;  // always build the dictionary and the remapped copy
    uint32_t debug_tail = 0, debug_tailseg = 0;
    uint32_t debug_ty = 0, debug_cseg = 0, debug_vseg = 0, debug_persist = 0, debug_np = 0, debug_extra = 0;  // 0 = not set
};

struct shf_filter {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;    // per block
    size_t smem_per_sm = 0;
    shf_toggles dbg;
};

struct shf_heightfield {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    uint32_t n_table = 0, grad_size = 0;
    DevBuf table, perm, grad;
};

// the biome-map producer: the layer chain, its seeds and per-layer device scratch
struct shf_biome_factory {
    int device = 0;
    int sm_count = 0;
    uint32_t width = 0, height = 0;
    std::vector<shf_biome_layer> layers;
    std::vector<uint64_t> seeds;          // STPLayer::seedLayer(global seed, salt) of every layer
    uint64_t voronoi_seed = 0;
    shf::BiomeIds ids{};
    std::vector<DevBuf> grid;             // per layer: the cells its descendants read, all maps of a batch
    DevBuf rects, jitter;
    PinBuf h_rects;
    cudaEvent_t rects_free = nullptr;     // the upload of the previous call's rectangles has left the pinned block
};

struct shf_buffer {
    unsigned char exec_type = SHF_EXEC_PARALLEL;
    int device = -1;
    cudaStream_t stream = nullptr;  // owned, used by the host-pointer entry points
    cudaStream_t copy_stream = nullptr;   // owned: carries finished ranges of a result to the host while the next is emitted
    cudaStream_t emit_stream = nullptr;   // owned, highest priority: the bulk of a large call's tiles (see launch_emit_split)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t aux_stream = nullptr;    // owned: work of a call that is off its critical path (the dictionary of a call
    cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;   // that reads the caller's sample values as compact ids)
    cudaEvent_t ev_totals = nullptr;
    bool aux_pending = false;             // the read-backs of the pending call are on aux_stream (ev_aux_join marks their end)
    cudaEvent_t range_ev[8] = {};
    DevBuf din, cmap, vstart, bitmap, prefix, nbiomes, dict, base, colmask, rowtotal, rowbase, chunktotal, chunkbase,
        bins, hso, gstate, evpool, rowinfo, cvt, vexit, sync;
    PinBuf h_small, h_bins, h_hso;
    std::vector<uint64_t> chunk_base;  // n_chunks + 1
    uint32_t n_chunks = 0, last_w = 0, last_h = 0;
    size_t n_bins = 0, n_offsets = 0;
    DevBuf hf_offsets;  // per-chunk noise offsets of shf_heightfield_run
    bool has_result = false, on_host = false;
    bool dev_valid = false;  // the device arrays hold this buffer's result (not so for the followers of shf_run_multi)
    uint32_t plan_k = 0, plan_ty = 0, plan_biomes = 0, plan_smem = 0;
    // the geometry and the plan of the last event-list call: a repeated call of the same shape runs with it straight
    // away and checks afterwards (run_speculative)
    shf::Geo last_g{};
    bool last_valid = false;
    // a call that ran ahead of its checks (enqueue_speculative) and what settle() needs to complete or repeat it
    struct Deferred {
        bool active = false;
        shf_filter filt{};
        shf::Geo g{};
        const uint16_t* in_dev = nullptr;
        bool vec8 = false;
        cudaStream_t stream = nullptr;
    } deferred;
    cudaEvent_t done_ev = nullptr;      // recorded behind the last operation of every call, on the call's stream
    cudaStream_t done_stream = nullptr;
    bool done_recorded = false;
    uint64_t spec_misses = 0;           // calls that ran ahead with a plan that did not fit and were repeated
    const uint16_t* ids = nullptr;      // where the current call's kernels read compact ids: cmap, or the caller's maps
    cudaEvent_t ev[kEvRing][kPhases + 1] = {};
    uint64_t ev_calls = 0;   // profiled calls so far; call c uses ring slot c % kEvRing
    bool ev_valid = false;
    cudaError_t mark(int i, cudaStream_t s) {
        if (!g_profiling) return cudaSuccess;
        if (i == 0) ev_calls++;
        cudaEvent_t& e = ev[(ev_calls - 1u) % kEvRing][i];
        if (!e) {
            const cudaError_t err = cudaEventCreate(&e);
            if (err != cudaSuccess) return err;
        }
        return cudaEventRecord(e, s);
    }
    void release_all() {
        for (auto& set : ev)
            for (cudaEvent_t& e : set) {
                if (e) cudaEventDestroy(e);
                e = nullptr;
            }
        ev_valid = false;
        ev_calls = 0;
        last_valid = false;
        deferred.active = false;
        if (done_ev) cudaEventDestroy(done_ev);
        done_ev = nullptr;
        done_recorded = false;
        DevBuf* d[] = {&din, &cmap, &vstart, &bitmap, &prefix, &nbiomes, &dict, &base,
                       &colmask, &rowtotal, &rowbase, &chunktotal, &chunkbase, &bins, &hso, &gstate, &hf_offsets,
                       &evpool, &rowinfo, &cvt, &vexit, &sync};
        for (DevBuf* b : d) b->release();
        h_small.release();
        h_bins.release();
        h_hso.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        if (copy_stream) cudaStreamDestroy(copy_stream);
        copy_stream = nullptr;
        if (emit_stream) cudaStreamDestroy(emit_stream);
        emit_stream = nullptr;
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        ev_fork = ev_join = nullptr;
        if (aux_stream) cudaStreamDestroy(aux_stream);
        aux_stream = nullptr;
        if (ev_aux_fork) cudaEventDestroy(ev_aux_fork);
        if (ev_aux_join) cudaEventDestroy(ev_aux_join);
        if (ev_totals) cudaEventDestroy(ev_totals);
        ev_aux_fork = ev_aux_join = ev_totals = nullptr;
        aux_pending = false;
        for (cudaEvent_t& e : range_ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
    }
};

namespace {

using shf::Geo;

// CTAs per chunk of the row-streaming kernels (dictionary, remap): about 8 CTAs per SM over the whole batch; every CTA
// carries a fixed cost (zeroing / loading 8-16 KB of tables), so large batches use few CTAs per chunk
uint32_t shf_rows_grid(uint32_t PH, uint32_t n_chunks, int sm_count) {
    const uint32_t want = (8u * (uint32_t)sm_count + n_chunks - 1u) / n_chunks;
    return std::max(1u, std::min(std::min(want, PH), 256u));
}

// state table + transposed masks + the staging tile of one vscan CTA (one warp)
size_t vscan_smem(int K) {
    return (size_t)shf::kVscanThreads * (4 * 32 * K) + 4 * 32 * K + (size_t)32 * shf::kVscanStagePitch * 4;
}

size_t emit_smem(uint32_t ty, uint32_t R, int K, int FW) {
    return (size_t)ty * R * 32 * K * FW / 8 + (size_t)ty * shf::kMarchNB * shf::emit_sbuf_stride(K, FW) +
           (size_t)ty * shf::emit_act_cap(K) * 8 + 128 +
           4 * 1024;  // barriers + the producers' sample scratch (<= 1 KB per producer warp)
}

template <int K>
int launch_events(shf_buffer* b, const Geo& g, cudaStream_t s) {
    // sync: [0..1] the event counter (u64), [2] chunks finished, [3 + n] CTAs of chunk n finished, [3 + n_chunks + n]
    // presence CTAs of chunk n finished; all left at zero by their kernels
    uint32_t* sync = b->sync.as<uint32_t>();
    const dim3 egrid((g.H + shf::kEventWarps - 1) / shf::kEventWarps, g.n_chunks);
    const bool fold = (uint64_t)egrid.x * egrid.y <= 1024ull;   // small calls: the scans ride on the last CTAs
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(sync);
#define SHF_EVENTS_ARGS                                                                                                \
    g, b->colmask.as<uint32_t>(), b->cvt.as<uint32_t>(), b->vexit.as<uint32_t>(), b->dict.as<uint16_t>(), 32 * K,       \
        b->evpool.as<uint2>(), (unsigned long long)(b->evpool.cap / 8), counter, b->rowinfo.as<uint2>(),               \
        b->rowtotal.as<uint32_t>(), b->rowbase.as<uint32_t>(), b->chunktotal.as<unsigned long long>(),                 \
        b->chunkbase.as<unsigned long long>(), b->hso.as<uint32_t>(), sync + 2
    if (fold) {
        shf::events_kernel<K, true><<<egrid, shf::kEventWarps * 32, 0, s>>>(SHF_EVENTS_ARGS);
    } else {
        shf::events_kernel<K, false><<<egrid, shf::kEventWarps * 32, 0, s>>>(SHF_EVENTS_ARGS);
        shf::bases_kernel<<<(g.n_chunks + 7u) / 8u, 256, 0, s>>>(g, b->rowtotal.as<uint32_t>(), b->rowbase.as<uint32_t>(),
                                                     b->chunktotal.as<unsigned long long>(),
                                                     b->chunkbase.as<unsigned long long>(), b->hso.as<uint32_t>(), counter,
                                                     sync + 2);
        tls_launches++;
    }
#undef SHF_EVENTS_ARGS
    tls_launches++;
    SHF_CUDA(cudaGetLastError());
    return SHF_OK;
}

int dispatch_events(int K, shf_buffer* b, const Geo& g, cudaStream_t s) {
    switch (K) {
        case 1: return launch_events<1>(b, g, s);
        case 2: return launch_events<2>(b, g, s);
        case 4: return launch_events<4>(b, g, s);
        case 8: return launch_events<8>(b, g, s);
    }
    return fail(SHF_ERR_UNSUPPORTED, "K", "no kernel instance");
}

// phase 0: vertical scan + event lists (bins per row); phase 1: emit
template <int K>
int launch_chain(shf_buffer* b, const Geo& g, cudaStream_t s, int phase) {
    if (phase == 0) {
        const uint32_t nblk = (g.PW + shf::kVscanThreads - 1) / shf::kVscanThreads;
        const dim3 vgrid(nblk, g.n_chunks, g.vseg);
        const size_t vsmem = vscan_smem(K);
        if (g.vseg > 1u) SHF_CUDA(b->vexit.ensure((size_t)g.n_chunks * (g.vseg - 1u) * nblk * 32 * K * 32 * 4));
        if (g.TY % 8u == 0u) {
            SHF_CUDA(cudaFuncSetAttribute(shf::vscan_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
            shf::vscan_kernel<K, true><<<vgrid, shf::kVscanThreads, vsmem, s>>>(g, b->ids, b->cvt.as<uint32_t>(),
                                                                               b->base.as<uint8_t>(), b->colmask.as<uint32_t>(),
                                                                               b->vexit.as<uint32_t>());
        } else {
            SHF_CUDA(cudaFuncSetAttribute(shf::vscan_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
            shf::vscan_kernel<K, false><<<vgrid, shf::kVscanThreads, vsmem, s>>>(g, b->ids, b->cvt.as<uint32_t>(),
                                                                                b->base.as<uint8_t>(), b->colmask.as<uint32_t>(),
                                                                                b->vexit.as<uint32_t>());
        }
        tls_launches++;
        SHF_CUDA(cudaGetLastError());
        SHF_CUDA(b->mark(2, s));
        return launch_events<K>(b, g, s);
    }
    const size_t smem = emit_smem(g.TY, g.R, K, (int)g.FW);
    Geo ge = g;
    ge.bins_cap = b->bins.cap / sizeof(shf_bin);
    ge.pool_cap = b->evpool.cap / 8;
    dim3 grid(g.T, g.emit_chunks, g.cseg);
    const uint32_t threads = (g.TY + g.producers) * 32;
    if (g.FW == 8u) SHF_CUDA(cudaFuncSetAttribute(shf::emit_kernel<K, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SHF_CUDA(cudaFuncSetAttribute(shf::emit_kernel<K, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (g.persist) {
        // persistent CTAs: as many as are resident at once (g.persist = SMs of the device), each walking the tile list
        int per_sm = 0;
        if (g.FW == 8u) SHF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shf::emit_kernel<K, 8>, (int)threads, smem));
        else SHF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shf::emit_kernel<K, 16>, (int)threads, smem));
        const uint64_t resident = (uint64_t)std::max(per_sm, 1) * (g.persist >> 1);
        grid = dim3((uint32_t)std::min<uint64_t>(resident, (uint64_t)g.T * g.emit_chunks), 1u, 1u);
    }
    if (g.FW == 8u) {
        shf::emit_kernel<K, 8><<<grid, threads, smem, s>>>(
            ge, b->ids, b->base.as<uint8_t>(), b->evpool.as<uint2>(), b->rowinfo.as<uint2>(),
            b->rowbase.as<uint32_t>(), b->chunkbase.as<uint64_t>(), b->bins.as<uint2>(), b->hso.as<uint32_t>());
    } else {
        shf::emit_kernel<K, 16><<<grid, threads, smem, s>>>(
            ge, b->ids, b->base.as<uint8_t>(), b->evpool.as<uint2>(), b->rowinfo.as<uint2>(),
            b->rowbase.as<uint32_t>(), b->chunkbase.as<uint64_t>(), b->bins.as<uint2>(), b->hso.as<uint32_t>());
    }
    tls_launches++;
    SHF_CUDA(cudaGetLastError());
    return SHF_OK;
}

int dispatch_chain(int K, shf_buffer* b, const Geo& g, cudaStream_t s, int phase) {
    switch (K) {
        case 1: return launch_chain<1>(b, g, s, phase);
        case 2: return launch_chain<2>(b, g, s, phase);
        case 4: return launch_chain<4>(b, g, s, phase);
        case 8: return launch_chain<8>(b, g, s, phase);
    }
    return fail(SHF_ERR_UNSUPPORTED, "K", "no kernel instance");
}

// Emit of a large call in two launches that run side by side: the bulk of the chunks as whole tiles on a stream of the
// highest priority, the last few chunks cut into column segments (items a quarter or half as long) on the call's own
// stream. Tiles are handed out dynamically, one CTA per SM, so a launch ends with the SMs finishing their last tile up to
// a whole tile time apart (~85 us of a 32-chunk launch's 790 us were this tail and the start-up); the low-priority
// launch's short items are only dispatched when no whole tile is waiting, i.e. into exactly those gaps.
// (dense histograms only -- `bins_per_pixel` from this call's totals, or from the buffer's last result when the call runs
// ahead of them: there the kernel waits for the HBM and the segments' extra window fill is free; sparse ones are bound by
// instruction issue and lose what the fill costs: +5 % on a blocky 32-chunk batch.)
int launch_emit_split(shf_filter* f, shf_buffer* b, const Geo& g, cudaStream_t s, double bins_per_pixel) {
    const uint64_t tiles = (uint64_t)g.T * g.n_chunks;
    uint32_t seg = f->dbg.debug_tailseg ? f->dbg.debug_tailseg : 4u;
    while (seg > 1u && (((g.W + seg - 1u) / seg + 15u) & ~15u) < std::max(64u, 2u * g.r)) seg--;
    // enough short items to cover the drain of the bulk launch: about one tile per two SMs, in whole chunks
    uint32_t tail_chunks = (uint32_t)std::min<uint64_t>(g.n_chunks / 4u, ((uint64_t)f->sm_count / 2u + g.T - 1u) / g.T);
    if (f->dbg.debug_tail) tail_chunks = std::min(f->dbg.debug_tail, g.n_chunks - 1u);   // measurements only
    if (g.persist || g.cseg != 1u || seg < 2u || tail_chunks == 0u || tiles < 4ull * (uint64_t)f->sm_count || f->dbg.no_split ||
        bins_per_pixel < 24.0)
        return dispatch_chain((int)g.K, b, g, s, 1);
    if (!b->emit_stream) {
        int least = 0, greatest = 0;
        SHF_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        SHF_CUDA(cudaStreamCreateWithPriority(&b->emit_stream, cudaStreamNonBlocking, greatest));
        SHF_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
        SHF_CUDA(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
    }
    SHF_CUDA(cudaEventRecord(b->ev_fork, s));
    SHF_CUDA(cudaStreamWaitEvent(b->emit_stream, b->ev_fork, 0));
    Geo bulk = g;
    bulk.emit_chunk0 = 0u;
    bulk.emit_chunks = g.n_chunks - tail_chunks;
    int st = dispatch_chain((int)g.K, b, bulk, b->emit_stream, 1);
    if (st != SHF_OK) return st;
    Geo tail = g;
    tail.emit_chunk0 = g.n_chunks - tail_chunks;
    tail.emit_chunks = tail_chunks;
    tail.cseg_px = ((g.W + seg - 1u) / seg + 15u) & ~15u;
    tail.cseg = (g.W + tail.cseg_px - 1u) / tail.cseg_px;
    st = dispatch_chain((int)g.K, b, tail, s, 1);
    if (st != SHF_OK) return st;
    SHF_CUDA(cudaEventRecord(b->ev_join, b->emit_stream));
    SHF_CUDA(cudaStreamWaitEvent(s, b->ev_join, 0));
    return SHF_OK;
}

// dictionary of every chunk: presence bitmap -> popcount prefix + number of distinct values (by the chunk's last CTA)
int launch_dictionary(shf_filter* f, shf_buffer* b, const Geo& g, const uint16_t* in_dev, bool vec8, cudaStream_t s) {
    const uint32_t n_chunks = g.n_chunks;
    SHF_CUDA(cudaMemsetAsync(b->bitmap.p, 0, (size_t)n_chunks * shf::kDictWords * 4, s));
    const dim3 pgrid(shf_rows_grid(g.PH, n_chunks, f->sm_count), n_chunks);
    uint32_t* done = b->sync.as<uint32_t>() + 3u + n_chunks;
    if (vec8)
        shf::presence_kernel<8><<<pgrid, 256, 0, s>>>(in_dev, g, b->bitmap.as<uint32_t>(), b->prefix.as<uint32_t>(),
                                                      b->nbiomes.as<uint32_t>(), done);
    else
        shf::presence_kernel<1><<<pgrid, 256, 0, s>>>(in_dev, g, b->bitmap.as<uint32_t>(), b->prefix.as<uint32_t>(),
                                                      b->nbiomes.as<uint32_t>(), done);
    tls_launches++;
    SHF_CUDA(cudaGetLastError());
    return SHF_OK;
}

// SHF.cpp:874-880
int validate(const uint32_t map_size[2], const uint32_t nn[2], uint32_t radius) {
    if (!(radius > 0u && (radius & 1u) == 0u))
        return fail(SHF_ERR_NUMERIC_DOMAIN, "radius > 0u && (radius & 0x01u) == 0x00u",
                    "The radius of the filter kernel must be a positive even number");
    const uint64_t sx = (uint64_t)map_size[0] * (nn[0] / 2u), sy = (uint64_t)map_size[1] * (nn[1] / 2u);
    if (radius > sx || radius > sy)
        return fail(SHF_ERR_NUMERIC_DOMAIN, "radius <= start_coord",
                    "The radius of the filter kernel is too large and will overflow the free-slip boundary");
    return SHF_OK;
}

// scratch every path needs + the compact-id map and the dictionary
int prepare_common(shf_filter* f, shf_buffer* b, const Geo& g, const uint16_t* in_dev, bool vec8, cudaStream_t s) {
    const size_t cells = (size_t)g.n_chunks * g.PH * g.P;
    SHF_CUDA(b->cmap.ensure(cells * 2 + 64));  // producers read whole 16/32-byte row segments
    SHF_CUDA(b->dict.ensure((size_t)g.n_chunks * g.Bpad * 2));
    SHF_CUDA(b->rowtotal.ensure((size_t)g.n_chunks * g.H * 4));
    SHF_CUDA(b->rowbase.ensure((size_t)g.n_chunks * g.H * 4));
    SHF_CUDA(b->chunktotal.ensure((size_t)(g.n_chunks + 1) * 8));  // + the event counter
    SHF_CUDA(b->chunkbase.ensure((size_t)(g.n_chunks + 1) * 8));
    SHF_CUDA(b->hso.ensure((size_t)g.n_chunks * ((size_t)g.W * g.H + 1u) * 4));
    if (g.small_ids) {   // the sample values are the compact ids: the kernels read the caller's maps in place
        b->ids = in_dev;
        return SHF_OK;
    }
    b->ids = b->cmap.as<uint16_t>();
    const dim3 pgrid(shf_rows_grid(g.PH, g.n_chunks, f->sm_count), g.n_chunks);
    if (vec8)
        shf::remap_kernel<8><<<pgrid, 256, 0, s>>>(in_dev, g, b->bitmap.as<uint32_t>(), b->prefix.as<uint32_t>(),
                                                   b->cmap.as<uint16_t>(), b->dict.as<uint16_t>(), g.Bpad);
    else
        shf::remap_kernel<1><<<pgrid, 256, 0, s>>>(in_dev, g, b->bitmap.as<uint32_t>(), b->prefix.as<uint32_t>(),
                                                   b->cmap.as<uint16_t>(), b->dict.as<uint16_t>(), g.Bpad);
    tls_launches++;
    SHF_CUDA(cudaGetLastError());
    return SHF_OK;
}

// bins per row -> first bin of every row, chunk totals on the host (u32 overflow check), the bin buffer
int size_output(shf_buffer* b, const Geo& g, uint64_t* h_totals, cudaStream_t s, bool with_events = false) {
    const uint32_t n_chunks = g.n_chunks;
    SHF_CUDA(b->mark(3, s));
    if (!with_events) {   // (events_kernel has scanned its own row totals)
        shf::rowscan_kernel<<<n_chunks, 1024, 0, s>>>(g, b->rowtotal.as<uint32_t>(), b->rowbase.as<uint32_t>(),
                                                      b->chunktotal.as<unsigned long long>(), b->hso.as<uint32_t>());
        tls_launches++;
        SHF_CUDA(cudaGetLastError());
    }
    SHF_CUDA(b->mark(4, s));
    const size_t n_read = (size_t)n_chunks + (with_events ? 1u : 0u);
    SHF_CUDA(cudaMemcpyAsync(h_totals, b->chunktotal.p, n_read * 8, cudaMemcpyDeviceToHost, s));
    tls_d2h += n_read * 8;
    SHF_CUDA(cudaStreamSynchronize(s));
    if (with_events) {
        // the event pool is sized by guess; a call that needs more grows it and builds the event lists again
        const uint64_t need = h_totals[n_chunks];
        if (need > 0xFFFFFFFFull)
            return fail(SHF_ERR_UNSUPPORTED, "presence chains per call < 2^32", "too many bin births in one batch");
        if (need > b->evpool.cap / 8) {
            SHF_CUDA(b->evpool.ensure((size_t)(need + need / 8u + 1024u) * 8));
            const int st = dispatch_events((int)g.K, b, g, s);   // (same totals; chunkbase is written again as well)
            if (st != SHF_OK) return st;
        }
    }
    b->chunk_base.assign(n_chunks + 1u, 0ull);
    for (uint32_t i = 0; i < n_chunks; i++) {
        if (h_totals[i] > 0xFFFFFFFFull)
            return fail(SHF_ERR_OFFSET_OVERFLOW, "bins per chunk < 2^32",
                        "HistogramStartOffset is 32 bits wide; this chunk has too many bins");
        b->chunk_base[i + 1u] = b->chunk_base[i] + h_totals[i];
    }
    const size_t total = (size_t)b->chunk_base[n_chunks];
    SHF_CUDA(b->bins.ensure(std::max<size_t>(total, 1) * sizeof(shf_bin)));
    if (!with_events) {   // (events_kernel has left the same numbers in chunkbase)
        SHF_CUDA(cudaMemcpyAsync(b->chunkbase.p, b->chunk_base.data(), (size_t)(n_chunks + 1) * 8, cudaMemcpyHostToDevice, s));
        tls_h2d += (size_t)(n_chunks + 1) * 8;
    }
    return SHF_OK;
}

int publish(shf_buffer* b, const Geo& g) {
    b->ev_valid = g_profiling != 0;
    b->n_chunks = g.n_chunks;
    b->last_w = g.W;
    b->last_h = g.H;
    b->n_bins = (size_t)b->chunk_base[g.n_chunks];
    b->n_offsets = (size_t)g.n_chunks * ((size_t)g.W * g.H + 1u);
    b->has_result = true;
    b->dev_valid = true;
    b->on_host = false;
    return SHF_OK;
}

// The wide path (shf_generic.cuh): any radius, up to kGenericMaxBiomes distinct values.
int run_generic(shf_filter* f, shf_buffer* b, const Geo& g, const uint16_t* in_dev, bool vec8, uint32_t bmax,
                uint64_t* h_totals, cudaStream_t s) {
    const uint32_t nb = g.Bpad;
    if (bmax > shf::kGenericMaxBiomes)
        return fail(SHF_ERR_UNSUPPORTED, "distinct samples per chunk <= 16384",
                    "too many distinct sample values in one neighbourhood for the per-row shared-memory tables");
    const size_t smem = shf::march_generic_smem(nb, g.span);
    if (smem > f->smem_optin)
        return fail(SHF_ERR_UNSUPPORTED, "per-row tables fit shared memory", "radius x biome count too large for one CTA");
    b->plan_smem = (uint32_t)smem;
    SHF_CUDA(b->vstart.ensure((size_t)g.n_chunks * g.PH * g.P * 2));
    int st = prepare_common(f, b, g, in_dev, vec8, s);
    if (st != SHF_OK) return st;
    // vertical chain starts; per-(biome, column) state for a sub-batch of chunks at a time (<= 1 GiB of scratch)
    const size_t per_chunk = (size_t)nb * g.PW * 4;
    const uint32_t sub = (uint32_t)std::max<size_t>(1, std::min<size_t>(g.n_chunks, (size_t(1) << 30) / per_chunk));
    SHF_CUDA(b->gstate.ensure(per_chunk * sub));
    for (uint32_t first = 0; first < g.n_chunks; first += sub) {
        const uint32_t cnt = std::min(sub, g.n_chunks - first);
        SHF_CUDA(cudaMemsetAsync(b->gstate.p, 0xFF, per_chunk * cnt, s));
        shf::vstart_generic_kernel<<<dim3((g.PW + 127u) / 128u, cnt), 128, 0, s>>>(
            g, first, nb, b->cmap.as<uint16_t>(), b->vstart.as<uint16_t>(), b->gstate.as<uint32_t>());
        tls_launches++;
        SHF_CUDA(cudaGetLastError());
    }
    SHF_CUDA(b->mark(2, s));
    SHF_CUDA(cudaFuncSetAttribute(shf::march_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SHF_CUDA(cudaFuncSetAttribute(shf::march_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(g.H, g.n_chunks);
    shf::march_generic_kernel<false><<<grid, shf::kGenericThreads, smem, s>>>(
        g, nb, b->cmap.as<uint16_t>(), b->vstart.as<uint16_t>(), b->dict.as<uint16_t>(), nb, nullptr, nullptr, nullptr,
        nullptr, b->rowtotal.as<uint32_t>());
    tls_launches++;
    SHF_CUDA(cudaGetLastError());
    st = size_output(b, g, h_totals, s);
    if (st != SHF_OK) return st;
    SHF_CUDA(b->mark(5, s));
    shf::march_generic_kernel<true><<<grid, shf::kGenericThreads, smem, s>>>(
        g, nb, b->cmap.as<uint16_t>(), b->vstart.as<uint16_t>(), b->dict.as<uint16_t>(), nb, b->rowbase.as<uint32_t>(),
        b->chunkbase.as<uint64_t>(), b->bins.as<uint2>(), b->hso.as<uint32_t>(), nullptr);
    tls_launches++;
    SHF_CUDA(cudaGetLastError());
    SHF_CUDA(b->mark(6, s));
    return publish(b, g);
}

// Emit + copy of a result that goes to the host (north_star: "result copies overlap compute via pinned async transfers
// on per-GPU streams"): the chunks are emitted in up to four ranges; as soon as a range is done its bins and offsets start
// towards the buffer's page-locked memory on a second stream while the next range is emitted. b->chunk_base is known
// (the caller has seen the totals), the bin buffer is large enough.
int emit_to_host(shf_buffer* b, Geo g, cudaStream_t s) {
    const uint32_t n_chunks = g.n_chunks;
    const size_t n_bins = (size_t)b->chunk_base[n_chunks];
    const size_t per = (size_t)g.W * g.H + 1u;
    SHF_CUDA(b->h_bins.ensure(std::max<size_t>(n_bins, 1) * sizeof(shf_bin)));
    SHF_CUDA(b->h_hso.ensure((size_t)n_chunks * per * 4));
    if (!b->copy_stream) SHF_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    const uint32_t n_ranges = std::min<uint32_t>(4u, n_chunks);
    SHF_CUDA(b->mark(5, s));
    for (uint32_t i = 0; i < n_ranges; i++) {
        const uint32_t c0 = (uint32_t)((uint64_t)n_chunks * i / n_ranges), c1 = (uint32_t)((uint64_t)n_chunks * (i + 1u) / n_ranges);
        g.emit_chunk0 = c0;
        g.emit_chunks = c1 - c0;
        const int st = dispatch_chain((int)g.K, b, g, s, 1);
        if (st != SHF_OK) return st;
        if (!b->range_ev[i]) SHF_CUDA(cudaEventCreateWithFlags(&b->range_ev[i], cudaEventDisableTiming));
        SHF_CUDA(cudaEventRecord(b->range_ev[i], s));
        SHF_CUDA(cudaStreamWaitEvent(b->copy_stream, b->range_ev[i], 0));
        const size_t b0 = (size_t)b->chunk_base[c0], b1 = (size_t)b->chunk_base[c1];
        if (b1 > b0) {
            SHF_CUDA(cudaMemcpyAsync(b->h_bins.as<shf_bin>() + b0, b->bins.as<shf_bin>() + b0, (b1 - b0) * sizeof(shf_bin),
                                     cudaMemcpyDeviceToHost, b->copy_stream));
            tls_d2h += (b1 - b0) * sizeof(shf_bin);
        }
        SHF_CUDA(cudaMemcpyAsync(b->h_hso.as<uint32_t>() + c0 * per, b->hso.as<uint32_t>() + c0 * per, (c1 - c0) * per * 4,
                                 cudaMemcpyDeviceToHost, b->copy_stream));
        tls_d2h += (c1 - c0) * per * 4;
    }
    SHF_CUDA(b->mark(6, s));
    SHF_CUDA(cudaStreamSynchronize(b->copy_stream));
    return SHF_OK;
}

// A repeated call of the same shape on the same buffer (the pipeline's pooled buffers see chunk after chunk of one
// geometry) runs the whole kernel sequence with the previous call's plan and buffer sizes and only then looks at the
// numbers the checked path waits for twice (distinct values per chunk, bins per chunk, events): no host round trip
// inside the sequence. Every kernel stays inside its buffers whatever the data (compact ids are clamped, events beyond
// the pool and chunks beyond the bin buffer are skipped). The look happens in settle(): right away for the blocking
// entry points, at the first query of the result for shf_run_device_async; when a check fails the result is discarded
// and the checked path runs on the same input.
int enqueue_speculative(shf_filter* f, shf_buffer* b, const Geo& g, const uint16_t* in_dev, bool vec8, uint32_t* h_nbiomes,
                        uint64_t* h_totals, cudaStream_t s, bool with_emit = true) {
    const uint32_t n_chunks = g.n_chunks;
    const int K = (int)g.K;
    b->ev_valid = false;
    SHF_CUDA(b->mark(0, s));
    int st;
    const bool side_dictionary = g.small_ids != 0u;   // (its phase then reads ~0 in the profile: it is off the stream)
    if (side_dictionary) {
        // nothing downstream reads the dictionary when the sample values are the ids: it only has to confirm that they
        // still are (distinct values, largest value), so it runs beside the scan instead of in front of it
        if (!b->aux_stream) {
            SHF_CUDA(cudaStreamCreateWithFlags(&b->aux_stream, cudaStreamNonBlocking));
            SHF_CUDA(cudaEventCreateWithFlags(&b->ev_aux_fork, cudaEventDisableTiming));
            SHF_CUDA(cudaEventCreateWithFlags(&b->ev_aux_join, cudaEventDisableTiming));
        }
        SHF_CUDA(cudaEventRecord(b->ev_aux_fork, s));
        SHF_CUDA(cudaStreamWaitEvent(b->aux_stream, b->ev_aux_fork, 0));
        st = launch_dictionary(f, b, g, in_dev, vec8, b->aux_stream);
        if (st != SHF_OK) return st;
        SHF_CUDA(cudaEventRecord(b->ev_aux_join, b->aux_stream));
    } else {
        st = launch_dictionary(f, b, g, in_dev, vec8, s);
        if (st != SHF_OK) return st;
    }
    SHF_CUDA(b->mark(1, s));
    st = prepare_common(f, b, g, in_dev, vec8, s);
    if (st != SHF_OK) return st;
    st = dispatch_chain(K, b, g, s, 0);   // (events_kernel leaves the row and chunk bases behind)
    if (st != SHF_OK) return st;
    SHF_CUDA(b->mark(3, s));
    if (side_dictionary && with_emit) {
        if (!b->ev_totals) SHF_CUDA(cudaEventCreateWithFlags(&b->ev_totals, cudaEventDisableTiming));
        SHF_CUDA(cudaEventRecord(b->ev_totals, s));   // the event lists and the totals are done
    }
    SHF_CUDA(b->mark(4, s));
    if (with_emit) {
        SHF_CUDA(b->mark(5, s));
        // (density of the buffer's last result: the pipeline filters chunk after chunk of one world)
        const double last_density = b->n_offsets ? (double)b->n_bins / (double)b->n_offsets : 0.0;
        st = launch_emit_split(f, b, g, s, last_density);
        if (st != SHF_OK) return st;
        SHF_CUDA(b->mark(6, s));
    }
    if (side_dictionary && with_emit) {
        // the numbers settle() looks at travel on the side stream as well: behind the dictionary (already there) and the
        // event lists, beside the emitting kernel -- nothing of this call is left on the caller's stream after emit, so the
        // next call's scan starts the moment emit ends
        SHF_CUDA(cudaStreamWaitEvent(b->aux_stream, b->ev_totals, 0));
        SHF_CUDA(cudaMemcpyAsync(h_nbiomes, b->nbiomes.p, (size_t)n_chunks * 8, cudaMemcpyDeviceToHost, b->aux_stream));
        SHF_CUDA(cudaMemcpyAsync(h_totals, b->chunktotal.p, ((size_t)n_chunks + 1u) * 8, cudaMemcpyDeviceToHost, b->aux_stream));
        SHF_CUDA(cudaEventRecord(b->ev_aux_join, b->aux_stream));
        b->aux_pending = true;
    } else {
        if (side_dictionary) SHF_CUDA(cudaStreamWaitEvent(s, b->ev_aux_join, 0));
        SHF_CUDA(cudaMemcpyAsync(h_nbiomes, b->nbiomes.p, (size_t)n_chunks * 8, cudaMemcpyDeviceToHost, s));
        SHF_CUDA(cudaMemcpyAsync(h_totals, b->chunktotal.p, ((size_t)n_chunks + 1u) * 8, cudaMemcpyDeviceToHost, s));
    }
    tls_d2h += (size_t)n_chunks * 16 + 8;
    return SHF_OK;
}

// page-locked scratch of a call: distinct values and largest value per chunk (2 x u32), bins per chunk + event count (u64)
void small_host_views(shf_buffer* b, uint32_t n_chunks, uint32_t** h_nbiomes, uint64_t** h_totals) {
    *h_nbiomes = b->h_small.as<uint32_t>();
    *h_totals = reinterpret_cast<uint64_t*>(b->h_small.as<uint8_t>() + (((size_t)n_chunks * 8 + 15) & ~size_t(15)));
}

int run_checked(shf_filter* f, shf_buffer* b, Geo g, const uint16_t* in_dev, bool vec8, cudaStream_t s, bool to_host = false);
int result_to_host(shf_buffer* b, cudaStream_t s);

// Completes a call enqueued by enqueue_speculative: waits for it, makes the checks the other path makes before it
// launches and publishes the result -- or runs the checked path when the previous call's plan or buffer sizes did not
// fit this input. SHF_OK without work when nothing is pending.
// the checks of a call that ran ahead with the plan in `g`: does the plan fit what the kernels found? On a hit
// b->chunk_base and b->plan_biomes are set. emitted: the emitting kernel ran too (the bin buffer must have been large enough).
bool speculation_holds(shf_buffer* b, const Geo& g, bool emitted) {
    const uint32_t n_chunks = g.n_chunks;
    uint32_t* h_nbiomes;
    uint64_t* h_totals;
    small_host_views(b, n_chunks, &h_nbiomes, &h_totals);
    uint32_t bmax = 0, vmax = 0;
    for (uint32_t i = 0; i < n_chunks; i++) {
        bmax = std::max(bmax, h_nbiomes[i]);
        vmax = std::max(vmax, h_nbiomes[n_chunks + i]);
    }
    const uint32_t need_k = bmax <= 32u ? 1u : bmax <= 64u ? 2u : bmax <= 128u ? 4u : 8u;
    bool hit = bmax <= 256u && need_k == g.K;                              // else another plan (or the wide path) is due
    hit = hit && (!g.small_ids || vmax < 32u * g.K);                       // sample values still usable as compact ids
    hit = hit && h_totals[n_chunks] <= b->evpool.cap / 8;                  // event pool large enough
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_chunks && hit; i++) {
        hit = h_totals[i] <= 0xFFFFFFFFull;                                // (the checked path reports the overflow)
        total += h_totals[i];
    }
    hit = hit && (!emitted || total <= b->bins.cap / sizeof(shf_bin));     // bin buffer large enough
    if (!hit) return false;
    b->chunk_base.assign(n_chunks + 1u, 0ull);
    for (uint32_t i = 0; i < n_chunks; i++) b->chunk_base[i + 1u] = b->chunk_base[i] + h_totals[i];
    b->plan_biomes = bmax;
    return true;
}

int settle(shf_buffer* b) {
    if (!b->deferred.active) return SHF_OK;
    shf_buffer::Deferred d = b->deferred;
    b->deferred.active = false;
    SHF_CUDA(cudaEventSynchronize(b->done_ev));
    if (b->aux_pending) {
        SHF_CUDA(cudaEventSynchronize(b->ev_aux_join));
        b->aux_pending = false;
    }
    const Geo& g = d.g;
    if (!speculation_holds(b, g, true)) {
        b->last_valid = false;
        b->spec_misses++;
        const int st = run_checked(&d.filt, b, g, d.in_dev, d.vec8, d.stream);
        if (st == SHF_OK) SHF_CUDA(cudaEventRecord(b->done_ev, d.stream));
        return st;
    }
    return publish(b, g);
}

// Run the whole kernel sequence. `in_dev` views the halo-extended region of every chunk in device memory.
// deferred: return as soon as everything is enqueued when the call can run ahead (settle() completes it later).
int run_on_device(shf_filter* f, shf_buffer* b, const uint16_t* in_dev, uint64_t in_chunk_stride, uint32_t in_row_stride,
                  uint32_t n_chunks, uint32_t W, uint32_t H, uint32_t r, cudaStream_t s, bool deferred = false,
                  bool to_host = false) {
    Geo g{};
    g.W = W;
    g.H = H;
    g.r = r;
    g.span = 2u * r + 1u;
    g.PW = W + 2u * r;
    g.PH = H + 2u * r;
    g.P = (g.PW + 7u) & ~7u;
    g.n_chunks = n_chunks;
    g.in_row_stride = in_row_stride;
    g.in_chunk_stride = in_chunk_stride;
    g.inv_total = 1.0f / (float)(g.span * g.span);
    if (W == 0u || H == 0u || n_chunks == 0u) return fail(SHF_ERR_INVALID_ARGUMENT, "W*H*n_chunks > 0", "empty input");
    if (n_chunks > 65535u) return fail(SHF_ERR_UNSUPPORTED, "n_chunks <= 65535", "too many chunks in one batch");
    if (g.PH >= 65535u) return fail(SHF_ERR_UNSUPPORTED, "H + 2*radius < 65535", "map too tall for 16-bit row keys");

    // A call still waiting for its checks is simply dropped: this call replaces its result. Work of an earlier call may
    // still be running on another stream (the device entry points leave their last kernels in flight on the caller's
    // stream): this call's stream waits for it before it touches the buffer's scratch.
    b->deferred.active = false;
    b->has_result = false;
    if (!b->done_ev) SHF_CUDA(cudaEventCreateWithFlags(&b->done_ev, cudaEventDisableTiming));
    if (b->done_recorded && b->done_stream != s) SHF_CUDA(cudaStreamWaitEvent(s, b->done_ev, 0));
    if (b->aux_pending) SHF_CUDA(cudaStreamWaitEvent(s, b->ev_aux_join, 0));   // (read-backs of the call before, long done)
    if ((size_t)n_chunks * 24 + 64 > b->h_small.cap && b->done_recorded) {
        SHF_CUDA(cudaEventSynchronize(b->done_ev));  // (a copy into the old page-locked block may be in flight)
        if (b->aux_pending) SHF_CUDA(cudaEventSynchronize(b->ev_aux_join));
    }

    SHF_CUDA(b->bitmap.ensure((size_t)n_chunks * shf::kDictWords * 4));
    SHF_CUDA(b->prefix.ensure((size_t)n_chunks * shf::kDictWords * 4));
    SHF_CUDA(b->nbiomes.ensure((size_t)n_chunks * 8));
    SHF_CUDA(b->sync.ensure_zeroed(((size_t)n_chunks * 2 + 4) * 4, s));   // arrival counters, left at zero by their kernels
    SHF_CUDA(b->h_small.ensure((size_t)n_chunks * 24 + 64));
    uint32_t* h_nbiomes;
    uint64_t* h_totals;
    small_host_views(b, n_chunks, &h_nbiomes, &h_totals);

    // 16-byte loads when the halo view allows it
    const bool vec8 = (reinterpret_cast<uintptr_t>(in_dev) % 16u == 0u) && (in_row_stride % 8u == 0u) && (in_chunk_stride % 8u == 0u);

    // ---- same shape as the last call on this buffer: run with its plan, check afterwards ----
    int st;
    if (b->last_valid && b->last_g.W == W && b->last_g.H == H && b->last_g.r == r && b->last_g.n_chunks == n_chunks &&
        !f->dbg.no_speculation) {
        Geo gs = b->last_g;
        gs.in_row_stride = in_row_stride;
        gs.in_chunk_stride = in_chunk_stride;
        if (gs.small_ids && !(vec8 && in_row_stride >= gs.PW + 32u)) {   // this view cannot be read in place
            gs.small_ids = 0u;
        }
        gs.ids_row_stride = gs.small_ids ? in_row_stride : gs.P;
        gs.ids_chunk_stride = gs.small_ids ? in_chunk_stride : (uint64_t)gs.PH * gs.P;
        if (to_host) {
            // the result goes to the host: run ahead up to the event lists, look at the totals (the page-locked output
            // must be sized anyway), then emit in ranges whose copies overlap the next range's kernel
            st = enqueue_speculative(f, b, gs, in_dev, vec8, h_nbiomes, h_totals, s, false);
            if (st != SHF_OK) return st;
            SHF_CUDA(cudaStreamSynchronize(s));
            if (speculation_holds(b, gs, false)) {
                SHF_CUDA(b->bins.ensure(std::max<size_t>((size_t)b->chunk_base[n_chunks], 1) * sizeof(shf_bin)));
                st = emit_to_host(b, gs, s);
                if (st != SHF_OK) return st;
                SHF_CUDA(cudaEventRecord(b->done_ev, s));
                b->done_recorded = true;
                b->done_stream = s;
                st = publish(b, gs);
                b->on_host = true;
                return st;
            }
            b->last_valid = false;
            b->spec_misses++;
            st = run_checked(f, b, g, in_dev, vec8, s, true);
            if (st != SHF_OK) return st;
            SHF_CUDA(cudaEventRecord(b->done_ev, s));
            b->done_recorded = true;
            b->done_stream = s;
            return SHF_OK;
        }
        st = enqueue_speculative(f, b, gs, in_dev, vec8, h_nbiomes, h_totals, s);
        if (st != SHF_OK) return st;
        SHF_CUDA(cudaEventRecord(b->done_ev, s));
        b->done_recorded = true;
        b->done_stream = s;
        b->deferred.active = true;
        b->deferred.filt = *f;
        b->deferred.g = gs;
        b->deferred.in_dev = in_dev;
        b->deferred.vec8 = vec8;
        b->deferred.stream = s;
        return deferred ? SHF_OK : settle(b);
    }
    st = run_checked(f, b, g, in_dev, vec8, s, to_host);
    if (st != SHF_OK) return st;
    SHF_CUDA(cudaEventRecord(b->done_ev, s));
    b->done_recorded = true;
    b->done_stream = s;
    return SHF_OK;
}

// The checked path: two host round trips (distinct values per chunk -> plan; bins per chunk -> bin buffer).
int run_checked(shf_filter* f, shf_buffer* b, Geo g, const uint16_t* in_dev, bool vec8, cudaStream_t s, bool to_host) {
    const uint32_t n_chunks = g.n_chunks, W = g.W, H = g.H, r = g.r;
    (void)r;
    uint32_t* h_nbiomes;
    uint64_t* h_totals;
    small_host_views(b, n_chunks, &h_nbiomes, &h_totals);
    // ---- dictionary ----
    b->ev_valid = false;
    SHF_CUDA(b->mark(0, s));
    {
        const int st = launch_dictionary(f, b, g, in_dev, vec8, s);
        if (st != SHF_OK) return st;
    }
    SHF_CUDA(b->mark(1, s));
    SHF_CUDA(cudaMemcpyAsync(h_nbiomes, b->nbiomes.p, (size_t)n_chunks * 8, cudaMemcpyDeviceToHost, s));
    tls_d2h += (size_t)n_chunks * 8;
    SHF_CUDA(cudaStreamSynchronize(s));
    uint32_t bmax = 0, vmax = 0;
    for (uint32_t i = 0; i < n_chunks; i++) {
        bmax = std::max(bmax, h_nbiomes[i]);
        vmax = std::max(vmax, h_nbiomes[n_chunks + i]);
    }
    if (bmax > 65535u) return fail(SHF_ERR_UNSUPPORTED, "distinct samples per chunk <= 65535", "compact ids are 16 bits wide");
    const int K = bmax <= 32u ? 1 : bmax <= 64u ? 2 : bmax <= 128u ? 4 : 8;
    g.K = K;
    g.Bpad = (bmax + 31u) & ~31u;
    // The event-list path covers up to 256 distinct values and 2r+1 <= 511 (8-bit vertical window counts up to 255,
    // 16-bit ones beyond) as long as one row's ring fits shared memory; everything else takes the wide path.
    g.FW = g.span <= 255u ? 8u : 16u;
    // (the event records pack pixel columns into 16 bits: wider maps take the wide path, which does not)
    bool generic = bmax > 256u || g.span > 511u || g.PW > 65535u;
    if (!generic) {
        g.Bpad = 32u * K;
        // emit kernel plan: rows per CTA (<= 16), producer warps, ring depth. A batch of 16 columns is produced in
        // `ppb` passes; with `np` producer warps ceil(np / ppb) batches are in production at once, and the ring must
        // hold one more batch than that besides the 2r+1 columns the consumers still read.
        const uint32_t ppb = (uint32_t)K * g.FW / 8u;
        uint32_t ty = std::min<uint32_t>(16u, H);
        auto plan = [&](uint32_t np, uint32_t extra) {
            g.producers = np;
            g.stages = (np + ppb - 1u) / ppb + extra;
            g.stages_magic = (uint32_t)((0x100000000ull + g.stages - 1u) / g.stages);
            g.R = g.span + shf::kMarchNB * g.stages;
        };
        auto smem_of = [&](uint32_t t) { return emit_smem(t, g.R, K, (int)g.FW); };
        // Keep the four producer warps and give up rows per CTA first (a starved consumer warp costs more than a smaller
        // tile): of the two ring depths with four producers take the one that fits more rows; fewer producers only
        // when not even two rows fit with four.
        const uint32_t ty_max = ty;
        auto rows_that_fit = [&](uint32_t np, uint32_t extra) {
            plan(np, extra);
            uint32_t t = ty_max;
            while (t > 1u && smem_of(t) > f->smem_optin) t--;
            return smem_of(t) <= f->smem_optin ? t : 0u;
        };
        const uint32_t deep = rows_that_fit(4u, 2u), shallow = rows_that_fit(4u, 1u);
        if (std::max(deep, shallow) >= std::min(2u, ty_max)) {
            ty = std::max(deep, shallow);
            plan(4u, deep >= shallow ? 2u : 1u);
        } else if ((ty = rows_that_fit(2u, 1u)) == 0u) {
            ty = std::max(1u, rows_that_fit(1u, 1u));   // (does not fit either: the check below sends it to the wide path)
        }
        // small calls: 8 rows per CTA when 16 would leave more than half of the SMs idle (a single 512x512 chunk has 32
        // tiles of 16 rows: emit 0.054 -> 0.046 ms; a 2048x2048 chunk with its 128 tiles is better off with 16)
        // (8, not fewer: vscan dumps a base vector per tile and its fast variant wants multiples of 8)
        if (f->dbg.debug_np) plan(std::min(f->dbg.debug_np, 4u), std::max(1u, f->dbg.debug_extra));  // measurements only
        if (f->dbg.debug_ty) ty = std::max(1u, std::min(ty, f->dbg.debug_ty));  // measurements only
        else if (ty > 8u && (uint64_t)n_chunks * ((H + ty - 1u) / ty) * 2u <= (uint64_t)f->sm_count) ty = 8u;
        if (smem_of(ty) > f->smem_optin) {
            generic = true;
            g.Bpad = (bmax + 31u) & ~31u;
        } else {
            g.TY = ty;
            g.T = (H + ty - 1u) / ty;
            b->plan_smem = (uint32_t)smem_of(ty);
        }
        // emit walks every row left to right, one consumer warp per row: a small call (a single 1024x1024 chunk has 128
        // tiles of 8 rows) is split into column segments, each sliding over the 2r columns before its first pixel
        g.cseg = 1u;
        g.cseg_px = W;
        if (!generic && !f->dbg.no_cseg) {
            const uint64_t ctas = (uint64_t)n_chunks * g.T;
            // (measured: 2 segments for 64..147 CTAs, 4 below; nothing to gain when one CTA fills an SM's shared memory)
            uint32_t want = 2u * smem_of(ty) > f->smem_optin ? 1u
                            : ctas * 2u <= (uint64_t)f->sm_count ? 4u
                            : ctas < (uint64_t)f->sm_count ? 2u : 1u;
            if (f->dbg.debug_cseg) want = f->dbg.debug_cseg;  // measurements only
            want = std::min(want, 8u);
            while (want > 1u) {
                const uint32_t px = ((W + want - 1u) / want + 15u) & ~15u;
                if (px >= std::max(64u, 2u * r)) {
                    g.cseg_px = px;
                    g.cseg = (W + px - 1u) / px;
                    break;
                }
                want--;
            }
        }
        // large calls: persistent emit CTAs walking the flat tile list out of phase with each other (EmitItem)
        // (opt-in: a static tile list per CTA measured 9 % slower on a 256-chunk batch than one CTA per tile, which the
        // hardware hands out dynamically -- with or without the phase shift; kept for experiments, SHF_PERSIST=1)
        g.persist = 0u;
        if (!generic && g.cseg == 1u && W >= 64u && (uint64_t)n_chunks * g.T >= 4ull * (uint64_t)f->sm_count && f->dbg.persist)
            g.persist = 2u * (uint32_t)f->sm_count + (f->dbg.persist > 1u ? 0u : 1u);
        // (g.persist = 2 * SMs to occupy + 1 if every CTA's first tile is cut in two; tests: persistent CTAs on few "SMs")
        if (!generic && g.cseg == 1u && f->dbg.debug_persist) g.persist = 2u * f->dbg.debug_persist + 1u;
        // vscan walks every column top to bottom, one warp per 32 columns: a small call (a single 1024x1024 chunk has 36
        // such warps) is split into up to eight row segments per block, each replaying the 2r+1 rows above its first
        g.vseg = 1u;
        g.vseg_rows = 0u;
        // (the kernel is a per-thread latency chain: as long as all its warps are resident at once, more and shorter
        // ones finish sooner although every segment adds 2r+1 replayed rows)
        const uint64_t vwarps = (uint64_t)n_chunks * ((g.PW + 31u) / 32u);
        const uint64_t vresident = (uint64_t)f->sm_count * std::min<uint64_t>(32u, f->smem_per_sm / (vscan_smem(K) + 1024u));
        if (!generic && vwarps * 2u <= vresident && !f->dbg.no_vseg) {
            uint32_t S = 8u;
            if (f->dbg.debug_vseg) S = f->dbg.debug_vseg;  // measurements only
            while (S > 1u) {
                const uint32_t L = (((H - 1u) + S - 1u) / S + 31u) & ~31u;   // output rows per segment
                if (L >= std::max(64u, g.span / 2u) && 1u + (S - 1u) * L < H && (vwarps * S <= vresident || f->dbg.debug_vseg)) {
                    g.vseg = S;
                    g.vseg_rows = L;
                    break;
                }
                S--;
            }
        }
    }
    // Sample values that are small numbers (biome ids counted from 0, as the demo's registry does) serve as compact ids
    // themselves when that costs no wider plan: no remapped copy, the kernels read the caller's maps where they lie
    // (16-byte aligned rows with room for the producers' whole-segment loads behind the halo).
    g.small_ids = !generic && vec8 && vmax < 32u * (uint32_t)K && g.in_row_stride >= g.PW + 32u && !f->dbg.no_small_ids;
    g.ids_row_stride = g.small_ids ? g.in_row_stride : g.P;
    g.ids_chunk_stride = g.small_ids ? g.in_chunk_stride : (uint64_t)g.PH * g.P;
    b->plan_k = generic ? 0u : (uint32_t)K;
    b->plan_ty = generic ? 1u : g.TY;
    b->plan_biomes = bmax;
    g.emit_chunk0 = 0u;
    g.emit_chunks = n_chunks;
    if (generic) {
        b->last_valid = false;
        const int st = run_generic(f, b, g, in_dev, vec8, bmax, h_totals, s);
        return st == SHF_OK && to_host ? result_to_host(b, s) : st;
    }

    SHF_CUDA(b->base.ensure((size_t)n_chunks * g.T * g.PW * g.Bpad * g.FW / 8u));
    SHF_CUDA(b->colmask.ensure((size_t)n_chunks * H * ((g.PW + 31u) / 32u) * 32u * K * 4));  // transposed masks per 32-column block
    {
        g.cv_pad = (32u - g.span % 32u) % 32u;
        g.cv_pitch = (g.PH + g.cv_pad + 31u) & ~31u;
        SHF_CUDA(b->cvt.ensure((size_t)n_chunks * g.PW * g.cv_pitch * 4));
        SHF_CUDA(b->rowinfo.ensure((size_t)n_chunks * H * 8));
        SHF_CUDA(b->evpool.ensure((size_t)n_chunks * H * 32u * (K + 1) * 8));
    }
    int st = prepare_common(f, b, g, in_dev, vec8, s);
    if (st != SHF_OK) return st;

    // ---- vertical scan + bins per row ----
    st = dispatch_chain(K, b, g, s, 0);
    if (st != SHF_OK) return st;
    st = size_output(b, g, h_totals, s, true);
    if (st != SHF_OK) return st;

    // ---- emit ----
    if (to_host) {
        st = emit_to_host(b, g, s);
        if (st != SHF_OK) return st;
    } else {
        SHF_CUDA(b->mark(5, s));
        st = launch_emit_split(f, b, g, s, (double)b->chunk_base[n_chunks] / ((double)n_chunks * W * H));
        if (st != SHF_OK) return st;
        SHF_CUDA(b->mark(6, s));
    }
    b->last_g = g;
    b->last_valid = true;
    st = publish(b, g);
    b->on_host = to_host;
    return st;
}

// (the calling entry point holds a DeviceGuard on f->device)
int bind_device(const shf_filter* f, shf_buffer* b) {
    if (b->device != f->device) {
        if (b->device >= 0) {
            cudaSetDevice(b->device);
            b->release_all();
            cudaSetDevice(f->device);
        }
        b->device = f->device;
    }
    if (!b->stream) SHF_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    return SHF_OK;
}

// the result of the last run_on_device into the buffer's page-locked host memory (what readHistogram() hands out)
int result_to_host(shf_buffer* b, cudaStream_t s) {
    SHF_CUDA(b->h_bins.ensure(std::max<size_t>(b->n_bins, 1) * sizeof(shf_bin)));
    SHF_CUDA(b->h_hso.ensure(b->n_offsets * 4));
    if (b->n_bins) {
        SHF_CUDA(cudaMemcpyAsync(b->h_bins.p, b->bins.p, b->n_bins * sizeof(shf_bin), cudaMemcpyDeviceToHost, s));
        tls_d2h += b->n_bins * sizeof(shf_bin);
    }
    SHF_CUDA(cudaMemcpyAsync(b->h_hso.p, b->hso.p, b->n_offsets * 4, cudaMemcpyDeviceToHost, s));
    tls_d2h += b->n_offsets * 4;
    SHF_CUDA(cudaStreamSynchronize(s));
    b->on_host = true;
    return SHF_OK;
}

int run_host(shf_filter* f, const uint16_t* const* maps, uint32_t n_chunks, const uint32_t map_size[2],
             const uint32_t nn[2], const uint32_t total[2], shf_buffer* b, uint32_t radius, bool to_host = true) {
    if (!f || !b || !maps || !map_size || !nn || !total)
        return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    int st = validate(map_size, nn, radius);
    if (st != SHF_OK) return st;
    DeviceGuard guard(f->device);
    st = bind_device(f, b);
    if (st != SHF_OK) return st;
    const uint32_t W = map_size[0], H = map_size[1], r = radius;
    // (row pitch of the device copy: 16-byte rows with 32 samples of slack, so that the kernels can read it in place)
    const uint32_t PW = W + 2u * r, PH = H + 2u * r, P = (PW + 39u) & ~7u;
    const size_t sx = (size_t)W * (nn[0] / 2u), sy = (size_t)H * (nn[1] / 2u);
    const size_t S = total[0];
    if (S < sx + W + r) return fail(SHF_ERR_INVALID_ARGUMENT, "TotalMapSize.x >= start + W + r", "row stride too small");
    if (total[1] < sy + H + r)
        return fail(SHF_ERR_INVALID_ARGUMENT, "TotalMapSize.y >= start + H + r", "the bottom halo lies outside the merged map");
    b->has_result = false;
    cudaStream_t s = b->stream;
    const size_t cells = (size_t)PH * P;
    SHF_CUDA(b->din.ensure((size_t)n_chunks * cells * 2));
    for (uint32_t i = 0; i < n_chunks; i++) {
        if (!maps[i]) return fail(SHF_ERR_INVALID_ARGUMENT, "samplemap != NULL", "null sample map");
        const uint16_t* src = maps[i] + (sy - r) * S + (sx - r);
        SHF_CUDA(cudaMemcpy2DAsync(b->din.as<uint16_t>() + (size_t)i * cells, (size_t)P * 2, src, S * 2, (size_t)PW * 2, PH,
                                   cudaMemcpyHostToDevice, s));
        tls_h2d += (size_t)PW * 2 * PH;
    }
    return run_on_device(f, b, b->din.as<uint16_t>(), cells, P, n_chunks, W, H, r, s, false, to_host);
}

// Several concurrent operator() calls of the world pipeline served by one pass: the
// neighbourhoods go through the kernels as one batch on buffers[0], then every caller's buffer receives its own chunk's
// bins and offsets in its own page-locked memory, exactly as if shf_run had been called on it.
int run_multi(shf_filter* f, const uint16_t* const* maps, shf_buffer* const* buffers, uint32_t n_calls,
              const uint32_t map_size[2], const uint32_t nn[2], const uint32_t total[2], uint32_t radius) {
    if (!f || !maps || !buffers || !map_size || !nn || !total)
        return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (n_calls == 0u) return fail(SHF_ERR_INVALID_ARGUMENT, "n_calls > 0", "empty batch");
    for (uint32_t i = 0; i < n_calls; i++) {
        if (!buffers[i]) return fail(SHF_ERR_INVALID_ARGUMENT, "buffer != NULL", "null filter buffer");
        for (uint32_t j = 0; j < i; j++)
            if (buffers[i] == buffers[j])
                return fail(SHF_ERR_INVALID_ARGUMENT, "buffers are distinct", "one filter buffer serves one call at a time");
    }
    shf_buffer* lead = buffers[0];
    int st = run_host(f, maps, n_calls, map_size, nn, total, lead, radius, false);
    if (st != SHF_OK) return st;
    DeviceGuard guard(f->device);
    cudaStream_t s = lead->stream;
    const size_t per = (size_t)map_size[0] * map_size[1] + 1u;
    const std::vector<uint64_t> base = lead->chunk_base;
    for (uint32_t i = 0; i < n_calls; i++) {
        shf_buffer* b = buffers[i];
        if (i) {
            st = bind_device(f, b);
            if (st != SHF_OK) return st;
            b->has_result = false;
        }
        const size_t nb = (size_t)(base[i + 1u] - base[i]);
        SHF_CUDA(b->h_bins.ensure(std::max<size_t>(nb, 1) * sizeof(shf_bin)));
        SHF_CUDA(b->h_hso.ensure(per * 4));
        if (nb) {
            SHF_CUDA(cudaMemcpyAsync(b->h_bins.p, lead->bins.as<shf_bin>() + base[i], nb * sizeof(shf_bin), cudaMemcpyDeviceToHost, s));
            tls_d2h += nb * sizeof(shf_bin);
        }
        SHF_CUDA(cudaMemcpyAsync(b->h_hso.p, lead->hso.as<uint32_t>() + (size_t)i * per, per * 4, cudaMemcpyDeviceToHost, s));
        tls_d2h += per * 4;
    }
    SHF_CUDA(cudaStreamSynchronize(s));
    for (uint32_t i = 0; i < n_calls; i++) {
        shf_buffer* b = buffers[i];
        b->n_chunks = 1u;
        b->last_w = map_size[0];
        b->last_h = map_size[1];
        b->n_bins = (size_t)(base[i + 1u] - base[i]);
        b->n_offsets = per;
        b->chunk_base.assign({0ull, (uint64_t)b->n_bins});
        b->has_result = true;
        b->on_host = true;
        b->dev_valid = i == 0u;  // chunk 0 sits at the start of the leader's device arrays
        b->ev_valid = b->ev_valid && i == 0u;
    }
    return SHF_OK;
}

// Neighbour merge on the device side of the bus. The reference packs the nn.x * nn.y chunk
// maps into one page-locked host buffer first (STPNearestNeighbourTextureBuffer.cpp:70-113: cudaMallocHost + 9 host to
// host 2D copies + sync) and the filter then walks that buffer. Here only the cells the filter reads -- the centre chunk
// plus a halo of `radius` cells -- are copied, straight from every neighbour's own map into the halo-extended device
// input: no merged host buffer, no page-locked allocation, (W+2r)(H+2r) instead of nn.x*nn.y*W*H cells moved.
// chunk_maps: n_chunks * nn.x * nn.y pointers (host or device memory, cudaMemcpyDefault), neighbour i of a neighbourhood
// at local coordinate (i % nn.x, i / nn.x) as in STPChunk::calcLocalChunkCoordinate, each map W x H row-major.
int gather_neighbours(shf_filter* f, shf_buffer* b, const uint16_t* const* chunk_maps, uint32_t n_chunks,
                      const uint32_t map_size[2], const uint32_t nn[2], uint32_t radius, cudaStream_t s) {
    const uint32_t W = map_size[0], H = map_size[1], r = radius;
    const uint32_t PW = W + 2u * r, PH = H + 2u * r, P = (PW + 39u) & ~7u;
    const size_t x0 = (size_t)W * (nn[0] / 2u) - r, y0 = (size_t)H * (nn[1] / 2u) - r;  // halo origin in the merged map
    const size_t cells = (size_t)PH * P;
    const uint32_t per = nn[0] * nn[1];
    SHF_CUDA(b->din.ensure((size_t)n_chunks * cells * 2));
    for (uint32_t i = 0; i < n_chunks; i++) {
        uint16_t* dst0 = b->din.as<uint16_t>() + (size_t)i * cells;
        for (uint32_t k = 0; k < per; k++) {
            const size_t cx = (size_t)(k % nn[0]) * W, cy = (size_t)(k / nn[0]) * H;  // the neighbour's origin
            const size_t xa = std::max(cx, x0), xb = std::min(cx + W, x0 + PW);
            const size_t ya = std::max(cy, y0), yb = std::min(cy + H, y0 + PH);
            if (xa >= xb || ya >= yb) continue;  // this neighbour lies outside the halo
            const uint16_t* map = chunk_maps[(size_t)i * per + k];
            if (!map) return fail(SHF_ERR_INVALID_ARGUMENT, "neighbour map != NULL", "null neighbour sample map");
            SHF_CUDA(cudaMemcpy2DAsync(dst0 + (ya - y0) * P + (xa - x0), (size_t)P * 2, map + (ya - cy) * W + (xa - cx),
                                       (size_t)W * 2, (xb - xa) * 2, yb - ya, cudaMemcpyDefault, s));
            tls_h2d += (xb - xa) * 2 * (yb - ya);
        }
    }
    (void)f;
    return SHF_OK;
}

int run_neighbours(shf_filter* f, const uint16_t* const* chunk_maps, uint32_t n_chunks, const uint32_t map_size[2],
                   const uint32_t nn[2], shf_buffer* b, uint32_t radius, cudaStream_t user_stream, bool to_host) {
    if (!f || !b || !chunk_maps || !map_size || !nn) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (n_chunks == 0u) return fail(SHF_ERR_INVALID_ARGUMENT, "n_chunks > 0", "empty batch");
    int st = validate(map_size, nn, radius);
    if (st != SHF_OK) return st;
    // the right / bottom halo must lie inside the neighbourhood as well (validate() only looks at the left / top margin;
    // an even neighbour count or a large radius would leave halo cells without a source map)
    if ((uint64_t)map_size[0] * (nn[0] / 2u) + map_size[0] + radius > (uint64_t)nn[0] * map_size[0] ||
        (uint64_t)map_size[1] * (nn[1] / 2u) + map_size[1] + radius > (uint64_t)nn[1] * map_size[1])
        return fail(SHF_ERR_INVALID_ARGUMENT, "start + MapSize + radius <= NearestNeighbour * MapSize",
                    "the right / bottom halo lies outside the chunk neighbourhood");
    DeviceGuard guard(f->device);
    st = bind_device(f, b);
    if (st != SHF_OK) return st;
    b->has_result = false;
    cudaStream_t s = to_host ? b->stream : user_stream;
    st = gather_neighbours(f, b, chunk_maps, n_chunks, map_size, nn, radius, s);
    if (st != SHF_OK) return st;
    const uint32_t PW = map_size[0] + 2u * radius, PH = map_size[1] + 2u * radius, P = (PW + 39u) & ~7u;
    return run_on_device(f, b, b->din.as<uint16_t>(), (size_t)PH * P, P, n_chunks, map_size[0], map_size[1], radius, s, false,
                         to_host);
}

// the const accessors complete a call that is still waiting for its checks
int settle_const(const shf_buffer* b) {
    if (!b->deferred.active) return SHF_OK;
    DeviceGuard guard(b->device);
    return settle(const_cast<shf_buffer*>(b));
}

}  // namespace

extern "C" {

int shf_filter_create(shf_filter** out, int device) {
    if (!out) return fail(SHF_ERR_INVALID_ARGUMENT, "out != NULL", "null argument");
    *out = nullptr;
    int count = 0;
    SHF_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) return fail(SHF_ERR_CUDA, "cudaGetDeviceCount", "no CUDA device: this filter has no CPU path");
    if (device < 0) SHF_CUDA(cudaGetDevice(&device));
    if (device >= count) return fail(SHF_ERR_INVALID_ARGUMENT, "device < device count", "no such device");
    shf_filter* f = new (std::nothrow) shf_filter();
    if (!f) return fail(SHF_ERR_INVALID_ARGUMENT, "new shf_filter", "out of host memory");
    f->device = device;
    int v = 0;
    cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) {
        f->sm_count = v;
        e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    }
    if (e != cudaSuccess) {
        delete f;
        return fail(SHF_ERR_CUDA, "cudaDeviceGetAttribute", cudaGetErrorString(e));
    }
    f->smem_optin = (size_t)v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device) != cudaSuccess) v = (int)f->smem_optin;
    f->smem_per_sm = (size_t)v;
    auto env_u32 = [](const char* name) {
        const char* e = getenv(name);
        return e ? (uint32_t)std::max(1, atoi(e)) : 0u;
    };
    f->dbg.no_speculation = getenv("SHF_NO_SPECULATION") != nullptr;
    f->dbg.no_cseg = getenv("SHF_NO_CSEG") != nullptr;
    f->dbg.no_vseg = getenv("SHF_NO_VSEG") != nullptr;
    f->dbg.persist = env_u32("SHF_PERSIST");
    f->dbg.no_split = getenv("SHF_NO_SPLIT") != nullptr;
    f->dbg.no_small_ids = getenv("SHF_NO_SMALL_IDS") != nullptr;
    f->dbg.debug_ty = env_u32("SHF_DEBUG_TY");
    f->dbg.debug_tail = env_u32("SHF_DEBUG_TAIL");
    f->dbg.debug_tailseg = env_u32("SHF_DEBUG_TAILSEG");
    f->dbg.debug_np = env_u32("SHF_DEBUG_NP");         // producer warps of an emit CTA
    f->dbg.debug_extra = env_u32("SHF_DEBUG_EXTRA");   // ring batches beyond those in production
    f->dbg.debug_cseg = env_u32("SHF_DEBUG_CSEG");
    f->dbg.debug_vseg = env_u32("SHF_DEBUG_VSEG");
    f->dbg.debug_persist = env_u32("SHF_DEBUG_PERSIST");
    *out = f;
    return SHF_OK;
}

void shf_filter_destroy(shf_filter* filter) { delete filter; }

int shf_buffer_create(shf_buffer** out, unsigned char execution_type) {
    if (!out) return fail(SHF_ERR_INVALID_ARGUMENT, "out != NULL", "null argument");
    *out = nullptr;
    if (execution_type != SHF_EXEC_SERIAL && execution_type != SHF_EXEC_PARALLEL)
        return fail(SHF_ERR_INVALID_ENUM, "STPExecutionType", "value is not a valid STPFilterBuffer::STPExecutionType");
    shf_buffer* b = new (std::nothrow) shf_buffer();
    if (!b) return fail(SHF_ERR_INVALID_ARGUMENT, "new shf_buffer", "out of host memory");
    b->exec_type = execution_type;
    *out = b;
    return SHF_OK;
}

void shf_buffer_destroy(shf_buffer* buffer) {
    if (!buffer) return;
    if (buffer->device >= 0) {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(buffer->device);
        buffer->release_all();
        if (prev >= 0) cudaSetDevice(prev);
    }
    delete buffer;
}

int shf_buffer_read(const shf_buffer* b, const shf_bin** bins, const uint32_t** offsets) {
    if (!b || !bins || !offsets) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (const int st = settle_const(b)) return st;
    if (!b->has_result || !b->on_host) {
        *bins = nullptr;
        *offsets = nullptr;
        return SHF_OK;
    }
    *bins = b->h_bins.as<shf_bin>();
    *offsets = b->h_hso.as<uint32_t>();
    return SHF_OK;
}

int shf_buffer_size(const shf_buffer* b, size_t* n_bins, size_t* n_offsets) {
    if (!b || !n_bins || !n_offsets) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (const int st = settle_const(b)) return st;
    *n_bins = b->has_result ? b->n_bins : 0;
    *n_offsets = b->has_result ? b->n_offsets : 0;
    return SHF_OK;
}

unsigned char shf_buffer_type(const shf_buffer* b) { return b ? b->exec_type : 0; }

int shf_run(shf_filter* filter, const uint16_t* samplemap, const uint32_t map_size[2], const uint32_t nn[2],
            const uint32_t total[2], shf_buffer* buffer, uint32_t radius) {
    const uint16_t* maps[1] = {samplemap};
    return run_host(filter, maps, 1u, map_size, nn, total, buffer, radius);
}

int shf_run_batch(shf_filter* filter, const uint16_t* const* samplemaps, uint32_t n_chunks, const uint32_t map_size[2],
                  const uint32_t nn[2], const uint32_t total[2], shf_buffer* buffer, uint32_t radius) {
    if (n_chunks == 0u) return fail(SHF_ERR_INVALID_ARGUMENT, "n_chunks > 0", "empty batch");
    return run_host(filter, samplemaps, n_chunks, map_size, nn, total, buffer, radius);
}

static int run_device_entry(shf_filter* f, const uint16_t* maps_dev, uint64_t chunk_stride, uint32_t n_chunks,
                            const uint32_t map_size[2], const uint32_t nn[2], const uint32_t total[2], shf_buffer* b,
                            uint32_t radius, void* stream, bool deferred) {
    if (!f || !b || !maps_dev || !map_size || !nn || !total)
        return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    int st = validate(map_size, nn, radius);
    if (st != SHF_OK) return st;
    DeviceGuard guard(f->device);
    st = bind_device(f, b);
    if (st != SHF_OK) return st;
    const uint32_t W = map_size[0], H = map_size[1];
    const size_t sx = (size_t)W * (nn[0] / 2u), sy = (size_t)H * (nn[1] / 2u);
    const size_t S = total[0];
    if (S < sx + W + radius) return fail(SHF_ERR_INVALID_ARGUMENT, "TotalMapSize.x >= start + W + r", "row stride too small");
    if (total[1] < sy + H + radius)
        return fail(SHF_ERR_INVALID_ARGUMENT, "TotalMapSize.y >= start + H + r", "the bottom halo lies outside the merged map");
    b->has_result = false;
    const uint16_t* view = maps_dev + (sy - radius) * S + (sx - radius);
    return run_on_device(f, b, view, chunk_stride, (uint32_t)S, n_chunks, W, H, radius, static_cast<cudaStream_t>(stream),
                         deferred);
}

int shf_run_device(shf_filter* f, const uint16_t* maps_dev, uint64_t chunk_stride, uint32_t n_chunks,
                   const uint32_t map_size[2], const uint32_t nn[2], const uint32_t total[2], shf_buffer* b,
                   uint32_t radius, void* stream) {
    return run_device_entry(f, maps_dev, chunk_stride, n_chunks, map_size, nn, total, b, radius, stream, false);
}

int shf_run_device_async(shf_filter* f, const uint16_t* maps_dev, uint64_t chunk_stride, uint32_t n_chunks,
                         const uint32_t map_size[2], const uint32_t nn[2], const uint32_t total[2], shf_buffer* b,
                         uint32_t radius, void* stream) {
    return run_device_entry(f, maps_dev, chunk_stride, n_chunks, map_size, nn, total, b, radius, stream, true);
}

int shf_buffer_wait(shf_buffer* b, uint64_t* repeated_calls) {
    if (!b) return fail(SHF_ERR_INVALID_ARGUMENT, "buffer != NULL", "null argument");
    const int st = settle_const(b);
    if (repeated_calls) *repeated_calls = b->spec_misses;
    return st;
}

int shf_run_multi(shf_filter* filter, const uint16_t* const* samplemaps, shf_buffer* const* buffers, uint32_t n_calls,
                  const uint32_t map_size[2], const uint32_t nn[2], const uint32_t total[2], uint32_t radius) {
    return run_multi(filter, samplemaps, buffers, n_calls, map_size, nn, total, radius);
}

int shf_run_neighbours(shf_filter* filter, const uint16_t* const* neighbour_maps, uint32_t n_chunks,
                       const uint32_t map_size[2], const uint32_t nn[2], shf_buffer* buffer, uint32_t radius) {
    return run_neighbours(filter, neighbour_maps, n_chunks, map_size, nn, buffer, radius, nullptr, true);
}

int shf_run_neighbours_device(shf_filter* filter, const uint16_t* const* neighbour_maps, uint32_t n_chunks,
                              const uint32_t map_size[2], const uint32_t nn[2], shf_buffer* buffer, uint32_t radius,
                              void* stream) {
    return run_neighbours(filter, neighbour_maps, n_chunks, map_size, nn, buffer, radius, static_cast<cudaStream_t>(stream), false);
}

int shf_buffer_read_device(const shf_buffer* b, const shf_bin** bins_dev, const uint32_t** offsets_dev) {
    if (!b || !bins_dev || !offsets_dev) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (const int st = settle_const(b)) return st;
    *bins_dev = b->has_result && b->dev_valid ? b->bins.as<shf_bin>() : nullptr;
    *offsets_dev = b->has_result && b->dev_valid ? b->hso.as<uint32_t>() : nullptr;
    return SHF_OK;
}

int shf_buffer_chunk_base(const shf_buffer* b, const uint64_t** chunk_base, uint32_t* n_chunks) {
    if (!b || !chunk_base || !n_chunks) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (const int st = settle_const(b)) return st;
    *chunk_base = b->has_result ? b->chunk_base.data() : nullptr;
    *n_chunks = b->has_result ? b->n_chunks : 0u;
    return SHF_OK;
}

int shf_heightfield_create(shf_heightfield** out, shf_filter* f, const shf_biome_property* table, uint32_t n_table,
                           const unsigned char* permutation, const float* gradient2d, uint32_t gradient2d_size) {
    if (!out || !f || !table || !permutation || !gradient2d)
        return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    *out = nullptr;
    if (n_table == 0u || gradient2d_size == 0u)
        return fail(SHF_ERR_INVALID_ARGUMENT, "n_table > 0 && gradient2d_size > 0", "empty table");
    const size_t smem = 1024 + (size_t)gradient2d_size * 8 + (size_t)n_table * sizeof(shf_biome_property);
    if (smem > f->smem_optin)
        return fail(SHF_ERR_UNSUPPORTED, "tables fit shared memory", "biome / gradient tables too large for one CTA");
    DeviceGuard guard(f->device);
    shf_heightfield* h = new (std::nothrow) shf_heightfield();
    if (!h) return fail(SHF_ERR_INVALID_ARGUMENT, "new shf_heightfield", "out of host memory");
    h->device = f->device;
    h->sm_count = f->sm_count;
    h->smem_optin = f->smem_optin;
    h->n_table = n_table;
    h->grad_size = gradient2d_size;
    cudaError_t e = h->table.ensure((size_t)n_table * sizeof(shf_biome_property));
    if (e == cudaSuccess) e = h->perm.ensure(512);
    if (e == cudaSuccess) e = h->grad.ensure((size_t)gradient2d_size * 8);
    if (e == cudaSuccess) e = cudaMemcpy(h->table.p, table, (size_t)n_table * sizeof(shf_biome_property), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->perm.p, permutation, 512, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->grad.p, gradient2d, (size_t)gradient2d_size * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        shf_heightfield_destroy(h);
        return fail(SHF_ERR_CUDA, "upload of the generator tables", cudaGetErrorString(e));
    }
    tls_h2d += (size_t)n_table * sizeof(shf_biome_property) + 512 + (size_t)gradient2d_size * 8;
    *out = h;
    return SHF_OK;
}

void shf_heightfield_destroy(shf_heightfield* h) {
    if (!h) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    h->table.release();
    h->perm.release();
    h->grad.release();
    if (prev >= 0) cudaSetDevice(prev);
    delete h;
}

int shf_heightfield_run(shf_heightfield* h, shf_buffer* b, uint32_t first_chunk, uint32_t n_chunks, const float* offsets_xy,
                        float* height_dev, void* stream) {
    if (!h || !b || !offsets_xy || !height_dev) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (const int st = settle_const(b)) return st;
    if (!b->has_result || !b->dev_valid)
        return fail(SHF_ERR_INVALID_ARGUMENT, "buffer holds a device result", "run the filter on this buffer first");
    if (b->device != h->device) return fail(SHF_ERR_INVALID_ARGUMENT, "same device", "histogram and generator live on different devices");
    if (n_chunks == 0u || first_chunk > b->n_chunks || n_chunks > b->n_chunks - first_chunk || n_chunks > 65535u)
        return fail(SHF_ERR_INVALID_ARGUMENT, "chunk range inside the last result", "no such chunks");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SHF_CUDA(b->hf_offsets.ensure((size_t)n_chunks * 8));
    SHF_CUDA(cudaMemcpyAsync(b->hf_offsets.p, offsets_xy, (size_t)n_chunks * 8, cudaMemcpyHostToDevice, s));
    tls_h2d += (size_t)n_chunks * 8;
    shf::HeightGeo g{};
    g.W = b->last_w;
    g.H = b->last_h;
    g.n_table = h->n_table;
    g.grad_size = h->grad_size;
    g.half_x = (float)g.W / 2.0f;
    g.half_y = (float)g.H / 2.0f;
    const size_t smem = 1024 + (size_t)h->grad_size * 8 + (size_t)h->n_table * sizeof(shf_biome_property);
    SHF_CUDA(cudaFuncSetAttribute(shf::heightfield_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t npx = g.W * g.H;
    const uint32_t gx = std::max(1u, std::min((npx + 255u) / 256u, (uint32_t)h->sm_count * 8u));
    shf::heightfield_kernel<<<dim3(gx, n_chunks), 256, smem, s>>>(
        g, b->bins.as<uint2>(), b->hso.as<uint32_t>(), b->chunkbase.as<uint64_t>(), first_chunk,
        h->table.as<shf_biome_property>(), h->perm.as<unsigned char>(), h->grad.as<float>(), b->hf_offsets.as<float2>(),
        height_dev);
    tls_launches++;
    SHF_CUDA(cudaGetLastError());
    return SHF_OK;
}

// STPLayer::mixSeed / seedLayer (STPLayer.cpp:143-151,178-182) on the host: one seed per layer
static uint64_t biome_mix_host(uint64_t s, uint64_t fac) {
    s *= s * 6364136223846793005ull + 1442695040888963407ull;
    return s + fac;
}

int shf_biome_factory_create(shf_biome_factory** out, shf_filter* f, uint32_t width, uint32_t height,
                             const shf_biome_layer* layers, uint32_t n_layers, uint64_t global_seed, uint64_t voronoi_seed,
                             const shf_biome_ids* ids) {
    if (!out || !f || !layers || !ids) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    *out = nullptr;
    // STPBiomeFactory.cpp:20-22
    if (width == 0u || height == 0u)
        return fail(SHF_ERR_NUMERIC_DOMAIN, "dimension.x > 0u && dimension.y > 0u",
                    "Biomemap should have strictly positive dimension in both vector components");
    if (n_layers == 0u || n_layers > 1024u) return fail(SHF_ERR_INVALID_ARGUMENT, "0 < n_layers <= 1024", "bad layer count");
    for (uint32_t i = 0; i < n_layers; i++) {
        if (layers[i].kind > SHF_LAYER_VORONOI)
            return fail(SHF_ERR_INVALID_ENUM, "shf_biome_layer_kind", "value is not a layer kind this producer implements");
        if (layers[i].kind != SHF_LAYER_CONTINENT && layers[i].parent >= i)
            return fail(SHF_ERR_INVALID_ARGUMENT, "layers[i].parent < i", "an ascendant must precede its descendants");
    }
    shf_biome_factory* bf = new (std::nothrow) shf_biome_factory();
    if (!bf) return fail(SHF_ERR_INVALID_ARGUMENT, "new shf_biome_factory", "out of host memory");
    bf->device = f->device;
    bf->sm_count = f->sm_count;
    bf->width = width;
    bf->height = height;
    bf->layers.assign(layers, layers + n_layers);
    bf->seeds.resize(n_layers);
    for (uint32_t i = 0; i < n_layers; i++) {
        const uint64_t salt = layers[i].salt;
        uint64_t mid = biome_mix_host(salt, salt);
        mid = biome_mix_host(mid, mid);
        mid = biome_mix_host(mid, mid);
        uint64_t s = biome_mix_host(global_seed, mid);
        s = biome_mix_host(s, mid);
        bf->seeds[i] = biome_mix_host(s, mid);
    }
    bf->voronoi_seed = voronoi_seed;
    bf->ids = shf::BiomeIds{ids->ocean, ids->plains, ids->forest, ids->frozen_ocean, ids->warm_ocean, ids->lukewarm_ocean,
                            ids->cold_ocean};
    bf->grid.resize(n_layers);
    *out = bf;
    return SHF_OK;
}

void shf_biome_factory_destroy(shf_biome_factory* bf) {
    if (!bf) return;
    {
        DeviceGuard guard(bf->device);
        for (DevBuf& g : bf->grid) g.release();
        bf->rects.release();
        bf->jitter.release();
        bf->h_rects.release();
        if (bf->rects_free) cudaEventDestroy(bf->rects_free);
    }
    delete bf;
}

int shf_biome_factory_run(shf_biome_factory* bf, uint16_t* biomemap_dev, uint32_t row_stride, uint64_t map_stride,
                          uint32_t n_maps, const int32_t* offsets_xz, void* stream) {
    if (!bf || !biomemap_dev || !offsets_xz) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (n_maps == 0u || n_maps > 65535u) return fail(SHF_ERR_INVALID_ARGUMENT, "0 < n_maps <= 65535", "bad map count");
    if (row_stride == 0u) row_stride = bf->width;
    if (row_stride < bf->width) return fail(SHF_ERR_INVALID_ARGUMENT, "row_stride >= BiomeDimension.x", "row stride too small");
    DeviceGuard guard(bf->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t L = (uint32_t)bf->layers.size();
    // ---- the rectangle every layer must supply, per map, from the root down ----
    if (!bf->rects_free) SHF_CUDA(cudaEventCreateWithFlags(&bf->rects_free, cudaEventDisableTiming));
    else SHF_CUDA(cudaEventSynchronize(bf->rects_free));
    const size_t n_rects = (size_t)L * n_maps;
    SHF_CUDA(bf->h_rects.ensure(n_rects * sizeof(shf::BiomeRect)));
    SHF_CUDA(bf->rects.ensure(n_rects * sizeof(shf::BiomeRect)));
    shf::BiomeRect* R = bf->h_rects.as<shf::BiomeRect>();
    std::vector<uint64_t> max_cells(L, 0ull);
    for (uint32_t m = 0; m < n_maps; m++) {
        struct Box { int64_t x0, z0, x1, z1; bool used; };   // inclusive
        std::vector<Box> box(L, Box{0, 0, 0, 0, false});
        box[L - 1u] = Box{offsets_xz[2 * m], offsets_xz[2 * m + 1], (int64_t)offsets_xz[2 * m] + bf->width - 1,
                          (int64_t)offsets_xz[2 * m + 1] + bf->height - 1, true};
        for (uint32_t i = L; i-- > 0u;) {
            const Box& b = box[i];
            if (!b.used || bf->layers[i].kind == SHF_LAYER_CONTINENT) continue;
            Box need{};
            switch (bf->layers[i].kind) {
                case SHF_LAYER_SCALE_NORMAL:
                case SHF_LAYER_SCALE_FUZZY: need = Box{b.x0 >> 1, b.z0 >> 1, (b.x1 + 1) >> 1, (b.z1 + 1) >> 1, true}; break;
                case SHF_LAYER_VORONOI: need = Box{(b.x0 - 2) >> 2, (b.z0 - 2) >> 2, ((b.x1 - 2) >> 2) + 1, ((b.z1 - 2) >> 2) + 1, true}; break;
                default: need = Box{b.x0 - 1, b.z0 - 1, b.x1 + 1, b.z1 + 1, true}; break;   // cross / X-cross layers
            }
            Box& p = box[bf->layers[i].parent];
            if (!p.used) p = need;
            else p = Box{std::min(p.x0, need.x0), std::min(p.z0, need.z0), std::max(p.x1, need.x1), std::max(p.z1, need.z1), true};
        }
        for (uint32_t i = 0; i < L; i++) {
            const Box& b = box[i];
            if (b.used && (b.x0 < INT32_MIN || b.x1 > INT32_MAX || b.z0 < INT32_MIN || b.z1 > INT32_MAX))
                return fail(SHF_ERR_INVALID_ARGUMENT, "coordinates fit int", "offset + dimension overflows the layer coordinates");
            shf::BiomeRect r{0, 0, 0u, 0u};
            if (b.used) r = shf::BiomeRect{(int32_t)b.x0, (int32_t)b.z0, (uint32_t)(b.x1 - b.x0 + 1), (uint32_t)(b.z1 - b.z0 + 1)};
            R[(size_t)i * n_maps + m] = r;
            max_cells[i] = std::max<uint64_t>(max_cells[i], (uint64_t)r.w * r.h);
        }
    }
    SHF_CUDA(cudaMemcpyAsync(bf->rects.p, R, n_rects * sizeof(shf::BiomeRect), cudaMemcpyHostToDevice, s));
    SHF_CUDA(cudaEventRecord(bf->rects_free, s));
    tls_h2d += n_rects * sizeof(shf::BiomeRect);
    // ---- the grids, leaf to root: one launch per layer over all maps ----
    uint64_t jit_cells = 0;
    for (uint32_t i = 0; i + 1u < L; i++) {
        if (max_cells[i] == 0ull) continue;
        if (max_cells[i] > 0xFFFFFFFFull) return fail(SHF_ERR_UNSUPPORTED, "cells per layer grid < 2^32", "map too large");
        SHF_CUDA(bf->grid[i].ensure((size_t)max_cells[i] * n_maps * 2));
    }
    for (uint32_t i = 0; i < L; i++)
        if (max_cells[i] && bf->layers[i].kind == SHF_LAYER_VORONOI) jit_cells = std::max(jit_cells, max_cells[bf->layers[i].parent]);
    if (jit_cells) SHF_CUDA(bf->jitter.ensure((size_t)jit_cells * 2u * 3u * 8u * n_maps));
    const shf::BiomeRect* rd = bf->rects.as<shf::BiomeRect>();
    for (uint32_t i = 0; i < L; i++) {
        if (max_cells[i] == 0ull) continue;   // not reachable from the root
        const shf_biome_layer& ly = bf->layers[i];
        const bool root = i + 1u == L;
        shf::BiomeLaunch p{};
        p.kind = ly.kind;
        p.seed = bf->seeds[i];
        p.voronoi_seed = bf->voronoi_seed;
        p.ids = bf->ids;
        p.n_maps = n_maps;
        p.out_map_stride = root ? map_stride : max_cells[i];
        p.out_row_stride = root ? row_stride : 0u;
        const uint32_t par = ly.kind == SHF_LAYER_CONTINENT ? i : ly.parent;
        p.in_map_stride = max_cells[par];
        const uint64_t jit_stride = max_cells[par] * 2u * 3u;
        if (ly.kind == SHF_LAYER_VORONOI) {
            const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((max_cells[par] * 2u + 255u) / 256u, (uint64_t)bf->sm_count * 8u));
            shf::voronoi_jitter_kernel<<<dim3(gx, n_maps), 256, 0, s>>>(bf->voronoi_seed, n_maps, rd + (size_t)par * n_maps,
                                                                       jit_stride, bf->jitter.as<double>());
            tls_launches++;
        }
        const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((max_cells[i] + 255u) / 256u, (uint64_t)bf->sm_count * 8u));
        shf::biome_layer_kernel<<<dim3(gx, n_maps), 256, 0, s>>>(
            p, rd + (size_t)i * n_maps, rd + (size_t)par * n_maps, bf->grid[par].as<uint16_t>(), bf->jitter.as<double>(),
            jit_stride, root ? biomemap_dev : bf->grid[i].as<uint16_t>());
        tls_launches++;
        SHF_CUDA(cudaGetLastError());
    }
    return SHF_OK;
}

const char* shf_last_error(void) { return tls_error.c_str(); }

void shf_stats_reset(void) { tls_launches = tls_h2d = tls_d2h = 0; }

void shf_stats_get(uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (kernel_launches) *kernel_launches = tls_launches;
    if (h2d_bytes) *h2d_bytes = tls_h2d;
    if (d2h_bytes) *d2h_bytes = tls_d2h;
}

void shf_set_profiling(int enabled) { g_profiling = enabled ? 1 : 0; }

int shf_buffer_phase_history(const shf_buffer* b, uint32_t back, float* ms, uint32_t n) {
    if (!b || !ms) return fail(SHF_ERR_INVALID_ARGUMENT, "arguments != NULL", "null argument");
    if (const int st = settle_const(b)) return st;
    if (!b->ev_valid) return fail(SHF_ERR_INVALID_ARGUMENT, "profiling enabled", "no phase events recorded for the last call");
    if (back >= (uint32_t)kEvRing || back >= b->ev_calls)
        return fail(SHF_ERR_INVALID_ARGUMENT, "back < profiled calls kept", "no phase events kept that far back");
    const cudaEvent_t* set = b->ev[(b->ev_calls - 1u - back) % kEvRing];
    for (int i = 0; i <= kPhases; i++)
        if (!set[i]) return fail(SHF_ERR_INVALID_ARGUMENT, "all phases recorded", "that call did not run every phase");
    SHF_CUDA(cudaEventSynchronize(set[kPhases]));
    for (uint32_t i = 0; i < n && i < (uint32_t)kPhases; i++) SHF_CUDA(cudaEventElapsedTime(&ms[i], set[i], set[i + 1]));
    return SHF_OK;
}

int shf_buffer_phase_ms(const shf_buffer* b, float* ms, uint32_t n) { return shf_buffer_phase_history(b, 0u, ms, n); }

int shf_buffer_last_plan(const shf_buffer* b, uint32_t* k_sets, uint32_t* rows_per_cta, uint32_t* n_biomes,
                         uint32_t* smem_bytes) {
    if (!b) return fail(SHF_ERR_INVALID_ARGUMENT, "buffer != NULL", "null argument");
    if (const int st = settle_const(b)) return st;
    if (k_sets) *k_sets = b->plan_k;
    if (rows_per_cta) *rows_per_cta = b->plan_ty;
    if (n_biomes) *n_biomes = b->plan_biomes;
    if (smem_bytes) *smem_bytes = b->plan_smem;
    return SHF_OK;
}

}  // extern "C"
