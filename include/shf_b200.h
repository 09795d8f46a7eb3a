/*
 * shf_b200.h -- C ABI of the B200-native single histogram filter.
 *
 * This is the drop-in boundary: every entry point below is what a binding of SuperTerrain+'s
 * `SuperTerrainPlus::STPAlgorithm::STPSingleHistogramFilter` would call. The reference interface each one replaces is
 * cited as SHF.h / SHF.cpp / SH.hpp:
 *   SHF.h   = SuperTerrain+/SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPSingleHistogramFilter.h
 *   SHF.cpp = SuperTerrain+/SuperAlgorithm+/Host/Private/STPSingleHistogramFilter.cpp
 *   SH.hpp  = SuperTerrain+/SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPSingleHistogram.hpp
 * The C++ class with the reference's own names that sits on top of this ABI is in
 * include/SuperAlgorithm+Host/STPSingleHistogramFilter.h; INTEGRATION.md shows how a maintainer swaps it in.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary. All functions are thread-safe as long as
 * one shf_buffer is used by one call at a time (the reference's rule, SHF.h:33-34 and
 * SuperDemo+/World/Biomes/STPBiomefieldGenerator.cpp:102-124). There is no CPU fallback: without a CUDA device every
 * compute entry point returns SHF_ERR_CUDA.
 */
#ifndef SHF_B200_H
#define SHF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SHF_API __declspec(dllexport)
#else
#define SHF_API __attribute__((visibility("default")))
#endif

/* Layout-identical to STPSingleHistogram::STPBin (SH.hpp:23-37): {uint16 Item; 2 B padding; float Weight}, 8 bytes.
 * The padding bytes are written as zero by this implementation (the reference leaves them indeterminate). */
typedef struct shf_bin {
    uint16_t item;
    float weight;
} shf_bin;

typedef struct shf_filter shf_filter; /* <-> STPSingleHistogramFilter object (SHF.h:36) */
typedef struct shf_buffer shf_buffer; /* <-> STPSingleHistogramFilter::STPFilterBuffer (SHF.h:43-123) */

/* STPFilterBuffer::STPExecutionType (SHF.h:49-52). Both map to the same GPU path; the value is only echoed back. */
#define SHF_EXEC_SERIAL 0x00u
#define SHF_EXEC_PARALLEL 0xFFu

enum shf_status {
    SHF_OK = 0,
    SHF_ERR_NUMERIC_DOMAIN = 1,  /* <-> STPNumericDomainError, SHF.cpp:874,879-880 */
    SHF_ERR_INVALID_ENUM = 2,    /* <-> STPInvalidEnum, SHF.cpp:721 */
    SHF_ERR_CUDA = 3,            /* <-> STPCUDAError (STP_CHECK_CUDA) */
    SHF_ERR_OFFSET_OVERFLOW = 4, /* total bins of one chunk do not fit HistogramStartOffset's uint32 (SH.hpp:48) */
    SHF_ERR_UNSUPPORTED = 5,     /* shape outside what the kernels implement; message says which limit */
    SHF_ERR_INVALID_ARGUMENT = 6
};

/* ---- filter object: constructor / destructor of STPSingleHistogramFilter (SHF.h:160-170). `device` < 0 = the
 * calling thread's current CUDA device. The object holds no per-call state and may be shared by threads. */
SHF_API int shf_filter_create(shf_filter** out, int device);
SHF_API void shf_filter_destroy(shf_filter* filter);

/* ---- filter buffer: STPFilterBuffer(STPExecutionType) / ~STPFilterBuffer (SHF.h:96-106). A C handle is moved by
 * copying the pointer. A fresh buffer reads as {NULL, NULL} with size (0, 0) (SHF.cpp:735-750). */
SHF_API int shf_buffer_create(shf_buffer** out, unsigned char execution_type);
SHF_API void shf_buffer_destroy(shf_buffer* buffer);
/* STPFilterBuffer::readHistogram (SHF.h:108-114): pointers into PAGE-LOCKED HOST memory owned by the buffer, valid
 * until the buffer is reused or destroyed. */
SHF_API int shf_buffer_read(const shf_buffer* buffer, const shf_bin** bins, const uint32_t** offsets);
/* STPFilterBuffer::size (SHF.h:116-120): (number of bins, number of offsets = W*H+1 per chunk). */
SHF_API int shf_buffer_size(const shf_buffer* buffer, size_t* n_bins, size_t* n_offsets);
/* STPFilterBuffer::type (SHF.h:122-126). */
SHF_API unsigned char shf_buffer_type(const shf_buffer* buffer);

/* ---- STPSingleHistogramFilter::operator() (SHF.h:172-188, SHF.cpp:870-902).
 * samplemap: HOST pointer, row-major, row stride total_map_size[0], not retained after return.
 * map_size / nearest_neighbour / total_map_size: the three uvec2 of STPNearestNeighbourInformation.
 * Synchronous: on SHF_OK the histogram is complete in the buffer's page-locked host memory. */
SHF_API int shf_run(shf_filter* filter, const uint16_t* samplemap, const uint32_t map_size[2],
                    const uint32_t nearest_neighbour[2], const uint32_t total_map_size[2], shf_buffer* buffer,
                    uint32_t radius);

/* ---- additive entry points (not in the reference) ---- */

/* N independent chunk neighbourhoods of identical geometry in one call. samplemaps[i] is a HOST pointer like shf_run's.
 * Result in page-locked host memory: bins of all chunks concatenated; offsets as n_chunks blocks of (W*H+1) uint32,
 * each block relative to its own chunk's first bin; shf_buffer_chunk_base() gives every chunk's first-bin index. */
SHF_API int shf_run_batch(shf_filter* filter, const uint16_t* const* samplemaps, uint32_t n_chunks,
                          const uint32_t map_size[2], const uint32_t nearest_neighbour[2],
                          const uint32_t total_map_size[2], shf_buffer* buffer, uint32_t radius);

/* Device-resident variant: samplemaps_dev points at n_chunks merged maps in DEVICE memory, chunk i starting at
 * samplemaps_dev + i * chunk_stride (in samples). The result stays in device memory (shf_buffer_read_device); nothing is
 * copied to the host except per-chunk distinct-value counts and bin totals. `stream` is a cudaStream_t (NULL = the
 * legacy default stream). The call BLOCKS THE HOST: the first call of a shape waits for the stream twice (plan, bin
 * buffer size) and returns with the emitting kernel still in flight; a repeated call of the same shape on the same
 * buffer runs ahead with the previous plan and waits once, at its end, to check it. It may free and reallocate device
 * scratch, so it cannot be stream-captured. A later call on the same buffer from another stream waits (on the device)
 * for the work this one left in flight. */
SHF_API int shf_run_device(shf_filter* filter, const uint16_t* samplemaps_dev, uint64_t chunk_stride, uint32_t n_chunks,
                           const uint32_t map_size[2], const uint32_t nearest_neighbour[2],
                           const uint32_t total_map_size[2], shf_buffer* buffer, uint32_t radius, void* stream);

/* The same without any host synchronisation when the call can run ahead (same shape as the last COMPLETED call on this
 * buffer; otherwise it behaves like shf_run_device): everything is enqueued on `stream` and the call returns. The
 * checks of the plan it ran with are made by the first query of the result -- shf_buffer_wait, shf_buffer_size,
 * shf_buffer_read_device, shf_buffer_chunk_base, shf_heightfield_run -- which waits for the call; if the input needed
 * another plan or larger buffers, that query repeats the call on the checked path, so `samplemaps_dev` must stay
 * unchanged until then (or until the next call on the buffer, which simply replaces the pending one). Back-to-back
 * calls on one stream therefore keep the device busy without a bubble between them. */
SHF_API int shf_run_device_async(shf_filter* filter, const uint16_t* samplemaps_dev, uint64_t chunk_stride,
                                 uint32_t n_chunks, const uint32_t map_size[2], const uint32_t nearest_neighbour[2],
                                 const uint32_t total_map_size[2], shf_buffer* buffer, uint32_t radius, void* stream);
/* Completes a pending shf_run_device_async on the buffer and returns its status (SHF_OK at once when nothing is
 * pending). On return the result is complete in stream order (a repeated call may still have its emitting kernel in
 * flight on the call's stream). repeated_calls (optional): how many calls on this buffer ran ahead with a plan that did
 * not fit and were repeated so far -- work enqueued behind such a call read a discarded result and must be re-issued. */
SHF_API int shf_buffer_wait(shf_buffer* buffer, uint64_t* repeated_calls);

/* n_calls concurrent operator() calls served by ONE pass over the device (the world
 * pipeline runs up to five STPBiomefieldGenerator::operator() at once, SuperDemo+/World/Biomes/
 * STPBiomefieldGenerator.cpp:79-104, SuperTerrain+/Private/World/STPWorldPipeline.cpp:448-474, every one with its own
 * pooled STPFilterBuffer). samplemaps[i] / buffers[i] are the arguments call i would pass to shf_run; all calls share the
 * geometry and the radius. On SHF_OK every buffers[i] reads exactly as after shf_run(filter, samplemaps[i], ...,
 * buffers[i], radius): its own chunk's bins and offsets in its own page-locked memory, size() = (its bins, W*H+1).
 * Device scratch is taken from buffers[0]; only buffers[0] keeps a device-resident view (of its own chunk). The buffers
 * must be distinct. */
SHF_API int shf_run_multi(shf_filter* filter, const uint16_t* const* samplemaps, shf_buffer* const* buffers,
                          uint32_t n_calls, const uint32_t map_size[2], const uint32_t nearest_neighbour[2],
                          const uint32_t total_map_size[2], uint32_t radius);

/* Neighbour merge without the merged host buffer (replaces the pack step of
 * STPNearestNeighbourTextureBuffer::STPMergedBuffer, SuperTerrain+/Private/World/Chunk/
 * STPNearestNeighbourTextureBuffer.cpp:18-39,70-113, in front of the filter call at SuperDemo+/World/Biomes/
 * STPBiomefieldGenerator.cpp:94-104). neighbour_maps holds n_chunks * nn.x * nn.y pointers: neighbour i of a
 * neighbourhood sits at local coordinate (i % nn.x, i / nn.x) (STPChunk::calcLocalChunkCoordinate, STPChunk.cpp:71-73),
 * every map is MapSize.x * MapSize.y samples, row-major, in HOST or DEVICE memory. Only the centre chunk and a halo of
 * `radius` samples are copied, directly into the device input. Same result and errors as shf_run_batch on the merged
 * maps; result in page-locked host memory. */
SHF_API int shf_run_neighbours(shf_filter* filter, const uint16_t* const* neighbour_maps, uint32_t n_chunks,
                               const uint32_t map_size[2], const uint32_t nearest_neighbour[2], shf_buffer* buffer,
                               uint32_t radius);
/* Same, result left in device memory (shf_buffer_read_device); copies and kernels are enqueued on `stream`. */
SHF_API int shf_run_neighbours_device(shf_filter* filter, const uint16_t* const* neighbour_maps, uint32_t n_chunks,
                                      const uint32_t map_size[2], const uint32_t nearest_neighbour[2],
                                      shf_buffer* buffer, uint32_t radius, void* stream);

/* Device pointers of the last result (bins concatenated, offsets n_chunks x (W*H+1)). */
SHF_API int shf_buffer_read_device(const shf_buffer* buffer, const shf_bin** bins_dev, const uint32_t** offsets_dev);
/* Host array of n_chunks+1 first-bin indices (last = total bins) of the last result. */
SHF_API int shf_buffer_chunk_base(const shf_buffer* buffer, const uint64_t** chunk_base, uint32_t* n_chunks);

/* ---- device-side consumer of the histogram (additive) ----
 * The multi-biome heightfield of SuperDemo+/Script/STPMultiHeightGenerator.cu:35-71, computed from the DEVICE-RESIDENT
 * result of the last shf_run_device / shf_run_batch / shf_run on a buffer, so that bins and offsets never travel to the
 * host and back (the copy at SuperDemo+/World/Biomes/STPBiomefieldGenerator.cpp:108-123 disappears):
 *   height(x, y) = sum over the pixel's bins in bin order of Weight * (simplex2DFractal(x, y; biome) * Variation + Depth)
 * with simplex2DFractal as in SuperAlgorithm+/Device/Private/STPSimplexNoise.cu:84-109. */

/* Layout-identical to STPDemo::STPBiomeProperty (SuperDemo+/World/Biomes/STPBiomeProperty.hpp:10-27). */
typedef struct shf_biome_property {
    float scale;
    uint32_t octave;
    float persistence;
    float lacunarity;
    float depth;
    float variation;
} shf_biome_property;

typedef struct shf_heightfield shf_heightfield; /* tables of one generator, uploaded once */

/* table[i] = properties of sample value i (the reference's __constant__ BiomeTable, STPMultiHeightGenerator.cu:14);
 * bins whose item is >= n_table contribute nothing. permutation: 512 bytes (a permutation of 0..255, repeated) and
 * gradient2d: 2 * gradient2d_size floats, as produced by STPPermutationGenerator (SuperAlgorithm+/Host/Private/
 * STPPermutationGenerator.cpp:40-93) -- all three are HOST pointers, copied to the filter's device. */
SHF_API int shf_heightfield_create(shf_heightfield** out, shf_filter* filter, const shf_biome_property* table,
                                   uint32_t n_table, const unsigned char* permutation, const float* gradient2d,
                                   uint32_t gradient2d_size);
SHF_API void shf_heightfield_destroy(shf_heightfield* generator);
/* Heightfields of chunks [first_chunk, first_chunk + n_chunks) of the buffer's last result. offsets_xy: HOST array of
 * 2 * n_chunks floats (the `offset` argument of the reference kernel per chunk). height_dev: DEVICE pointer, n_chunks
 * blocks of MapSize.x * MapSize.y floats, row-major. `stream` is a cudaStream_t; returns once the work is enqueued. */
SHF_API int shf_heightfield_run(shf_heightfield* generator, shf_buffer* buffer, uint32_t first_chunk, uint32_t n_chunks,
                                const float* offsets_xy, float* height_dev, void* stream);

/* ---- biome-map producer on the device (additive) ----
 * What STPBiomeFactory::operator()(STPSample_t* biomemap, glm::ivec2 offset) computes on the CPU
 * (SuperTerrain+/SuperTerrain+/Private/World/Diversity/STPBiomeFactory.cpp:24-42: biomemap[x + z * dim.x] =
 * tree.retrieve(x + offset.x, 0, z + offset.y) through a chain of STPLayer objects, STPLayer.cpp:118-196), written
 * straight into DEVICE memory, so that the uint16 sample map the filter reads never exists on the host. The reference's
 * layers are virtual C++ classes supplied by the application (STPBiomeFactory::supply()); here the chain is DATA: a list
 * of the demo's layer kinds (SuperDemo+/World/Layers/) with their salts, in construction order -- every layer names its
 * ascendant by index, the last layer is the root the factory samples (STPAllLayers.cpp:61-109). */
enum shf_biome_layer_kind {
    SHF_LAYER_CONTINENT = 0,    /* STPContinentLayer.h:17-23, no ascendant */
    SHF_LAYER_SCALE_NORMAL = 1, /* STPScaleLayer.h:31-98, STPScaleType::NORMAL */
    SHF_LAYER_SCALE_FUZZY = 2,  /* STPScaleLayer.h:31-98, STPScaleType::FUZZY */
    SHF_LAYER_LAND = 3,         /* STPXCrossLayer.h:26-37 + STPLandLayer.h:20-70 */
    SHF_LAYER_ISLAND = 4,       /* STPCrossLayer.h:26-37 + STPIslandLayer.h:20-27 */
    SHF_LAYER_VORONOI = 5       /* STPVoronoiLayer.h:56-101 with Is3D = false */
};
typedef struct shf_biome_layer {
    uint32_t kind;   /* shf_biome_layer_kind */
    uint32_t parent; /* index of the ascendant layer (< this layer's index); ignored for SHF_LAYER_CONTINENT */
    uint64_t salt;   /* STPLayer's salt; the layer seed is STPLayer::seedLayer(global seed, salt) */
} shf_biome_layer;
/* Sample values the layers test and produce: STPBiomeRegistry::{Ocean, Plains, Forest, FrozenOcean, WarmOcean,
 * LukewarmOcean, ColdOcean}.ID (SuperDemo+/World/Biomes/STPBiomeRegistry.cpp:112-116; loaded from Biome.ini by the demo). */
typedef struct shf_biome_ids {
    uint16_t ocean, plains, forest, frozen_ocean, warm_ocean, lukewarm_ocean, cold_ocean;
} shf_biome_ids;
typedef struct shf_biome_factory shf_biome_factory; /* <-> STPBiomeFactory (STPBiomeFactory.h:21-73) */

/* STPBiomeFactory(dimension) + the chain supply() would build. voronoi_seed = std::hash<STPSeed_t>{}(global_seed)
 * (STPVoronoiLayer.h:52), which is standard-library defined: the identity with libstdc++. */
SHF_API int shf_biome_factory_create(shf_biome_factory** out, shf_filter* filter, uint32_t width, uint32_t height,
                                     const shf_biome_layer* layers, uint32_t n_layers, uint64_t global_seed,
                                     uint64_t voronoi_seed, const shf_biome_ids* ids);
SHF_API void shf_biome_factory_destroy(shf_biome_factory* factory);
/* operator()(biomemap, offset) for n_maps offsets at once. biomemap_dev: DEVICE memory, map i at biomemap_dev +
 * i * map_stride, row z of a map at + z * row_stride (0 = width), i.e. a map may be a window of a larger image -- e.g.
 * only the (W+2r) x (H+2r) cells the filter reads of a merged neighbourhood map. offsets_xz: HOST array of 2 * n_maps
 * ints, the world coordinate of every map's first cell. Work is enqueued on `stream` (cudaStream_t); the call returns
 * without waiting for it. One call at a time per factory. */
SHF_API int shf_biome_factory_run(shf_biome_factory* factory, uint16_t* biomemap_dev, uint32_t row_stride,
                                  uint64_t map_stride, uint32_t n_maps, const int32_t* offsets_xz, void* stream);

/* Message of the last failure on the calling thread: "<expression>: <description>" in the spirit of
 * STPException::STPBasic::what(), so that the C++ shim can throw the matching exception type. */
SHF_API const char* shf_last_error(void);

/* Counters for the benchmark harness: kernels launched / bytes copied by this library on the calling thread since the
 * last reset. */
SHF_API void shf_stats_reset(void);
SHF_API void shf_stats_get(uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* Kernel configuration chosen by the last call on this buffer (for reports): K = 32-biome register sets per row,
 * rows per CTA, distinct biomes found, shared memory per CTA. */
SHF_API int shf_buffer_last_plan(const shf_buffer* buffer, uint32_t* k_sets, uint32_t* rows_per_cta,
                                 uint32_t* n_biomes, uint32_t* smem_bytes);

/* Measurement hook: when enabled, every call records CUDA events on its stream between the kernel phases
 * (0 dictionary, 1 remap + vertical scan, 2 event lists (wide path: counting pass), 3 row scan, 4 host round trip for
 * the bin totals, 5 emit); shf_buffer_phase_ms waits for the last call on the buffer and returns the milliseconds of
 * the first n phases. */
SHF_API void shf_set_profiling(int enabled);
SHF_API int shf_buffer_phase_ms(const shf_buffer* buffer, float* ms, uint32_t n);
/* Same for the call `back` calls before the last one (0 = the last; the events of the last 64 profiled calls are kept),
 * so that a timed loop can be read back afterwards without a host synchronisation per call. */
SHF_API int shf_buffer_phase_history(const shf_buffer* buffer, uint32_t back, float* ms, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* SHF_B200_H */
