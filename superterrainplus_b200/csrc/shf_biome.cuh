// shf_biome.cuh -- the biome-map producer on the device: what
// STPBiomeFactory::operator() (SuperTerrain+/SuperTerrain+/Private/World/Diversity/STPBiomeFactory.cpp:24-42) computes
// pixel by pixel on the CPU -- `tree.retrieve(x + offset.x, 0, z + offset.y)` through a chain of STPLayer objects with
// hashed per-cell random numbers (STPLayer.cpp:118-196) -- evaluated here level by level as dense grids.
//
// A layer is a pure function of (x, z) and of its ascendant's values at a few neighbouring cells (the per-layer cache
// of the reference, STPLayer.cpp:77-92, only saves recomputation). So for a requested output rectangle the host works
// out, from the root down, the rectangle every ascendant must supply (scale layers halve it, Voronoi layers quarter it,
// cross layers grow it by one cell each way), and the kernels below then fill the rectangles from the leaf up, one
// thread per cell, one launch per layer for all maps of a batch. Nothing is cached, nothing recursive, every cell is
// computed exactly once; the u16 map is born in HBM where the filter reads it.
//
// Layer semantics follow the demo's layers (SuperDemo+/World/Layers/): STPContinentLayer.h:17-23, STPScaleLayer.h:31-98,
// STPXCrossLayer.h:26-37 + STPLandLayer.h:20-70, STPCrossLayer.h:26-37 + STPIslandLayer.h:20-27, STPVoronoiLayer.h:19-101
// (2-D: every Voronoi layer of the demo chain has Is3D = false). Integer work throughout, except the Voronoi layer's
// nearest-corner test, which compares sums of three squared doubles: every operation is written with an explicit
// round-to-nearest intrinsic (no fused multiply-add), i.e. evaluated as the C++ source is written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shf {

enum BiomeLayerKind : uint32_t {
    kLayerContinent = 0, kLayerScaleNormal = 1, kLayerScaleFuzzy = 2, kLayerLand = 3, kLayerIsland = 4, kLayerVoronoi = 5
};

struct BiomeIds {
    uint16_t ocean, plains, forest, frozen_ocean, warm_ocean, lukewarm_ocean, cold_ocean;
};

// the rectangle of one layer's grid one map needs: cells [x0, x0 + w) x [z0, z0 + h) in the layer's own coordinates
struct BiomeRect {
    int32_t x0, z0;
    uint32_t w, h;
};

struct BiomeLaunch {
    uint32_t kind;
    uint64_t seed;           // the layer's seed (STPLayer::seedLayer)
    uint64_t voronoi_seed;   // std::hash<STPSeed_t>{}(global seed), STPVoronoiLayer.h:52
    BiomeIds ids;
    uint32_t n_maps;
    uint64_t out_map_stride, in_map_stride;   // elements between the grids of consecutive maps
    uint32_t out_row_stride;                  // 0 = the rectangle's own width (scratch grids)
};

// STPLayer::mixSeed, STPLayer.cpp:178-182
__device__ __forceinline__ uint64_t biome_mix(uint64_t s, long long fac) {
    s *= s * 6364136223846793005ull + 1442695040888963407ull;
    return s + (uint64_t)fac;
}
// STPLayer::seedLocal, STPLayer.cpp:153-159
__device__ __forceinline__ uint64_t biome_seed_local(uint64_t layer_seed, int x, int z) {
    uint64_t s = biome_mix(layer_seed, x);
    s = biome_mix(s, z);
    s = biome_mix(s, x);
    return biome_mix(s, z);
}
// STPLocalSampler::nextValue, STPLayer.cpp:118-127
__device__ __forceinline__ uint32_t biome_next(uint64_t layer_seed, uint64_t& local, uint32_t range) {
    const uint32_t v = (uint32_t)((local >> 24) % (uint64_t)range);
    local = biome_mix(local, (long long)layer_seed);
    return v;
}
__device__ __forceinline__ bool biome_shallow(const BiomeIds& id, uint32_t v) {   // STPBiomeRegistry.cpp:112-116
    return v == id.ocean || v == id.frozen_ocean || v == id.warm_ocean || v == id.lukewarm_ocean || v == id.cold_ocean;
}

// the three jitters of one Voronoi lattice corner (STPVoronoiLayer.h:19-40): (d, e, f) for the x, y, z distance terms
__device__ __forceinline__ void voronoi_jitter(uint64_t seed, int x, int y, int z, double& d, double& e, double& f) {
    uint64_t m = biome_mix(seed, x);
    m = biome_mix(m, y);
    m = biome_mix(m, z);
    m = biome_mix(m, x);
    m = biome_mix(m, y);
    m = biome_mix(m, z);
    // (k / 1024.0 and the subtraction are exact; the product rounds once)
    d = __dmul_rn(__dadd_rn((double)(uint32_t)((m >> 24) & 1023ull) / 1024.0, -0.5), 0.9);
    m = biome_mix(m, (long long)seed);
    e = __dmul_rn(__dadd_rn((double)(uint32_t)((m >> 24) & 1023ull) / 1024.0, -0.5), 0.9);
    m = biome_mix(m, (long long)seed);
    f = __dmul_rn(__dadd_rn((double)(uint32_t)((m >> 24) & 1023ull) / 1024.0, -0.5), 0.9);
}

// Jitters of every lattice corner a Voronoi layer's rectangle touches: corner (X, y, Z) with y in {-1, 0} (the 2-D
// layer evaluates the 3-D cell at y = 0: (0 - 2) >> 2 = -1), X / Z over the ascendant's rectangle. Sixteen pixels share
// a cell's eight corners, so hashing them once per corner instead of once per pixel removes ~15/16 of the 64-bit work.
// jit[((m * 2 + yi) * ph + zi) * pw + xi] = (d, e, f)
__global__ void voronoi_jitter_kernel(uint64_t voronoi_seed, uint32_t n_maps, const BiomeRect* __restrict__ parent_rect,
                                      uint64_t jit_map_stride, double* __restrict__ jit) {
    const uint32_t m = blockIdx.y;
    const BiomeRect pr = parent_rect[m];
    const uint32_t cells = pr.w * pr.h * 2u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
        const uint32_t xi = i % pr.w, zi = (i / pr.w) % pr.h, yi = i / (pr.w * pr.h);
        double d, e, f;
        voronoi_jitter(voronoi_seed, pr.x0 + (int)xi, (int)yi - 1, pr.z0 + (int)zi, d, e, f);
        double* o = jit + (size_t)m * jit_map_stride + (size_t)i * 3u;
        o[0] = d;
        o[1] = e;
        o[2] = f;
    }
    (void)n_maps;
}

__global__ void __launch_bounds__(256) biome_layer_kernel(BiomeLaunch p, const BiomeRect* __restrict__ self_rect,
                                                          const BiomeRect* __restrict__ parent_rect,
                                                          const uint16_t* __restrict__ parent,
                                                          const double* __restrict__ jit, uint64_t jit_map_stride,
                                                          uint16_t* __restrict__ out) {
    const uint32_t m = blockIdx.y;
    const BiomeRect sr = self_rect[m];
    BiomeRect pr{0, 0, 1u, 1u};
    if (p.kind != kLayerContinent) pr = parent_rect[m];
    const uint16_t* par = parent + (size_t)m * p.in_map_stride;
    uint16_t* dst = out + (size_t)m * p.out_map_stride;
    const uint32_t row_stride = p.out_row_stride ? p.out_row_stride : sr.w;
    // ascendant's value at layer coordinate (x, z) (STPLayer::retrieve)
    auto up = [&](int x, int z) -> uint32_t { return par[(size_t)(uint32_t)(z - pr.z0) * pr.w + (uint32_t)(x - pr.x0)]; };
    const uint32_t cells = sr.w * sr.h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
        const uint32_t cx = i % sr.w, cz = i / sr.w;
        const int x = sr.x0 + (int)cx, z = sr.z0 + (int)cz;
        uint32_t v;
        switch (p.kind) {
        case kLayerContinent: {   // STPContinentLayer.h:17-23: one cell in ten is land
            uint64_t local = biome_seed_local(p.seed, x, z);
            v = biome_next(p.seed, local, 10u) == 0u ? p.ids.plains : p.ids.ocean;
            break;
        }
        case kLayerScaleNormal:
        case kLayerScaleFuzzy: {  // STPScaleLayer.h:31-98
            const uint32_t c = up(x >> 1, z >> 1);
            const int xb = x & 1, zb = z & 1;
            v = c;
            if (xb | zb) {
                uint64_t local = biome_seed_local(p.seed, x & -2, z & -2);
                const uint32_t s = up(x >> 1, (z + 1) >> 1);
                const uint32_t mm = biome_next(p.seed, local, 2u) == 0u ? c : s;
                if (xb == 0) {
                    v = mm;
                } else {
                    const uint32_t e = up((x + 1) >> 1, z >> 1);
                    const uint32_t o = biome_next(p.seed, local, 2u) == 0u ? c : e;
                    if (zb == 0) {
                        v = o;
                    } else {
                        const uint32_t se = up((x + 1) >> 1, (z + 1) >> 1);
                        const uint32_t pick = biome_next(p.seed, local, 4u);
                        const uint32_t ret = pick == 0u ? c : pick == 1u ? e : pick == 2u ? s : se;
                        if (p.kind == kLayerScaleFuzzy) v = ret;
                        else if (e == s && e == se) v = e;
                        else if (c == e && (c == se || s != se)) v = c;
                        else if (c == s && (c == se || e != se)) v = c;
                        else if (c == se && e != s) v = c;
                        else if (e == s && c != se) v = e;
                        else if (e == se && c != s) v = e;
                        else if (s == se && c != e) v = s;
                        else v = ret;
                    }
                }
            }
            break;
        }
        case kLayerLand: {        // STPXCrossLayer.h:26-37, STPLandLayer.h:20-70
            const uint32_t c = up(x, z), ne = up(x + 1, z - 1), se = up(x + 1, z + 1), sw = up(x - 1, z + 1), nw = up(x - 1, z - 1);
            uint64_t local = biome_seed_local(p.seed, x, z);
            const uint32_t forest = p.ids.forest;
            const bool oc = biome_shallow(p.ids, c), osw = biome_shallow(p.ids, sw), ose = biome_shallow(p.ids, se),
                       one = biome_shallow(p.ids, ne), onw = biome_shallow(p.ids, nw);
            if (!oc || (osw && ose && one && onw)) {
                if (oc || (!osw && !ose && !one && !onw) || biome_next(p.seed, local, 5u) != 0u) v = c;
                else if (onw) v = c == forest ? forest : nw;
                else if (osw) v = c == forest ? forest : sw;
                else if (one) v = c == forest ? forest : ne;
                else if (ose) v = c == forest ? forest : se;
                else v = c;
            } else {
                uint32_t range = 1u, j = 1u;   // land grows out of the neighbours with an ever smaller chance
                if (!onw && biome_next(p.seed, local, range++) == 0u) j = nw;
                if (!one && biome_next(p.seed, local, range++) == 0u) j = ne;
                if (!osw && biome_next(p.seed, local, range++) == 0u) j = sw;
                if (!ose && biome_next(p.seed, local, range) == 0u) j = se;
                if (biome_next(p.seed, local, 3u) == 0u) v = j;
                else v = j == forest ? forest : c;
            }
            break;
        }
        case kLayerIsland: {      // STPCrossLayer.h:26-37, STPIslandLayer.h:20-27
            const uint32_t c = up(x, z), north = up(x, z - 1), east = up(x + 1, z), south = up(x, z + 1), west = up(x - 1, z);
            v = c;
            if (biome_shallow(p.ids, c) && biome_shallow(p.ids, north) && biome_shallow(p.ids, east) &&
                biome_shallow(p.ids, south) && biome_shallow(p.ids, west)) {
                uint64_t local = biome_seed_local(p.seed, x, z);
                if (biome_next(p.seed, local, 2u) == 0u) v = p.ids.plains;
            }
            break;
        }
        default: {                // STPVoronoiLayer.h:56-101 with y = 0
            const int i0 = x - 2, k0 = z - 2;
            const int l = i0 >> 2, n = k0 >> 2;          // the cell; its y index is (0 - 2) >> 2 = -1
            const double fx = (double)(i0 & 3) / 4.0, fy = 0.5, fz = (double)(k0 & 3) / 4.0;   // ((0 - 2) & 3) / 4.0 = 0.5
            const double* jm = jit + (size_t)m * jit_map_stride;
            uint32_t best = 0u;
            double min = 0.0;
#pragma unroll
            for (uint32_t c = 0u; c < 8u; c++) {
                const uint32_t dx = (c >> 2) & 1u, dy = (c >> 1) & 1u, dz = c & 1u;
                const double* q = jm + ((((size_t)dy * pr.h) + (uint32_t)(n - pr.z0) + dz) * pr.w + (uint32_t)(l - pr.x0) + dx) * 3u;
                const double gx = dx ? fx - 1.0 : fx, gy = dy ? fy - 1.0 : fy, gz = dz ? fz - 1.0 : fz;
                const double a = __dadd_rn(gz, q[2]), b = __dadd_rn(gy, q[1]), cc = __dadd_rn(gx, q[0]);
                const double dist = __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(cc, cc));
                if (c == 0u || !(dist >= min)) {   // "if (ds[c] >= min) continue;" keeps the first of equal minima
                    best = c;
                    min = dist;
                }
            }
            v = up((best & 4u) ? l + 1 : l, (best & 1u) ? n + 1 : n);
            break;
        }
        }
        dst[(size_t)cz * row_stride + cx] = (uint16_t)v;
    }
}

}  // namespace shf
