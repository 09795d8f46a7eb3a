/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's biome-map producer, the checker of the device
 * producer in superterrainplus_b200/csrc/shf_biome.cuh. Nothing in the product path links, loads or calls this file.
 *
 * What it restates (reference paths under /root/reference):
 *   SuperTerrain+/SuperTerrain+/Private/World/Diversity/STPBiomeFactory.cpp:24-42   the per-pixel loop over the layer tree
 *   SuperTerrain+/SuperTerrain+/Private/World/Diversity/STPLayer.cpp:118-196        sampler, layer / local seeds, mixSeed
 *   SuperDemo+/World/Layers/STPContinentLayer.h:17-23, STPScaleLayer.h:31-98, STPXCrossLayer.h:26-37,
 *   STPLandLayer.h:20-70, STPCrossLayer.h:26-37, STPIslandLayer.h:20-27, STPVoronoiLayer.h:19-101   the layers the
 *                                                                                   demo's chain is built from
 *   SuperDemo+/World/Biomes/STPBiomeRegistry.cpp:112-144                            isShallowOcean, CAS
 * Structure differs on purpose from both the reference (virtual layer objects) and the device code (dense level-by-level
 * grids): a table-driven recursion with a small direct-mapped memo per layer. Layers are pure functions of (x, z), so
 * the memo (like the reference's STPLayerCache, STPLayer.cpp:77-92) only saves time.
 *
 * Pinned by tests/test_biome_oracle.py against the reference's own compiled chain (oracle/_ref/libbiome_ref.so, built
 * from /root/reference by oracle/Makefile) and against its stored outputs in tests/golden/biome_vectors.npz.
 * Floating point: the Voronoi layer compares sums of squared doubles; this file is compiled with -ffp-contract=off so
 * that every operation rounds once, as the C++ source is written. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { LAYER_CONTINENT = 0, LAYER_SCALE_NORMAL = 1, LAYER_SCALE_FUZZY = 2, LAYER_LAND = 3, LAYER_ISLAND = 4, LAYER_VORONOI = 5 };

typedef struct {
    uint32_t kind;
    uint32_t parent;
    uint64_t salt;
} biome_layer_desc;

enum { ID_OCEAN, ID_PLAINS, ID_FOREST, ID_FROZEN_OCEAN, ID_WARM_OCEAN, ID_LUKEWARM_OCEAN, ID_COLD_OCEAN, ID_COUNT };

#define MEMO 4096u

typedef struct {
    uint32_t kind, parent;
    uint64_t seed;
    int32_t* kx;
    int32_t* kz;
    uint16_t* val;
    uint8_t* used;
} layer_t;

typedef struct {
    layer_t* layer;
    uint32_t n;
    uint64_t voronoi_seed;
    uint16_t id[ID_COUNT];
} chain_t;

/* STPLayer.cpp:178-182 */
static uint64_t mix_seed(uint64_t s, int64_t fac) {
    s *= s * 6364136223846793005ull + 1442695040888963407ull;
    s += (uint64_t)fac;
    return s;
}

/* STPLayer.cpp:143-151 */
static uint64_t seed_layer(uint64_t global_seed, uint64_t salt) {
    uint64_t mid = mix_seed(salt, (int64_t)salt);
    mid = mix_seed(mid, (int64_t)mid);
    mid = mix_seed(mid, (int64_t)mid);
    uint64_t s = mix_seed(global_seed, (int64_t)mid);
    s = mix_seed(s, (int64_t)mid);
    s = mix_seed(s, (int64_t)mid);
    return s;
}

/* STPLayer.cpp:153-159 */
static uint64_t seed_local(uint64_t layer_seed, int x, int z) {
    uint64_t s = mix_seed(layer_seed, x);
    s = mix_seed(s, z);
    s = mix_seed(s, x);
    s = mix_seed(s, z);
    return s;
}

/* STPLayer.cpp:118-127: the value, then the sequence advances */
static uint16_t next_value(uint64_t layer_seed, uint64_t* local, uint16_t range) {
    const uint16_t v = (uint16_t)((*local >> 24) % (uint64_t)range);
    *local = mix_seed(*local, (int64_t)layer_seed);
    return v;
}

static int shallow_ocean(const chain_t* c, uint16_t v) { /* STPBiomeRegistry.cpp:112-116 */
    return v == c->id[ID_OCEAN] || v == c->id[ID_FROZEN_OCEAN] || v == c->id[ID_WARM_OCEAN] ||
           v == c->id[ID_LUKEWARM_OCEAN] || v == c->id[ID_COLD_OCEAN];
}

static uint16_t retrieve(chain_t* c, uint32_t li, int x, int z);

/* STPScaleLayer.h:31-98 */
static uint16_t scale_sample(chain_t* c, const layer_t* L, int x, int z) {
    const uint16_t i = retrieve(c, L->parent, x >> 1, z >> 1);
    const int xb = x & 1, zb = z & 1;
    uint64_t local = seed_local(L->seed, x & -2, z & -2);
    if (xb == 0 && zb == 0) return i;
    const uint16_t l = retrieve(c, L->parent, x >> 1, (z + 1) >> 1);
    const uint16_t m = next_value(L->seed, &local, 2) == 0 ? i : l;
    if (xb == 0) return m;
    const uint16_t n = retrieve(c, L->parent, (x + 1) >> 1, z >> 1);
    const uint16_t o = next_value(L->seed, &local, 2) == 0 ? i : n;
    if (zb == 0) return o;
    const uint16_t p = retrieve(c, L->parent, (x + 1) >> 1, (z + 1) >> 1);
    /* sample(center = i, e = n, s = l, se = p) */
    const uint16_t pick = next_value(L->seed, &local, 4);
    const uint16_t ret = pick == 0 ? i : pick == 1 ? n : pick == 2 ? l : p;
    if (L->kind == LAYER_SCALE_FUZZY) return ret;
    const uint16_t center = i, e = n, s = l, se = p;
    if (e == s && e == se) return e;
    if (center == e && (center == se || s != se)) return center;
    if (center == s && (center == se || e != se)) return center;
    if (center == se && e != s) return center;
    if (e == s && center != se) return e;
    if (e == se && center != s) return e;
    if (s == se && center != e) return s;
    return ret;
}

/* STPXCrossLayer.h:26-37 + STPLandLayer.h:20-70 */
static uint16_t land_sample(chain_t* c, const layer_t* L, int x, int z) {
    const uint16_t center = retrieve(c, L->parent, x, z);
    const uint16_t ne = retrieve(c, L->parent, x + 1, z - 1);
    const uint16_t se = retrieve(c, L->parent, x + 1, z + 1);
    const uint16_t sw = retrieve(c, L->parent, x - 1, z + 1);
    const uint16_t nw = retrieve(c, L->parent, x - 1, z - 1);
    uint64_t local = seed_local(L->seed, x, z);
    const uint16_t forest = c->id[ID_FOREST];
    const int oc = shallow_ocean(c, center);
    const int osw = shallow_ocean(c, sw), ose = shallow_ocean(c, se), one = shallow_ocean(c, ne), onw = shallow_ocean(c, nw);
    if (!oc || (osw && ose && one && onw)) {
        if (oc || (!osw && !ose && !one && !onw) || next_value(L->seed, &local, 5) != 0) return center;
        /* CAS(comparator, comparable, fallback) = comparator == comparable ? comparable : fallback */
        if (onw) return center == forest ? forest : nw;
        if (osw) return center == forest ? forest : sw;
        if (one) return center == forest ? forest : ne;
        if (ose) return center == forest ? forest : se;
        return center;
    }
    uint16_t i = 1, j = 1;
    if (!onw && next_value(L->seed, &local, i++) == 0) j = nw;
    if (!one && next_value(L->seed, &local, i++) == 0) j = ne;
    if (!osw && next_value(L->seed, &local, i++) == 0) j = sw;
    if (!ose && next_value(L->seed, &local, i) == 0) j = se;
    if (next_value(L->seed, &local, 3) == 0) return j;
    return j == forest ? forest : center;
}

/* STPCrossLayer.h:26-37 + STPIslandLayer.h:20-27 */
static uint16_t island_sample(chain_t* c, const layer_t* L, int x, int z) {
    const uint16_t center = retrieve(c, L->parent, x, z);
    const uint16_t north = retrieve(c, L->parent, x, z - 1);
    const uint16_t east = retrieve(c, L->parent, x + 1, z);
    const uint16_t south = retrieve(c, L->parent, x, z + 1);
    const uint16_t west = retrieve(c, L->parent, x - 1, z);
    uint64_t local = seed_local(L->seed, x, z);
    if (shallow_ocean(c, center) && shallow_ocean(c, north) && shallow_ocean(c, east) && shallow_ocean(c, south) &&
        shallow_ocean(c, west) && next_value(L->seed, &local, 2) == 0)
        return c->id[ID_PLAINS];
    return center;
}

/* STPVoronoiLayer.h:19-47 */
static double voronoi_dist(uint64_t seed, int x, int y, int z, double xf, double yf, double zf) {
    uint64_t m = mix_seed(seed, x);
    m = mix_seed(m, y);
    m = mix_seed(m, z);
    m = mix_seed(m, x);
    m = mix_seed(m, y);
    m = mix_seed(m, z);
    const double d = ((double)(uint32_t)((m >> 24) % 1024ull) / 1024.0 - 0.5) * 0.9;
    m = mix_seed(m, (int64_t)seed);
    const double e = ((double)(uint32_t)((m >> 24) % 1024ull) / 1024.0 - 0.5) * 0.9;
    m = mix_seed(m, (int64_t)seed);
    const double f = ((double)(uint32_t)((m >> 24) % 1024ull) / 1024.0 - 0.5) * 0.9;
    const double a = zf + f, b = yf + e, cc = xf + d;
    return a * a + b * b + cc * cc;
}

/* STPVoronoiLayer.h:56-101, Is3D == false (every Voronoi layer of the demo chain): y = 0 */
static uint16_t voronoi_sample(chain_t* c, const layer_t* L, int x, int z) {
    const int ijk[3] = {x - 2, 0 - 2, z - 2};
    const int lmn[3] = {ijk[0] >> 2, ijk[1] >> 2, ijk[2] >> 2};
    const double def[3] = {(double)(ijk[0] & 3) / 4.0, (double)(ijk[1] & 3) / 4.0, (double)(ijk[2] & 3) / 4.0};
    double ds[8];
    for (unsigned k = 0; k < 8u; k++) {
        const int b0 = (k & 4u) == 0u, b1 = (k & 2u) == 0u, b2 = (k & 1u) == 0u;
        ds[k] = voronoi_dist(c->voronoi_seed, b0 ? lmn[0] : lmn[0] + 1, b1 ? lmn[1] : lmn[1] + 1, b2 ? lmn[2] : lmn[2] + 1,
                             b0 ? def[0] : def[0] - 1.0, b1 ? def[1] : def[1] - 1.0, b2 ? def[2] : def[2] - 1.0);
    }
    unsigned index = 0;
    double min = ds[0];
    for (unsigned k = 1; k < 8u; k++) {
        if (ds[k] >= min) continue;
        index = k;
        min = ds[k];
    }
    return retrieve(c, L->parent, (index & 4u) == 0u ? lmn[0] : lmn[0] + 1, (index & 1u) == 0u ? lmn[2] : lmn[2] + 1);
}

static uint16_t sample(chain_t* c, uint32_t li, int x, int z) {
    const layer_t* L = &c->layer[li];
    switch (L->kind) {
        case LAYER_CONTINENT: { /* STPContinentLayer.h:17-23 */
            uint64_t local = seed_local(L->seed, x, z);
            return next_value(L->seed, &local, 10) == 0 ? c->id[ID_PLAINS] : c->id[ID_OCEAN];
        }
        case LAYER_SCALE_NORMAL:
        case LAYER_SCALE_FUZZY: return scale_sample(c, L, x, z);
        case LAYER_LAND: return land_sample(c, L, x, z);
        case LAYER_ISLAND: return island_sample(c, L, x, z);
        default: return voronoi_sample(c, L, x, z);
    }
}

static uint16_t retrieve(chain_t* c, uint32_t li, int x, int z) {
    layer_t* L = &c->layer[li];
    uint64_t h = ((uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)z * 0xC2B2AE3D27D4EB4Full);
    h ^= h >> 29;
    const uint32_t slot = (uint32_t)h & (MEMO - 1u);
    if (L->used[slot] && L->kx[slot] == x && L->kz[slot] == z) return L->val[slot];
    const uint16_t v = sample(c, li, x, z);
    L->used[slot] = 1;
    L->kx[slot] = x;
    L->kz[slot] = z;
    L->val[slot] = v;
    return v;
}

/* 0 ok, 1 bad chain description, 2 out of memory. The last layer is the root (STPAllLayers.cpp:74-76). */
int biome_oracle_run(const biome_layer_desc* desc, uint32_t n_layers, uint64_t global_seed, uint64_t voronoi_seed,
                     const uint16_t ids[ID_COUNT], int32_t offset_x, int32_t offset_z, uint32_t width, uint32_t height,
                     uint16_t* biomemap) {
    if (!desc || n_layers == 0 || !ids || !biomemap) return 1;
    for (uint32_t i = 0; i < n_layers; i++) {
        if (desc[i].kind > LAYER_VORONOI) return 1;
        if (desc[i].kind != LAYER_CONTINENT && desc[i].parent >= i) return 1;
    }
    chain_t c;
    c.n = n_layers;
    c.voronoi_seed = voronoi_seed;
    memcpy(c.id, ids, sizeof(c.id));
    c.layer = (layer_t*)calloc(n_layers, sizeof(layer_t));
    if (!c.layer) return 2;
    int st = 0;
    for (uint32_t i = 0; i < n_layers; i++) {
        layer_t* L = &c.layer[i];
        L->kind = desc[i].kind;
        L->parent = desc[i].parent;
        L->seed = seed_layer(global_seed, desc[i].salt);
        L->kx = (int32_t*)malloc(MEMO * sizeof(int32_t));
        L->kz = (int32_t*)malloc(MEMO * sizeof(int32_t));
        L->val = (uint16_t*)malloc(MEMO * sizeof(uint16_t));
        L->used = (uint8_t*)calloc(MEMO, 1);
        if (!L->kx || !L->kz || !L->val || !L->used) st = 2;
    }
    if (st == 0) {
        /* STPBiomeFactory.cpp:31-38 */
        for (uint32_t z = 0; z < height; z++)
            for (uint32_t x = 0; x < width; x++)
                biomemap[x + (size_t)z * width] = retrieve(&c, n_layers - 1u, (int)x + offset_x, (int)z + offset_z);
    }
    for (uint32_t i = 0; i < n_layers; i++) {
        free(c.layer[i].kx);
        free(c.layer[i].kz);
        free(c.layer[i].val);
        free(c.layer[i].used);
    }
    free(c.layer);
    return st;
}
