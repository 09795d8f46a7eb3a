/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the multi-biome heightfield that consumes the filter's histograms
 * (SURVEY.md section 8 row f1). Only tests/ and __graft_entry__.smoke() may load it, as the checker.
 *
 * Follows /root/reference/SuperDemo+/Script/STPMultiHeightGenerator.cu:35-71 (per pixel: sum over the bins in bin
 * order of weight * (fractal * variation + depth)), /root/reference/SuperTerrain+/SuperAlgorithm+/Device/Private/
 * STPSimplexNoise.cu:18-82 (simplex2D) and :84-109 (simplex2DFractal), and STPSingleHistogramWrapper.inl:5-20 (bin walk).
 *
 * Parity status: the reference compiles this arithmetic for the GPU with NVRTC and leaves multiply-add contraction to
 * the compiler, so its last bits are compiler-defined. This restatement fixes one rounding sequence (fmaf where a fused
 * multiply-add is used, plain IEEE float operations elsewhere; build with -ffp-contract=off) which the CUDA kernel
 * (superterrainplus_b200/csrc/shf_heightfield.cuh) reproduces bit for bit. It is PINNED against the reference's own device
 * code (compiled from /root/reference into oracle/_ref/libshf_ref_height.so, run on the GPU box) within 2e-5 absolute on
 * heights of order 1 (measured maximum 3.7e-6, at noise offsets of 1e5): tests/test_heightfield_gpu.py, and against vectors of that run committed in tests/golden/.
 */
#include <math.h>
#include <stdint.h>

typedef struct {
    float scale;
    uint32_t octave;
    float persistence, lacunarity, depth, variation;
} hf_biome_property; /* STPDemo::STPBiomeProperty, SuperDemo+/World/Biomes/STPBiomeProperty.hpp:10-27 */

typedef struct {
    uint16_t item;
    float weight;
} hf_bin;

static int floori(float x) { return x > 0.0f ? (int)x : (int)(x - 1.0f); } /* STPSimplexNoise.cu:14-16 */

static float simplex2d(const unsigned char* perm, const float* grad, uint32_t grad_size, float x, float y) {
    const float F2 = 0.3660254038f, G2 = 0.2113248654f, H2 = -1.0f + 2.0f * 0.2113248654f;
    const float s = (x + y) * F2;
    const int i = floori(x + s), j = floori(y + s);
    const float t = (float)(i + j) * G2;
    const float X0 = (float)i - t, Y0 = (float)j - t;
    float dx[3], dy[3], corner[3];
    dx[0] = x - X0;
    dy[0] = y - Y0;
    const uint32_t i1 = dx[0] > dy[0] ? 1u : 0u, j1 = 1u - i1;
    dx[1] = (dx[0] - (float)i1) + G2;
    dy[1] = (dy[0] - (float)j1) + G2;
    dx[2] = dx[0] + H2;
    dy[2] = dy[0] + H2;
    const uint32_t ii = (uint32_t)i & 255u, jj = (uint32_t)j & 255u;
    uint32_t gi[3];
    gi[0] = perm[ii + perm[jj]] % grad_size;
    gi[1] = perm[ii + i1 + perm[jj + j1]] % grad_size;
    gi[2] = perm[ii + 1u + perm[jj + 1u]] % grad_size;
    for (int v = 0; v < 3; v++) {
        float w = (0.5f - dx[v] * dx[v]) - dy[v] * dy[v];
        if (w <= 0.0f) {
            corner[v] = 0.0f;
        } else {
            w = w * w;
            const float dot = fmaf(grad[2u * gi[v]], dx[v], grad[2u * gi[v] + 1u] * dy[v]);
            corner[v] = (w * w) * dot;
        }
    }
    return 70.0f * ((corner[0] + corner[1]) + corner[2]);
}

static float saturate(float v) { return v > 0.0f ? (v < 1.0f ? v : 1.0f) : 0.0f; } /* NaN -> 0 like __saturatef */

static float fractal2d(const unsigned char* perm, const float* grad, uint32_t grad_size, float x, float y,
                       const hf_biome_property* p, float off_x, float off_y, float half_x, float half_y) {
    float fractal = 0.0f, amplitude = 1.0f, frequency = 1.0f, range = 0.0f;
    const float bx = (x - half_x) + off_x, by = (y - half_y) + off_y;
    for (uint32_t o = 0u; o < p->octave; o++) {
        const float sx = (bx / p->scale) * frequency, sy = (by / p->scale) * frequency;
        fractal = fmaf(simplex2d(perm, grad, grad_size, sx, sy), amplitude, fractal);
        range = range + amplitude;
        amplitude = amplitude * p->persistence;
        frequency = frequency * p->lacunarity;
    }
    return saturate((fractal + range) / (2.0f * range));
}

/* bins / offsets: one chunk's histogram (offsets relative to bins[0]); height: W*H floats, row-major */
/* only pixels [p_begin, p_end) are evaluated (the rest of `height` is left untouched) */
void shf_heightfield_oracle(const hf_bin* bins, const uint32_t* offsets, uint32_t W, uint32_t H,
                            const hf_biome_property* table, uint32_t n_table, const unsigned char* perm, const float* grad,
                            uint32_t grad_size, float off_x, float off_y, float* height, uint32_t p_begin, uint32_t p_end) {
    const float half_x = (float)W / 2.0f, half_y = (float)H / 2.0f;
    if (p_end > W * H) p_end = W * H;
    for (uint32_t p = p_begin; p < p_end; p++) {
        const float x = (float)(p % W), y = (float)(p / W);
        float h = 0.0f;
        for (uint32_t b = offsets[p]; b < offsets[p + 1u]; b++) {
            if (bins[b].item >= n_table) continue;
            const hf_biome_property* pr = &table[bins[b].item];
            const float noise = fractal2d(perm, grad, grad_size, x, y, pr, off_x, off_y, half_x, half_y);
            h = fmaf(bins[b].weight, fmaf(noise, pr->variation, pr->depth), h);
        }
        height[p] = h;
    }
}
