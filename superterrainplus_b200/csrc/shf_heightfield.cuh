// shf_heightfield.cuh -- device-side consumer of the filter output: the multi-biome
// heightfield of SuperDemo+/Script/STPMultiHeightGenerator.cu:35-71, fed from the histogram while it is still in HBM.
//
//   height(x, y) = sum over the pixel's bins, IN BIN ORDER, of weight * (fractal(x, y; biome) * variation + depth)
//   fractal      = saturate((sum_o simplex2D(sample_o) * amp_o + A) / (2A)),  A = sum_o amp_o        (STPSimplexNoise.cu:84-109)
//   simplex2D    = Gustavson-style 2D simplex noise over a 512-entry permutation and an N-entry unit-gradient table
//                                                                                                  (STPSimplexNoise.cu:18-82)
// The reference compiles its version with NVRTC and lets the compiler contract multiply-adds as it likes, so its low
// bits are compiler-defined. Here every rounding is spelled out (explicit fma / mul / add intrinsics, IEEE division),
// which makes the result reproducible bit for bit by the CPU oracle (oracle/shf_heightfield_oracle.c uses fmaf in the
// same places); against the reference's own device code the difference is a few ulp of the [0, 1] noise range.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/shf_b200.h"

namespace shf {

struct HeightGeo {
    uint32_t W, H;
    uint32_t n_table;      // entries of the biome property table (indexed by sample value)
    uint32_t grad_size;    // gradient table entries
    float half_x, half_y;  // W / 2, H / 2 (STPCommonCompiler.cpp:236)
};

__device__ __forceinline__ int hf_floori(float x) {  // STPSimplexNoise.cu:14-16
    return x > 0.0f ? (int)x : (int)__fsub_rn(x, 1.0f);
}

// STPSimplexNoise.cu:18-82 with explicit roundings
// `pgrad[i]` = perm[i] % grad_size, tabulated once per CTA (the reference takes the modulo at every corner)
__device__ __forceinline__ float hf_simplex2d(const unsigned char* __restrict__ perm, const unsigned char* __restrict__ pgrad,
                                              const float2* __restrict__ grad, float x, float y) {
    const float F2 = 0.3660254038f, G2 = 0.2113248654f, H2 = -1.0f + 2.0f * 0.2113248654f;
    const float s = __fmul_rn(__fadd_rn(x, y), F2);
    const int i = hf_floori(__fadd_rn(x, s)), j = hf_floori(__fadd_rn(y, s));
    const float t = __fmul_rn((float)(i + j), G2);
    const float X0 = __fsub_rn((float)i, t), Y0 = __fsub_rn((float)j, t);
    float dx[3], dy[3];
    dx[0] = __fsub_rn(x, X0);
    dy[0] = __fsub_rn(y, Y0);
    const uint32_t i1 = dx[0] > dy[0] ? 1u : 0u, j1 = 1u - i1;
    dx[1] = __fadd_rn(__fsub_rn(dx[0], (float)i1), G2);
    dy[1] = __fadd_rn(__fsub_rn(dy[0], (float)j1), G2);
    dx[2] = __fadd_rn(dx[0], H2);
    dy[2] = __fadd_rn(dy[0], H2);
    const uint32_t ii = (uint32_t)i & 255u, jj = (uint32_t)j & 255u;
    uint32_t gi[3];
    gi[0] = pgrad[ii + perm[jj]];
    gi[1] = pgrad[ii + i1 + perm[jj + j1]];
    gi[2] = pgrad[ii + 1u + perm[jj + 1u]];
    float corner[3];
#pragma unroll
    for (int v = 0; v < 3; v++) {
        // the reference returns 0 for w <= 0; clamping w to 0 gives (+-)0 * dot = (+-)0 instead, and a signed zero
        // changes neither the corner sum nor the fma into the octave sum -- no branch, same bits
        float w = fmaxf(__fsub_rn(__fsub_rn(0.5f, __fmul_rn(dx[v], dx[v])), __fmul_rn(dy[v], dy[v])), 0.0f);
        w = __fmul_rn(w, w);
        const float2 gr = grad[gi[v]];
        const float dot = __fmaf_rn(gr.x, dx[v], __fmul_rn(gr.y, dy[v]));
        corner[v] = __fmul_rn(__fmul_rn(w, w), dot);
    }
    return __fmul_rn(70.0f, __fadd_rn(__fadd_rn(corner[0], corner[1]), corner[2]));
}

// STPSimplexNoise.cu:84-109 (initial amplitude and frequency 1, STPSimplexNoise.cuh:46-51)
__device__ __forceinline__ float hf_fractal(const unsigned char* __restrict__ perm, const unsigned char* __restrict__ pgrad,
                                            const float2* __restrict__ grad, float bx, float by, const shf_biome_property& p) {
    float fractal = 0.0f, amplitude = 1.0f, frequency = 1.0f, range = 0.0f;
    // (x / scale) * frequency per octave in the reference: the quotient does not depend on the octave, so the two IEEE
    // divisions are done once per bin -- same operands, same roundings
    const float qx = __fdiv_rn(bx, p.scale), qy = __fdiv_rn(by, p.scale);
    for (uint32_t o = 0u; o < p.octave; o++) {
        const float sx = __fmul_rn(qx, frequency), sy = __fmul_rn(qy, frequency);
        fractal = __fmaf_rn(hf_simplex2d(perm, pgrad, grad, sx, sy), amplitude, fractal);
        range = __fadd_rn(range, amplitude);
        amplitude = __fmul_rn(amplitude, p.persistence);
        frequency = __fmul_rn(frequency, p.lacunarity);
    }
    return __saturatef(__fdiv_rn(__fadd_rn(fractal, range), __fmul_rn(2.0f, range)));
}

// One thread per pixel; the lookup tables sit in shared memory. blockIdx.y = chunk of the batch.
// smem: perm[512] | pgrad[512] | grad[grad_size] float2 | table[n_table] shf_biome_property
__global__ void __launch_bounds__(256) heightfield_kernel(HeightGeo g, const uint2* __restrict__ bins,
                                                          const uint32_t* __restrict__ hso,
                                                          const uint64_t* __restrict__ chunkbase, uint32_t first_chunk,
                                                          const shf_biome_property* __restrict__ table,
                                                          const unsigned char* __restrict__ perm_g,
                                                          const float* __restrict__ grad_g,
                                                          const float2* __restrict__ offsets, float* __restrict__ height) {
    extern __shared__ __align__(16) uint8_t smem[];
    unsigned char* perm = smem;
    unsigned char* pgrad = smem + 512;
    float2* grad = reinterpret_cast<float2*>(smem + 1024);
    shf_biome_property* tab = reinterpret_cast<shf_biome_property*>(smem + 1024 + (size_t)g.grad_size * 8);
    for (uint32_t i = threadIdx.x; i < 512u; i += blockDim.x) {
        const unsigned char pv = perm_g[i];
        perm[i] = pv;
        pgrad[i] = (unsigned char)((uint32_t)pv % g.grad_size);
    }
    for (uint32_t i = threadIdx.x; i < g.grad_size; i += blockDim.x) grad[i] = make_float2(grad_g[2u * i], grad_g[2u * i + 1u]);
    for (uint32_t i = threadIdx.x; i < g.n_table; i += blockDim.x) tab[i] = table[i];
    __syncthreads();
    const uint32_t chunk = first_chunk + blockIdx.y;
    const uint32_t npx = g.W * g.H;
    const uint2* cb = bins + chunkbase[chunk];
    const uint32_t* ho = hso + (size_t)chunk * ((size_t)npx + 1u);
    const float2 off = offsets[blockIdx.y];
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += gridDim.x * blockDim.x) {
        const uint32_t begin = ho[p], end = ho[p + 1u];  // STPSingleHistogramWrapper.inl:9-12
        const float x = (float)(p % g.W), y = (float)(p / g.W);
        // the sample position before scaling is the same for every bin of the pixel (STPSimplexNoise.cu:93-96)
        const float bx = __fadd_rn(__fsub_rn(x, g.half_x), off.x), by = __fadd_rn(__fsub_rn(y, g.half_y), off.y);
        float h = 0.0f;
        for (uint32_t b = begin; b < end; b++) {
            const uint2 bin = cb[b];
            const uint32_t item = bin.x & 0xFFFFu;
            if (item >= g.n_table) continue;  // the reference indexes its table unchecked; contribute nothing instead
            const shf_biome_property pr = tab[item];
            const float noise = hf_fractal(perm, pgrad, grad, bx, by, pr);
            h = __fmaf_rn(__uint_as_float(bin.y), __fmaf_rn(noise, pr.variation, pr.depth), h);
        }
        height[(size_t)blockIdx.y * npx + p] = h;
    }
}

}  // namespace shf
