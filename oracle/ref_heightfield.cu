// TEST INFRASTRUCTURE ONLY. The reference's OWN device code for the heightfield consumer, compiled where it lies:
// STPSimplexNoise.cu (simplex2D / simplex2DFractal) and STPSingleHistogramWrapper (bin walk) from /root/reference, driven
// by a kernel that restates generateMultiBiomeHeightmap (SuperDemo+/Script/STPMultiHeightGenerator.cu:35-71; that script
// itself only builds under NVRTC with the demo's __constant__ symbols, and its table holds two biomes). Built by
// oracle/Makefile into oracle/_ref/libshf_ref_height.so; runs on the GPU box only.
#include <cuda_runtime.h>
#include <cstdint>

#include <SuperAlgorithm+Device/STPSingleHistogramWrapper.cuh>
#include "STPSimplexNoise.cu"  // -I <reference>/SuperTerrain+/SuperAlgorithm+/Device/Private

using namespace SuperTerrainPlus::STPAlgorithm;

struct RefBiomeProperty {  // layout of STPDemo::STPBiomeProperty
    float Scale;
    unsigned int Octave;
    float Persistence, Lacunarity, Depth, Variation;
};

__global__ void refHeightKernel(float* height, STPSingleHistogram hist, uint2 dim, float2 half, const RefBiomeProperty* table,
                                unsigned int n_table, STPPermutation perm, float2 offset) {
    const unsigned int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dim.x || y >= dim.y) return;
    const unsigned int index = x + y * dim.x;
    float h = 0.0f;
    STPSingleHistogramWrapper::iterate(hist, index, [&](SuperTerrainPlus::STPSample_t biomeID, float weight) {
        if (biomeID >= n_table) return;
        const RefBiomeProperty& b = table[biomeID];
        STPSimplexNoise::STPFractalSimplexInformation desc = {};
        desc.Persistence = b.Persistence;
        desc.Lacunarity = b.Lacunarity;
        desc.Octave = b.Octave;
        desc.Scale = b.Scale;
        desc.Offset = offset;
        desc.HalfDimension = half;
        h += weight * (STPSimplexNoise::simplex2DFractal(perm, 1.0f * x, 1.0f * y, desc) * b.Variation + b.Depth);
    });
    height[index] = h;
}

#define CK(x) do { if ((x) != cudaSuccess) return 1; } while (0)

extern "C" int ref_heightfield_run(const void* bins, const uint32_t* offsets, uint64_t n_bins, uint32_t W, uint32_t H,
                                   const void* table, uint32_t n_table, const unsigned char* perm512, const float* grad,
                                   uint32_t grad_size, float off_x, float off_y, float* height) {
    void *d_bins, *d_off, *d_table, *d_perm, *d_grad, *d_h;
    CK(cudaMalloc(&d_bins, n_bins * 8 + 8));
    CK(cudaMalloc(&d_off, ((size_t)W * H + 1) * 4));
    CK(cudaMalloc(&d_table, (size_t)n_table * sizeof(RefBiomeProperty)));
    CK(cudaMalloc(&d_perm, 512));
    CK(cudaMalloc(&d_grad, (size_t)grad_size * 8));
    CK(cudaMalloc(&d_h, (size_t)W * H * 4));
    CK(cudaMemcpy(d_bins, bins, n_bins * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_off, offsets, ((size_t)W * H + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_table, table, (size_t)n_table * sizeof(RefBiomeProperty), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_perm, perm512, 512, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_grad, grad, (size_t)grad_size * 8, cudaMemcpyHostToDevice));
    STPSingleHistogram hist{static_cast<const STPSingleHistogram::STPBin*>(d_bins), static_cast<const unsigned int*>(d_off)};
    STPPermutation pm{static_cast<unsigned char*>(d_perm), static_cast<float*>(d_grad), grad_size};
    const dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    refHeightKernel<<<grid, block>>>(static_cast<float*>(d_h), hist, make_uint2(W, H), make_float2(W / 2.0f, H / 2.0f),
                                     static_cast<const RefBiomeProperty*>(d_table), n_table, pm, make_float2(off_x, off_y));
    CK(cudaGetLastError());
    CK(cudaMemcpy(height, d_h, (size_t)W * H * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_bins); cudaFree(d_off); cudaFree(d_table); cudaFree(d_perm); cudaFree(d_grad); cudaFree(d_h);
    return 0;
}
