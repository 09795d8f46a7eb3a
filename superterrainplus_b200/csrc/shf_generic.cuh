// shf_generic.cuh -- the wide path of the single histogram filter: any radius, up to kGenericMaxBiomes distinct sample
// values per neighbourhood. Used when the event-list path of shf_events.cuh does not cover the shape (2r+1 > 511,
// more than 256 distinct values, or a ring that does not fit shared memory). Same semantics, same outputs; it trades
// the shared-memory ring of vertical window counts for re-counting the entering and leaving columns cell by cell
// (O(r) work per pixel instead of O(bins)), which needs no storage proportional to radius x biomes.
//
//   vstart_generic_kernel : one thread per column, per-biome (chain start, last row) state in global scratch
//   march_generic_kernel  : one CTA per output row. Dense per-biome window counts and the ordered bin list live in shared
//                           memory; a step adds the 2r+1 cells of column c (a 0 -> 1 transition is a birth), removes
//                           the cells of column c - (2r+1) (a 1 -> 0 transition is a death), compacts the list if
//                           something died (SHF.cpp:435-445), appends the newborn ordered by their vertical chain start
//                           (SHF.cpp:411-415) and emits pixel c - 2r. COUNT mode only totals the bins of the row.
#pragma once
#include "shf_kernels.cuh"

namespace shf {

constexpr int kGenericThreads = 128;
constexpr uint32_t kGenericMaxBiomes = 16384;

// state(n, s, c) = chain start row << 16 | last row seen, 0xFFFFFFFF = the biome has not occurred yet in this column
__global__ void __launch_bounds__(128) vstart_generic_kernel(Geo g, uint32_t first_chunk, uint32_t n_biomes,
                                                             const uint16_t* __restrict__ cmap,
                                                             uint16_t* __restrict__ vstart, uint32_t* __restrict__ state) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.PW) return;
    const uint32_t n = first_chunk + blockIdx.y;
    const uint16_t* col = cmap + (size_t)n * g.PH * g.P + c;
    uint16_t* vcol = vstart + (size_t)n * g.PH * g.P + c;
    uint32_t* st = state + (size_t)blockIdx.y * n_biomes * g.PW + c;
    const uint32_t span = g.span;
    uint32_t s_next = col[0];
    for (uint32_t p = 0; p < g.PH; p++) {
        const uint32_t s = s_next;
        if (p + 1u < g.PH) s_next = col[(size_t)(p + 1u) * g.P];
        const uint32_t v = st[(size_t)s * g.PW];
        const uint32_t last = v & 0xFFFFu;
        const uint32_t start = (v == 0xFFFFFFFFu || p - last > span) ? p : (v >> 16);
        st[(size_t)s * g.PW] = (start << 16) | p;
        vcol[(size_t)p * g.P] = (uint16_t)start;
    }
}

// exclusive prefix over one value per thread; every thread gets its prefix and the block total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, off);
        if (lane >= (uint32_t)off) incl += t;
    }
    __syncthreads();  // warp_sums may still be read from the previous use
    if (lane == 31u) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0u, sum = 0u;
    for (uint32_t w = 0; w < n_warps; w++) {
        const uint32_t t = warp_sums[w];
        if (w < warp) before += t;
        sum += t;
    }
    total = sum;
    return before + incl - v;
}

template <bool EMIT>
__global__ void __launch_bounds__(kGenericThreads)
    march_generic_kernel(Geo g, uint32_t n_biomes, const uint16_t* __restrict__ cmap, const uint16_t* __restrict__ vstart,
                         const uint16_t* __restrict__ dict, uint32_t dict_stride, const uint32_t* __restrict__ rowbase,
                         const uint64_t* __restrict__ chunkbase, uint2* __restrict__ bins, uint32_t* __restrict__ hso,
                         uint32_t* __restrict__ rowtotal) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t T = blockDim.x, tid = threadIdx.x;
    const uint32_t y = blockIdx.x, n = blockIdx.y;
    const uint32_t span = g.span, two_r = 2u * g.r, PW = g.PW;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem);              // [n_biomes] window count by compact id
    uint32_t* ctrl = cnt + n_biomes;                                // [0..1] newborn counters, [2..3] death flags, by step parity
    uint32_t* warp_sums = ctrl + 4;                                 // [32]
    uint16_t* list_a = reinterpret_cast<uint16_t*>(warp_sums + 32); // [n_biomes] ordered bin list (compact ids)
    uint16_t* list_b = list_a + n_biomes;                           // [n_biomes] compaction target
    uint16_t* born_id = list_b + n_biomes;                          // [span]
    uint16_t* born_key = born_id + span;                            // [span]
    for (uint32_t i = tid; i < n_biomes; i += T) cnt[i] = 0u;
    if (tid < 4u) ctrl[tid] = 0u;
    const uint16_t* cm = cmap + ((size_t)n * g.PH + y) * g.P;  // window rows of this output row start at halo row y
    const uint16_t* vs = vstart + ((size_t)n * g.PH + y) * g.P;
    const uint16_t* dc = dict + (size_t)n * dict_stride;
    uint16_t* list = list_a;
    uint16_t* other = list_b;
    uint32_t n_list = 0u, rowpos = 0u;
    const uint32_t row_first = EMIT ? rowbase[(size_t)n * g.H + y] : 0u;
    uint2* dst = EMIT ? bins + (size_t)chunkbase[n] + row_first : nullptr;
    uint32_t* hso_row = EMIT ? hso + (size_t)n * ((size_t)g.W * g.H + 1u) + (size_t)y * g.W : nullptr;
    const float inv = g.inv_total;

    for (uint32_t c = 0u; c < PW; c++) {
        const uint32_t par = c & 1u;
        __syncthreads();  // the previous step's emit has read cnt / list
        // ---- add column c (SHF.cpp:650-657) ----
        for (uint32_t i = tid; i < span; i += T) {
            const uint32_t s = cm[(size_t)i * g.P + c];
            if (atomicAdd(&cnt[s], 1u) == 0u) {
                const uint32_t slot = atomicAdd(&ctrl[par], 1u);
                born_id[slot] = (uint16_t)s;
                // every occurrence of s inside one window belongs to the same vertical chain
                born_key[slot] = vs[(size_t)i * g.P + c];
            }
        }
        __syncthreads();
        // ---- remove column c - span (SHF.cpp:659-666) ----
        if (c >= span) {
            for (uint32_t i = tid; i < span; i += T) {
                const uint32_t s = cm[(size_t)i * g.P + (c - span)];
                if (atomicSub(&cnt[s], 1u) == 1u) ctrl[2u + par] = 1u;
            }
        }
        __syncthreads();
        const uint32_t nb = ctrl[par];
        const bool died = ctrl[2u + par] != 0u;
        if (tid == 0u) {  // counters of the next step (last read before the previous step's closing barrier)
            ctrl[par ^ 1u] = 0u;
            ctrl[2u + (par ^ 1u)] = 0u;
        }
        if (died) {
            // stable compaction of the survivors
            const uint32_t per = (n_list + T - 1u) / T;
            const uint32_t lo = min(tid * per, n_list), hi = min(lo + per, n_list);
            uint32_t alive = 0u;
            for (uint32_t i = lo; i < hi; i++) alive += cnt[list[i]] != 0u;
            uint32_t total;
            uint32_t at = block_exclusive_scan(alive, warp_sums, total);
            for (uint32_t i = lo; i < hi; i++) {
                const uint16_t s = list[i];
                if (cnt[s] != 0u) other[at++] = s;
            }
            uint16_t* t = list;
            list = other;
            other = t;
            n_list = total;
        }
        if (nb) {
            // newborn bins join in the order of the column's own vertical histogram: ascending chain start row
            for (uint32_t j = tid; j < nb; j += T) {
                const uint32_t key = born_key[j];
                uint32_t rank = 0u;
                for (uint32_t i = 0u; i < nb; i++) rank += born_key[i] < key;
                list[n_list + rank] = born_id[j];
            }
            n_list += nb;
        }
        __syncthreads();
        if (c >= two_r) {
            const uint32_t x = c - two_r;
            if (EMIT) {
                for (uint32_t i = tid; i < n_list; i += T) {
                    const uint32_t s = list[i];
                    dst[rowpos + i] = make_uint2((uint32_t)dc[s], __float_as_uint(__fmul_rn(__uint2float_rn(cnt[s]), inv)));
                }
                if (tid == 0u) hso_row[x] = row_first + rowpos;
            }
            rowpos += n_list;
        }
    }
    if (!EMIT && tid == 0u) rowtotal[(size_t)n * g.H + y] = rowpos;
}

inline size_t march_generic_smem(uint32_t n_biomes, uint32_t span) {
    return (size_t)n_biomes * 4 + 4 * 4 + 32 * 4 + (size_t)n_biomes * 2 * 2 + (size_t)span * 2 * 2 + 16;
}

}  // namespace shf
