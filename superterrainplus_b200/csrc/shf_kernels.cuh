// shf_kernels.cuh -- sm_100a kernels of the single histogram filter.
//
// What is computed (semantics, not code, follow SuperTerrain+/SuperAlgorithm+/Host/Private/STPSingleHistogramFilter.cpp,
// "SHF.cpp"; the closed form is SURVEY.md Appendix A): for every pixel (x, y) of the centre chunk the histogram of the
// uint16 samples in the (2r+1)^2 window, as an ORDERED sparse bin list. The order is the order in which the reference's
// running accumulator (SHF.cpp:400-450) would hold the bins after its vertical (SHF.cpp:522-563) and horizontal
// (SHF.cpp:608-679) sliding passes.
//
// Layout of the computation on the GPU (all coordinates halo-relative: column c in [0, W+2r), row p in [0, H+2r)):
//   presence/dictionary : distinct sample values of a chunk -> dense ids 0..B-1 ("compact ids"), cmap = remapped map
//   vscan               : one thread per column walks down the rows and writes
//                           vstart(c,p)  = first row of the vertical presence chain the sample at (c,p) belongs to
//                                          (the reference's vertical accumulator keeps a bin alive across gaps <= 2r+1),
//                           base(c,t)    = per-biome counts of the vertical window of the first row of row-tile t
//   march<count>        : per row, slide the window left to right keeping the ordered bin list in registers, only
//                         counting bins -> bins per row -> exclusive scan = first-bin index of every row
//   march<emit>         : the same march, now writing HistogramStartOffset and the normalised bins in place
// The march kernel is the hot one: a CTA owns TY consecutive rows (one warp per row); the vertical window counts of
// the columns it walks over live in a shared-memory ring of (2r+1+NB) columns x TY rows x B bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shf {

constexpr int kDictWords = 2048;      // 65536 possible sample values / 32
constexpr int kBatchCols = 32;        // NB: columns produced per phase of the march kernel (one per lane)
constexpr int kVscanThreads = 64;     // columns per CTA in vscan
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint16_t kNoRow = 0xFFFFu;

struct Geo {
    uint32_t W, H, r, span;        // chunk map size, radius, 2r+1
    uint32_t PW, PH;               // W+2r, H+2r : the halo-extended region
    uint32_t P;                    // row pitch (elements) of cmap / vstart
    uint32_t n_chunks;
    uint32_t in_row_stride;        // input view: sample(n,p,c) = in[n*in_chunk_stride + p*in_row_stride + c]
    uint64_t in_chunk_stride;
    uint32_t TY, T;                // rows per march CTA, tiles per chunk
    uint32_t K, Bpad;              // 32-biome sets, bytes per count vector (= 32*K)
    uint32_t R;                    // ring columns = span + kBatchCols
    uint32_t VS;                   // ring slot stride in bytes: 4*K mask + Bpad counts (+4 so that VS/4 is odd)
    float inv_total;               // 1.0f / float((2r+1)^2), SHF.cpp:495
};

// ------------------------------------------------------------------------------------------------------------------
// dictionary
// ------------------------------------------------------------------------------------------------------------------
__global__ void presence_kernel(const uint16_t* __restrict__ in, Geo g, uint32_t* __restrict__ bitmap) {
    __shared__ uint32_t bm[kDictWords];
    const uint32_t n = blockIdx.y;
    for (int i = threadIdx.x; i < kDictWords; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    const uint16_t* src = in + (size_t)n * g.in_chunk_stride;
    for (uint32_t p = blockIdx.x; p < g.PH; p += gridDim.x) {
        const uint16_t* row = src + (size_t)p * g.in_row_stride;
        for (uint32_t c = threadIdx.x; c < g.PW; c += blockDim.x) {
            const uint32_t s = row[c];
            const uint32_t bit = 1u << (s & 31u);
            if (!(((volatile uint32_t*)bm)[s >> 5] & bit)) atomicOr(&bm[s >> 5], bit);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kDictWords; i += blockDim.x)
        if (bm[i]) atomicOr(&bitmap[(size_t)n * kDictWords + i], bm[i]);
}

// one CTA of 256 threads per chunk: exclusive prefix of popcounts over the 2048 bitmap words, and the biome total
__global__ void dict_prefix_kernel(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix,
                                   uint32_t* __restrict__ n_biomes) {
    __shared__ uint32_t part[256];
    const uint32_t n = blockIdx.x, t = threadIdx.x;
    const uint32_t* bm = bitmap + (size_t)n * kDictWords;
    uint32_t local[8], sum = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        local[i] = sum;
        sum += __popc(bm[t * 8 + i]);
    }
    part[t] = sum;
    __syncthreads();
    // simple Hillis-Steele over 256 partial sums
    for (int off = 1; off < 256; off <<= 1) {
        uint32_t v = (t >= (uint32_t)off) ? part[t - off] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const uint32_t excl = part[t] - sum;
#pragma unroll
    for (int i = 0; i < 8; i++) prefix[(size_t)n * kDictWords + t * 8 + i] = excl + local[i];
    if (t == 255) n_biomes[n] = part[255];
}

// remap samples to compact ids; CTA x==0 of every chunk also writes the dictionary (compact id -> sample value)
__global__ void remap_kernel(const uint16_t* __restrict__ in, Geo g, const uint32_t* __restrict__ bitmap,
                             const uint32_t* __restrict__ prefix, uint16_t* __restrict__ cmap,
                             uint16_t* __restrict__ dict, uint32_t dict_stride) {
    __shared__ uint32_t bm[kDictWords];
    __shared__ uint32_t pf[kDictWords];
    const uint32_t n = blockIdx.y;
    for (int i = threadIdx.x; i < kDictWords; i += blockDim.x) {
        bm[i] = bitmap[(size_t)n * kDictWords + i];
        pf[i] = prefix[(size_t)n * kDictWords + i];
    }
    __syncthreads();
    const uint16_t* src = in + (size_t)n * g.in_chunk_stride;
    for (uint32_t p = blockIdx.x; p < g.PH; p += gridDim.x) {
        const uint16_t* row = src + (size_t)p * g.in_row_stride;
        uint16_t* dst = cmap + ((size_t)n * g.PH + p) * g.P;
        for (uint32_t c = threadIdx.x; c < g.PW; c += blockDim.x) {
            const uint32_t s = row[c];
            dst[c] = (uint16_t)(pf[s >> 5] + __popc(bm[s >> 5] & ((1u << (s & 31u)) - 1u)));
        }
    }
    if (blockIdx.x == 0) {
        for (int w = threadIdx.x; w < kDictWords; w += blockDim.x) {
            uint32_t bits = bm[w], at = pf[w];
            while (bits) {
                const int b = __ffs(bits) - 1;
                if (at < dict_stride) dict[(size_t)n * dict_stride + at] = (uint16_t)(w * 32 + b);
                at++;
                bits &= bits - 1u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// vertical scan: vstart map + per-tile base count vectors
// ------------------------------------------------------------------------------------------------------------------
// Per thread (= column) private shared memory: hist u8[Bpad] | last u16[Bpad] | start u16[Bpad], stride 5*Bpad+4 bytes
// (an odd number of words, so that equal offsets of neighbouring threads fall into different banks).
template <int K>
__global__ void __launch_bounds__(kVscanThreads) vscan_kernel(Geo g, const uint16_t* __restrict__ cmap,
                                                              uint16_t* __restrict__ vstart, uint8_t* __restrict__ base,
                                                              uint32_t* __restrict__ basemask) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int Bpad = 32 * K;
    constexpr int kStride = 5 * Bpad + 4;
    const uint32_t n = blockIdx.y;
    const uint32_t c = blockIdx.x * kVscanThreads + threadIdx.x;
    uint8_t* mine = smem + (size_t)threadIdx.x * kStride;
    uint8_t* hist = mine;
    uint16_t* last = reinterpret_cast<uint16_t*>(mine + Bpad);
    uint16_t* start = reinterpret_cast<uint16_t*>(mine + 3 * Bpad);
    for (int i = 0; i < Bpad; i++) {
        hist[i] = 0;
        last[i] = kNoRow;
        start[i] = 0;
    }
    if (c >= g.PW) return;
    uint32_t mask[K];
#pragma unroll
    for (int k = 0; k < K; k++) mask[k] = 0u;
    const uint16_t* col = cmap + (size_t)n * g.PH * g.P + c;
    uint16_t* vcol = vstart + (size_t)n * g.PH * g.P + c;
    const uint32_t span = g.span, two_r = 2u * g.r;
    for (uint32_t p = 0; p < g.PH; p++) {
        const uint32_t s = col[(size_t)p * g.P];
        const uint32_t l = last[s];
        if (l == kNoRow || p - l > span) start[s] = (uint16_t)p;
        last[s] = (uint16_t)p;
        vcol[(size_t)p * g.P] = start[s];
        if (++hist[s] == 1) {
#pragma unroll
            for (int k = 0; k < K; k++)
                if ((s >> 5) == (uint32_t)k) mask[k] |= 1u << (s & 31u);
        }
        if (p >= span) {
            const uint32_t o = col[(size_t)(p - span) * g.P];
            if (--hist[o] == 0) {
#pragma unroll
                for (int k = 0; k < K; k++)
                    if ((o >> 5) == (uint32_t)k) mask[k] &= ~(1u << (o & 31u));
            }
        }
        if (p >= two_r) {
            const uint32_t y = p - two_r;
            if (y < g.H && y % g.TY == 0u) {
                const uint32_t t = y / g.TY;
                const size_t slot = ((size_t)n * g.T + t) * g.PW + c;
                uint4* dst = reinterpret_cast<uint4*>(base + slot * Bpad);
                const uint32_t* hw = reinterpret_cast<const uint32_t*>(hist);
#pragma unroll
                for (int q = 0; q < 2 * K; q++) dst[q] = make_uint4(hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
#pragma unroll
                for (int k = 0; k < K; k++) basemask[slot * K + k] = mask[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// march: the horizontal sliding window with an ordered bin list per row
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <int K>
__device__ __forceinline__ bool test_bit(const uint32_t (&words)[K], uint32_t s) {
    bool hit = false;
#pragma unroll
    for (int k = 0; k < K; k++)
        if ((s >> 5) == (uint32_t)k) hit = (words[k] >> (s & 31u)) & 1u;
    return hit;
}

// A list entry is one register pair per lane and set: item = (sample value << 16) | compact id, cnt = window count.
// Entry e = k*32 + lane, valid iff e < n. The list order is the reference accumulator's bin order.
template <int K, bool EMIT>
__global__ void __launch_bounds__(512, 1)
    march_kernel(Geo g, const uint16_t* __restrict__ cmap, const uint16_t* __restrict__ vstart,
                 const uint8_t* __restrict__ base, const uint32_t* __restrict__ basemask,
                 const uint16_t* __restrict__ dict, uint32_t dict_stride, uint32_t* __restrict__ rowtotal,
                 const uint32_t* __restrict__ rowbase, const uint64_t* __restrict__ chunkbase,
                 uint2* __restrict__ bins, uint32_t* __restrict__ hso) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int Bpad = 32 * K;
    constexpr int E = 32 * K;  // list capacity
    const uint32_t TY = g.TY, R = g.R, VS = g.VS, span = g.span, two_r = 2u * g.r;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n_chunk = blockIdx.y, tile = blockIdx.x;
    const uint32_t y0 = tile * TY, y = y0 + warp;
    const bool row_active = y < g.H;

    uint8_t* ring = smem;                                                         // [TY][R][VS]
    uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + (size_t)TY * R * VS);  // [TY][2][E]
    uint16_t* sdict = reinterpret_cast<uint16_t*>(scratch + (size_t)TY * 2 * E);  // [E]
    uint32_t* sA = scratch + (size_t)warp * 2 * E;
    uint32_t* sB = sA + E;
    for (uint32_t i = threadIdx.x; i < (uint32_t)E; i += blockDim.x)
        sdict[i] = (i < dict_stride) ? dict[(size_t)n_chunk * dict_stride + i] : (uint16_t)0;

    uint8_t* myring = ring + (size_t)warp * R * VS;
    const uint16_t* cm = cmap + (size_t)n_chunk * g.PH * g.P;
    const uint16_t* vs = vstart + (size_t)n_chunk * g.PH * g.P;

    // ordered bin list of this row
    uint32_t item[K], cnt[K], listmask[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        item[k] = 0u;
        cnt[k] = 0u;
        listmask[k] = 0u;
    }
    uint32_t n = 0u;       // bins in the list
    uint32_t rowpos = 0u;  // bins emitted so far in this row
    uint32_t hso_reg = 0u;
    size_t out0 = 0;
    if (EMIT && row_active) out0 = (size_t)chunkbase[n_chunk] + rowbase[(size_t)n_chunk * g.H + y];
    const uint32_t first_row_bins = (EMIT && row_active) ? rowbase[(size_t)n_chunk * g.H + y] : 0u;
    const float inv = g.inv_total;

    uint32_t in_slot = 0u;  // ring slot of column c
    for (uint32_t cb = 0u; cb < g.PW; cb += kBatchCols) {
        __syncthreads();  // every warp is done with the slots about to be overwritten (and sdict is loaded)
        // ---- produce: vertical window counts of columns [cb, cb+32) for this warp's row (lane = column) ----
        {
            const uint32_t c = cb + lane;
            if (row_active && c < g.PW) {
                uint32_t slot_idx = in_slot + lane;
                if (slot_idx >= R) slot_idx -= R;
                uint8_t* slot = myring + (size_t)slot_idx * VS;
                uint32_t* slot_w = reinterpret_cast<uint32_t*>(slot);
                uint8_t* sc = slot + 4 * K;
                const size_t bslot = ((size_t)n_chunk * g.T + tile) * g.PW + c;
                const uint4* bsrc = reinterpret_cast<const uint4*>(base + bslot * Bpad);
#pragma unroll
                for (int q = 0; q < 2 * K; q++) {
                    const uint4 v = bsrc[q];
                    slot_w[K + 4 * q + 0] = v.x;
                    slot_w[K + 4 * q + 1] = v.y;
                    slot_w[K + 4 * q + 2] = v.z;
                    slot_w[K + 4 * q + 3] = v.w;
                }
                uint32_t mask[K];
#pragma unroll
                for (int k = 0; k < K; k++) mask[k] = basemask[bslot * K + k];
                // slide the vertical window down from the tile's first row to this warp's row
                for (uint32_t i0 = 0u; i0 < warp; i0 += 8u) {
                    uint32_t sin[8], sout[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint32_t i = i0 + j;
                        if (i < warp) {
                            sin[j] = cm[(size_t)(y0 + i + span) * g.P + c];
                            sout[j] = cm[(size_t)(y0 + i) * g.P + c];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint32_t i = i0 + j;
                        if (i < warp) {
                            const uint32_t a = sin[j], o = sout[j];
                            if (++sc[a] == 1) {
#pragma unroll
                                for (int k = 0; k < K; k++)
                                    if ((a >> 5) == (uint32_t)k) mask[k] |= 1u << (a & 31u);
                            }
                            if (--sc[o] == 0) {
#pragma unroll
                                for (int k = 0; k < K; k++)
                                    if ((o >> 5) == (uint32_t)k) mask[k] &= ~(1u << (o & 31u));
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < K; k++) slot_w[k] = mask[k];
            }
        }
        __syncthreads();
        // ---- consume: advance this row's window over the batch ----
        if (row_active) {
            const uint32_t c_end = min(cb + (uint32_t)kBatchCols, g.PW);
            for (uint32_t c = cb; c < c_end; c++) {
                const uint8_t* sin = myring + (size_t)in_slot * VS;
                const uint32_t* sin_w = reinterpret_cast<const uint32_t*>(sin);
                const uint8_t* sin_c = sin + 4 * K;
                uint32_t out_slot = in_slot + R - span;  // column c - span
                if (out_slot >= R) out_slot -= R;
                const uint8_t* sout_c = myring + (size_t)out_slot * VS + 4 * K;
                const bool has_out = c >= span;
                bool event = false;
                uint32_t born[K];
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const uint32_t id = item[k] & 0xFFFFu;
                    uint32_t v = cnt[k] + sin_c[id];
                    if (has_out) v -= sout_c[id];
                    cnt[k] = v;
                    born[k] = sin_w[k] & ~listmask[k];
                    const bool dead = (k * 32 + lane < n) && v == 0u;
                    event |= (born[k] != 0u) | (__ballot_sync(kFull, dead) != 0u);
                }
                if (event) {
                    // ---------------- slow path: bins die and/or are born ----------------
                    // (1) drop dead bins, keeping the order of the survivors (SHF.cpp:435-445)
                    {
                        uint32_t keep_base = 0u;
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const bool alive = (k * 32 + lane < n) && cnt[k] != 0u;
                            const unsigned am = __ballot_sync(kFull, alive);
                            if (alive) {
                                const uint32_t idx = keep_base + __popc(am & lanemask_lt());
                                sA[idx] = item[k];
                                sB[idx] = cnt[k];
                            }
                            keep_base += __popc(am);
                        }
                        __syncwarp();
                        if (keep_base != n) {
                            n = keep_base;
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                const uint32_t e = k * 32 + lane;
                                if (e < n) {
                                    item[k] = sA[e];
                                    cnt[k] = sB[e];
                                }
                            }
                        }
                        __syncwarp();
                    }
                    // (2) append the bins born in column c, ordered by their vertical chain start (SHF.cpp:411-415:
                    //     the column's own bin order is the order the horizontal pass inserts them in)
                    uint32_t nb = 0u;
#pragma unroll
                    for (int k = 0; k < K; k++) nb += __popc(born[k]);
                    if (nb) {
                        uint32_t pending[K];
#pragma unroll
                        for (int k = 0; k < K; k++) pending[k] = born[k];
                        uint32_t found = 0u;
                        // walk the window rows of column c bottom-up, 32 rows at a time: the lowest occurrence of a
                        // biome carries its chain start in vstart
                        for (uint32_t blk = 0u; blk * 32u < span && found < nb; blk++) {
                            const int32_t off = (int32_t)two_r - (int32_t)(blk * 32u + lane);
                            const bool valid = off >= 0;
                            const size_t at = (size_t)(y + (valid ? off : 0)) * g.P + c;
                            const uint32_t s = valid ? (uint32_t)cm[at] : 0xFFFFu;
                            const bool inpend = valid && test_bit<K>(pending, s);
                            const unsigned same = __match_any_sync(kFull, s);
                            const bool first = inpend && ((uint32_t)(__ffs(same) - 1) == lane);
                            const unsigned fm = __ballot_sync(kFull, first);
                            if (first) {
                                const uint32_t idx = found + __popc(fm & lanemask_lt());
                                sA[idx] = ((uint32_t)sdict[s] << 16) | s;
                                sB[idx] = ((uint32_t)vs[at] << 16) | (uint32_t)sin_c[s];
                            }
                            found += __popc(fm);
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                const uint32_t mine = (first && (s >> 5) == (uint32_t)k) ? (1u << (s & 31u)) : 0u;
                                pending[k] &= ~__reduce_or_sync(kFull, mine);
                            }
                        }
                        __syncwarp();
                        // rank by chain start row (unique per biome within a column)
                        uint32_t ra[K], rb[K], rank[K];
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const uint32_t idx = k * 32 + lane;
                            rank[k] = idx;
                            if (idx < nb) {
                                ra[k] = sA[idx];
                                rb[k] = sB[idx];
                                if (EMIT) {
                                    uint32_t rk = 0u;
                                    for (uint32_t i = 0u; i < nb; i++) rk += (sB[i] >> 16) < (rb[k] >> 16);
                                    rank[k] = rk;
                                }
                            }
                        }
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            if ((uint32_t)(k * 32) + lane < nb) {
                                sA[rank[k]] = ra[k];
                                sB[rank[k]] = rb[k];
                            }
                        }
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const uint32_t e = k * 32 + lane;
                            if (e >= n && e < n + nb) {
                                item[k] = sA[e - n];
                                cnt[k] = sB[e - n] & 0xFFFFu;
                            }
                        }
                        n += nb;
                        __syncwarp();
                    }
                    // (3) membership mask of the list
#pragma unroll
                    for (int kk = 0; kk < K; kk++) {
                        uint32_t mine = 0u;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const uint32_t id = item[k] & 0xFFFFu;
                            if ((uint32_t)(k * 32) + lane < n && (id >> 5) == (uint32_t)kk) mine |= 1u << (id & 31u);
                        }
                        listmask[kk] = __reduce_or_sync(kFull, mine);
                    }
                }
                // ---------------- emit pixel x = c - 2r ----------------
                if (c >= two_r) {
                    const uint32_t x = c - two_r;
                    if (EMIT) {
                        uint2* dst = bins + out0 + rowpos;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const uint32_t e = k * 32 + lane;
                            if (e < n) {
                                const float w = __fmul_rn(__uint2float_rn(cnt[k]), inv);
                                dst[e] = make_uint2(item[k] >> 16, __float_as_uint(w));
                            }
                        }
                        if ((x & 31u) == lane) hso_reg = first_row_bins + rowpos;
                        if ((x & 31u) == 31u || x == g.W - 1u) {
                            if (lane <= (x & 31u))
                                hso[(size_t)n_chunk * ((size_t)g.W * g.H + 1u) + (size_t)y * g.W + (x & ~31u) + lane] =
                                    hso_reg;
                        }
                    }
                    rowpos += n;
                }
                in_slot = (in_slot + 1u == R) ? 0u : in_slot + 1u;
            }
        } else {
            in_slot += kBatchCols;
            if (in_slot >= R) in_slot -= R;
        }
    }
    if (!EMIT && row_active && lane == 0u) rowtotal[(size_t)n_chunk * g.H + y] = rowpos;
}

// ------------------------------------------------------------------------------------------------------------------
// first-bin index of every row (exclusive scan of the bins-per-row), chunk totals, the closing offset entry
// ------------------------------------------------------------------------------------------------------------------
__global__ void rowscan_kernel(Geo g, const uint32_t* __restrict__ rowtotal, uint32_t* __restrict__ rowbase,
                               unsigned long long* __restrict__ chunktotal, uint32_t* __restrict__ hso) {
    __shared__ unsigned long long part[1024];
    __shared__ unsigned long long carry;
    const uint32_t n = blockIdx.x, t = threadIdx.x;
    if (t == 0) carry = 0ull;
    __syncthreads();
    for (uint32_t y0 = 0u; y0 < g.H; y0 += blockDim.x) {
        const uint32_t y = y0 + t;
        const unsigned long long v = (y < g.H) ? rowtotal[(size_t)n * g.H + y] : 0ull;
        part[t] = v;
        __syncthreads();
        for (uint32_t off = 1u; off < blockDim.x; off <<= 1) {
            const unsigned long long add = (t >= off) ? part[t - off] : 0ull;
            __syncthreads();
            part[t] += add;
            __syncthreads();
        }
        const unsigned long long excl = carry + part[t] - v;
        if (y < g.H) rowbase[(size_t)n * g.H + y] = (uint32_t)excl;  // garbage if the total overflows; host checks
        __syncthreads();
        if (t == blockDim.x - 1) carry += part[t];
        __syncthreads();
    }
    if (t == 0) {
        chunktotal[n] = carry;
        hso[(size_t)n * ((size_t)g.W * g.H + 1u) + (size_t)g.W * g.H] = (uint32_t)carry;
    }
}

}  // namespace shf
