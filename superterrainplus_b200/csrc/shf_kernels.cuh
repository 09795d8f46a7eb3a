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
    uint32_t R;                    // ring columns = span + 16 * stages
    uint32_t stages;               // batches the march producers may run ahead of the consumers (2..6)
    uint32_t producers;            // producer warps of the march kernel (1..4)
    uint32_t VS;                   // count vector stride in the ring (= Bpad)
    float inv_total;               // 1.0f / float((2r+1)^2), SHF.cpp:495
    uint32_t flags;                // bit 0: do not use the packed short-list path (A/B measurements)
    unsigned long long* dbg;       // measurement only: cycle counters of the march kernel, or null
};

// ------------------------------------------------------------------------------------------------------------------
// dictionary
// ------------------------------------------------------------------------------------------------------------------
// V = samples per load: 8 (one 16-byte load; needs a 16-byte aligned view and row stride) or 1
template <int V>
__global__ void presence_kernel(const uint16_t* __restrict__ in, Geo g, uint32_t* __restrict__ bitmap) {
    __shared__ uint32_t bm[kDictWords];
    const uint32_t n = blockIdx.y;
    for (int i = threadIdx.x; i < kDictWords; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    const uint16_t* src = in + (size_t)n * g.in_chunk_stride;
    auto mark = [&](uint32_t s) {
        const uint32_t bit = 1u << (s & 31u);
        if (!(((volatile uint32_t*)bm)[s >> 5] & bit)) atomicOr(&bm[s >> 5], bit);
    };
    for (uint32_t p = blockIdx.x; p < g.PH; p += gridDim.x) {
        const uint16_t* row = src + (size_t)p * g.in_row_stride;
        if (V == 8) {
            const uint32_t nv = g.PW / 8u;
            for (uint32_t v = threadIdx.x; v < nv; v += blockDim.x) {
                const uint4 q = reinterpret_cast<const uint4*>(row)[v];
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    mark(w[t] & 0xFFFFu);
                    if ((w[t] >> 16) != (w[t] & 0xFFFFu)) mark(w[t] >> 16);
                }
            }
            for (uint32_t c = nv * 8u + threadIdx.x; c < g.PW; c += blockDim.x) mark(row[c]);
        } else {
            for (uint32_t c = threadIdx.x; c < g.PW; c += blockDim.x) mark(row[c]);
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
template <int V>
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
    auto rank = [&](uint32_t s) { return pf[s >> 5] + __popc(bm[s >> 5] & ((1u << (s & 31u)) - 1u)); };
    const uint16_t* src = in + (size_t)n * g.in_chunk_stride;
    for (uint32_t p = blockIdx.x; p < g.PH; p += gridDim.x) {
        const uint16_t* row = src + (size_t)p * g.in_row_stride;
        uint16_t* dst = cmap + ((size_t)n * g.PH + p) * g.P;
        if (V == 8) {
            const uint32_t nv = g.PW / 8u;
            for (uint32_t v = threadIdx.x; v < nv; v += blockDim.x) {
                const uint4 q = reinterpret_cast<const uint4*>(row)[v];
                uint4 o;
                o.x = rank(q.x & 0xFFFFu) | (rank(q.x >> 16) << 16);
                o.y = rank(q.y & 0xFFFFu) | (rank(q.y >> 16) << 16);
                o.z = rank(q.z & 0xFFFFu) | (rank(q.z >> 16) << 16);
                o.w = rank(q.w & 0xFFFFu) | (rank(q.w >> 16) << 16);
                reinterpret_cast<uint4*>(dst)[v] = o;
            }
            for (uint32_t c = nv * 8u + threadIdx.x; c < g.PW; c += blockDim.x) dst[c] = (uint16_t)rank(row[c]);
        } else {
            for (uint32_t c = threadIdx.x; c < g.PW; c += blockDim.x) dst[c] = (uint16_t)rank(row[c]);
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
// vertical scan: vstart map, per-tile base count vectors, per-(row, column) presence masks
// ------------------------------------------------------------------------------------------------------------------
// Per thread (= column) private shared memory: hist u8[Bpad] | last u16[Bpad] | start u16[Bpad], stride 5*Bpad+4 bytes
// (an odd number of words, so that equal offsets of neighbouring threads fall into different banks).
// colmask(n, y, c) = K words, bit b of word k set iff compact id 32k+b occurs in column c, rows [y, y+2r].
template <int K>
__global__ void __launch_bounds__(kVscanThreads) vscan_kernel(Geo g, const uint16_t* __restrict__ cmap,
                                                              uint16_t* __restrict__ vstart, uint8_t* __restrict__ base,
                                                              uint32_t* __restrict__ colmask) {
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
    uint32_t* mcol = colmask + ((size_t)n * g.H * g.PW + c) * K;
    const uint32_t span = g.span, two_r = 2u * g.r;
    uint32_t next_tile_row = 0u, tile = 0u;
    for (uint32_t p0 = 0; p0 < g.PH; p0 += 8u) {
        // the loads of 8 rows go out together; the per-row work below is a chain of shared-memory updates
        uint32_t s_in[8], s_out[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t p = p0 + j;
            s_in[j] = p < g.PH ? (uint32_t)col[(size_t)p * g.P] : 0u;
            s_out[j] = (p < g.PH && p >= span) ? (uint32_t)col[(size_t)(p - span) * g.P] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t p = p0 + j;
            if (p >= g.PH) break;
            const uint32_t s = s_in[j];
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
                const uint32_t o = s_out[j];
                if (--hist[o] == 0) {
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if ((o >> 5) == (uint32_t)k) mask[k] &= ~(1u << (o & 31u));
                }
            }
            if (p >= two_r) {
                const uint32_t y = p - two_r;  // the window [y, y+2r] is complete
#pragma unroll
                for (int k = 0; k < K; k++) mcol[(size_t)y * g.PW * K + k] = mask[k];
                if (y == next_tile_row) {
                    const size_t slot = ((size_t)n * g.T + tile) * g.PW + c;
                    uint4* dst = reinterpret_cast<uint4*>(base + slot * Bpad);
                    const uint32_t* hw = reinterpret_cast<const uint32_t*>(hist);
#pragma unroll
                    for (int q = 0; q < 2 * K; q++)
                        dst[q] = make_uint4(hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
                    tile++;
                    next_tile_row += g.TY;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// bins per row without building any histogram: a pixel's bin count is the popcount of the OR of its window's column
// masks (the count is order-free, so this is exact by construction). The sliding OR of width span = 2r+1 is evaluated
// the van Herk / Gil-Werman way: columns are cut into blocks of `span`; with suf(i) = OR of block start..end from i on
// and pre(i) = OR of the block up to i, the window [x, x+span-1] is suf(x) | pre(x+span-1). A lane owns one chain
// (row, block, mask word): it builds the suffix of its block backwards, then walks forwards combining it with the
// running prefix of the next block. Two shared-memory ops per mask word and pass instead of a log-step tree.
// ------------------------------------------------------------------------------------------------------------------
struct RowcountPlan {
    uint32_t nblk;      // blocks of `span` columns per row
    uint32_t rows;      // rows one warp handles at a time
    uint32_t stride;    // words between two rows of a warp's staging area (padded against bank conflicts)
};

template <int K>
__global__ void __launch_bounds__(64) rowcount_kernel(Geo g, RowcountPlan rp, const uint32_t* __restrict__ colmask,
                                                      uint32_t* __restrict__ rowtotal) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, n_warps = blockDim.x >> 5;
    const uint32_t n = blockIdx.y;
    const uint32_t PW = g.PW, NW = PW * K, span = g.span;
    const uint32_t y0 = (blockIdx.x * n_warps + warp) * rp.rows;
    if (y0 >= g.H) return;
    const uint32_t rows = min(rp.rows, g.H - y0);
    uint32_t* org = reinterpret_cast<uint32_t*>(smem) + (size_t)warp * (2u * rp.rows * rp.stride + 32u);
    uint32_t* suf = org + (size_t)rp.rows * rp.stride;
    uint32_t* tot = suf + (size_t)rp.rows * rp.stride;  // [rows] bins per row
    for (uint32_t r = 0u; r < rows; r++) {
        const uint32_t* src = colmask + ((size_t)n * g.H + y0 + r) * NW;
        for (uint32_t i = lane; i < NW; i += 32u) org[r * rp.stride + i] = src[i];
    }
    if (lane < rp.rows) tot[lane] = 0u;
    __syncwarp();
    const uint32_t chains = rp.nblk * K;
    for (uint32_t q = lane; q < rows * chains; q += 32u) {  // suffix ORs, written next to the masks
        const uint32_t r = q / chains, ch = q - r * chains, blk = ch / K, k = ch - blk * K;
        const uint32_t lo = blk * span, hi = min(lo + span, PW);
        const uint32_t* a = org + r * rp.stride + k;
        uint32_t* s2 = suf + r * rp.stride + k;
        uint32_t acc = 0u;
        for (uint32_t i = hi; i > lo;) {  // 8 loads in flight per trip: the chain itself is only register ORs
            const uint32_t m = min(8u, i - lo);
            uint32_t v[8];
#pragma unroll
            for (uint32_t j = 0u; j < 8u; j++) v[j] = j < m ? a[(i - 1u - j) * K] : 0u;
#pragma unroll
            for (uint32_t j = 0u; j < 8u; j++) {
                acc |= v[j];
                if (j < m) s2[(i - 1u - j) * K] = acc;
            }
            i -= m;
        }
    }
    __syncwarp();
    for (uint32_t q = lane; q < rows * chains; q += 32u) {  // windows starting in my block
        const uint32_t r = q / chains, ch = q - r * chains, blk = ch / K, k = ch - blk * K;
        const uint32_t lo = blk * span, hi = min(lo + span, g.W);  // only windows of real pixels
        const uint32_t* nxt = org + r * rp.stride + (size_t)(lo + span) * K + k;  // next block's masks
        const uint32_t* s2 = suf + r * rp.stride + k;
        uint32_t pre = 0u, total = 0u;
        for (uint32_t x = lo; x < hi;) {
            const uint32_t m = min(8u, hi - x);
            uint32_t sv[8], nv[8];
#pragma unroll
            for (uint32_t j = 0u; j < 8u; j++) {
                sv[j] = j < m ? s2[(x + j) * K] : 0u;
                nv[j] = j < m ? nxt[(x - lo + j) * K] : 0u;  // column x + j + span joins the prefix of the next window
            }
#pragma unroll
            for (uint32_t j = 0u; j < 8u; j++) {
                if (j < m) total += __popc(sv[j] | pre);
                pre |= nv[j];
            }
            x += m;
        }
        if (hi > lo) atomicAdd(&tot[r], total);
    }
    __syncwarp();
    if (lane < rows) rowtotal[(size_t)n * g.H + y0 + lane] = tot[lane];
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

// shared-memory mbarriers (arrive = release, wait = acquire at CTA scope): the producer -> consumer hand-over must not
// make the consumer warps wait for each other, which a bar.sync would
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// PTX shl clamps shift amounts above 31 to 32, i.e. the result is 0 (C++ leaves such shifts undefined)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t sh) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(sh));
    return r;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// Shared memory of a march CTA (TY consumer warps = TY rows, plus one producer warp):
//   cring[TY][R][32K]   u8   vertical window counts of the last R = 2r+1 + NB*stages columns, per row, by compact id
//   mbuf[TY][NB][K]     u32  presence masks of the batch being consumed, per row
//   scratch[TY][2][32K] u32  per-warp staging for list compaction / birth ordering
//   sdict[32K]          u16  compact id -> sample value
// The producer warp runs `stages` batches of NB = 16 columns ahead of the consumers. For a column it keeps the count
// vector in registers (2K lanes x 16 bytes), starts from the tile's base vector, and walks down the tile's rows: store
// the row's slot (one 128-bit store per lane, neighbouring columns fill one 128-byte line), then apply the one sample
// entering and the one leaving the vertical window. Hand-over uses mbarriers: full[s] producer -> consumers (count 1),
// empty[s] consumers -> producers (count TY), s = batch % stages; a consumer warp never waits for another consumer.
//
// A consumer warp keeps its row's ordered bin list in registers: entry e = k*32 + lane (valid iff e < n) is the triple
// id (compact id) / hi (sample value) / cnt (window count). A step adds column c, removes column c - (2r+1), handles
// bin deaths/births if there are any (slow path), and emits pixel x = c - 2r. In the steady state four steps are done
// at once when none of them has a death or birth.
constexpr int kMarchNB = 16;

template <int K>
__global__ void __launch_bounds__(640, 1)
    march_kernel(Geo g, const uint16_t* __restrict__ cmap, const uint16_t* __restrict__ vstart,
                 const uint8_t* __restrict__ base, const uint32_t* __restrict__ colmask,
                 const uint16_t* __restrict__ dict, uint32_t dict_stride, const uint32_t* __restrict__ rowbase,
                 const uint64_t* __restrict__ chunkbase, uint2* __restrict__ bins, uint32_t* __restrict__ hso) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int CS = 32 * K;  // bytes per count vector
    constexpr int E = 32 * K;   // list capacity
    constexpr int NB = kMarchNB;
    const uint32_t TY = g.TY, R = g.R, span = g.span, two_r = 2u * g.r, PW = g.PW, stages = g.stages;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n_chunk = blockIdx.y, tile = blockIdx.x;
    const uint32_t y0 = tile * TY;
    const uint32_t NP = g.producers;                  // producer warps; producer p owns batches b = p (mod NP)
    const uint32_t n_batches = (PW + NB - 1u) / NB;

    uint8_t* cring = smem;
    uint32_t* mbuf_all = reinterpret_cast<uint32_t*>(smem + (size_t)TY * R * CS);
    uint32_t* scratch = mbuf_all + (size_t)TY * NB * K;
    uint16_t* sdict = reinterpret_cast<uint16_t*>(scratch + (size_t)TY * 2 * E);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sdict + E);  // [stages]
    uint64_t* empty_bar = full_bar + 8;                            // [stages]
    for (uint32_t i = threadIdx.x; i < (uint32_t)E; i += blockDim.x)
        sdict[i] = (i < dict_stride) ? dict[(size_t)n_chunk * dict_stride + i] : (uint16_t)0;
    if (threadIdx.x < stages) {
        mbar_init(&full_bar[threadIdx.x], (uint32_t)(NB / (32 / (2 * K))));  // one arrival per pass of the batch
        mbar_init(&empty_bar[threadIdx.x], TY);
    }
    __syncthreads();
    const uint16_t* cm = cmap + (size_t)n_chunk * g.PH * g.P;

    if (warp >= TY) {
        // =============================== producer warps ===============================
        // work item t = (batch b, pass q), t = b * PPB + q; producer p owns the items t = p (mod NP)
        constexpr int LPC = 2 * K;       // lanes per column, 16 bytes each
        constexpr int CPP = 32 / LPC;    // columns per pass
        constexpr int PPB = NB / CPP;    // passes per batch
        const uint32_t part = lane % LPC, colq = lane / LPC;
        const uint32_t tile_rows = min(TY, g.H - y0);
        const uint32_t bit0 = part * 128u;  // bit position of my first counter inside the column's count vector
        for (uint32_t t = warp - TY; t < n_batches * PPB; t += NP) {
            const uint32_t b = t / PPB, pass = t % PPB;
            const uint32_t s = b % stages, cb = b * NB;
            const long long tp0 = g.dbg ? clock64() : 0;
            const uint32_t slot0 = cb % R;  // ring slot of the batch's first column
            const uint32_t cu = pass * CPP + colq;
            const bool live = cb + cu < PW;
            const uint32_t c = live ? cb + cu : PW - 1u;
            uint32_t slot = slot0 + cu;
            if (slot >= R) slot -= R;
            uint4 v = *reinterpret_cast<const uint4*>(base + (((size_t)n_chunk * g.T + tile) * PW + c) * CS + part * 16u);
            // samples entering / leaving the vertical window when it moves from tile row i to i+1
            uint32_t sa[16], so[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t pi = min(y0 + i + span, g.PH - 1u), po = min(y0 + i, g.PH - 1u);
                sa[i] = cm[(size_t)pi * g.P + c];
                so[i] = cm[(size_t)po * g.P + c];
            }
            if (b >= stages) mbar_wait(&empty_bar[s], (b / stages - 1u) & 1u);  // batch b - stages is consumed
            const long long tp1 = g.dbg ? clock64() : 0;
            uint8_t* out = cring + (size_t)slot * CS + part * 16u;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if ((uint32_t)i < tile_rows) {
                    if (live) *reinterpret_cast<uint4*>(out) = v;
                    out += (size_t)R * CS;
                    // counter of compact id a sits at bit 8a of the column vector; a shift by >= 32 (or "negative",
                    // i.e. huge) yields 0 with PTX shl, so every word only sees its own counters
                    const uint32_t ba = sa[i] * 8u - bit0, bo = so[i] * 8u - bit0;
                    v.x += shl_clamp(1u, ba) - shl_clamp(1u, bo);
                    v.y += shl_clamp(1u, ba - 32u) - shl_clamp(1u, bo - 32u);
                    v.z += shl_clamp(1u, ba - 64u) - shl_clamp(1u, bo - 64u);
                    v.w += shl_clamp(1u, ba - 96u) - shl_clamp(1u, bo - 96u);
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0u) mbar_arrive(&full_bar[s]);
            if (g.dbg && lane == 0u) {
                const long long tp2 = clock64();
                atomicAdd(&g.dbg[2], (unsigned long long)(tp2 - tp0));
                atomicAdd(&g.dbg[3], (unsigned long long)(tp1 - tp0));
                atomicAdd(&g.dbg[4], (unsigned long long)(tp2 - tp1));
            }
        }
        return;
    }

    // =============================== consumer warps ===============================
    const uint32_t y = y0 + warp;
    const bool row_active = y < g.H;
    const long long tw0 = g.dbg ? clock64() : 0;
    uint8_t* crow = cring + (size_t)warp * R * CS;
    uint32_t* mbuf = mbuf_all + (size_t)warp * NB * K;
    uint32_t* sA = scratch + (size_t)warp * 2 * E;
    uint32_t* sB = sA + E;
    const uint16_t* vs = vstart + (size_t)n_chunk * g.PH * g.P;
    const uint32_t* cmask_row = colmask + ((size_t)n_chunk * g.H + (row_active ? y : 0u)) * PW * K;

    uint32_t id[K], hi[K], cnt[K], listmask[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        id[k] = 0u;
        hi[k] = 0u;
        cnt[k] = 0u;
        listmask[k] = 0u;
    }
    uint32_t n = 0u;       // bins in the list
    // packed view of a short list (n <= 16): the lanes >= n of set 0 replicate entry lane % n, so that S = 32 / n
    // consecutive pixels are produced by one pass over the warp (lane = step slot * n + entry)
    uint32_t pk_steps = 0u, pk_entry = 0u, pk_slot = 0u, pk_off = 0u;  // pk_off: my entry's byte in the slot's count vector
    uint32_t rowpos = 0u;  // bins emitted so far in this row
    const uint32_t row_first = row_active ? rowbase[(size_t)n_chunk * g.H + y] : 0u;
    uint2* dst = bins + (row_active ? (size_t)chunkbase[n_chunk] + row_first : (size_t)0) + lane;
    uint32_t* hso_row = hso + (size_t)n_chunk * ((size_t)g.W * g.H + 1u) + (size_t)(row_active ? y : 0u) * g.W;
    const float inv = g.inv_total;

    // presence masks of the next batch, one column per lane, fetched a batch ahead
    uint32_t mnext[K];
#pragma unroll
    for (int k = 0; k < K; k++) mnext[k] = (row_active && lane < (uint32_t)NB && lane < PW) ? cmask_row[(size_t)lane * K + k] : 0u;

    uint32_t in_slot = 0u;                       // ring slot of column c
    uint32_t out_slot = (R - span % R) % R;      // ring slot of column c - span (meaningful once c >= span)
    for (uint32_t b = 0u; b < n_batches; b++) {
        const uint32_t s = b % stages, cb = b * NB;
        const uint32_t ce = min(cb + (uint32_t)NB, PW);
        const long long tc0 = g.dbg ? clock64() : 0;
        mbar_wait(&full_bar[s], (b / stages) & 1u);
        if (g.dbg && lane == 0u) atomicAdd(&g.dbg[1], (unsigned long long)(clock64() - tc0));
        if (row_active) {
            if (lane < (uint32_t)NB) {
#pragma unroll
                for (int k = 0; k < K; k++) mbuf[lane * K + k] = mnext[k];
            }
            // bit i: column cb + i holds a biome that is not in the list (a birth is due there); kept current after
            // every change of the list, so the fast paths test a register instead of re-reading the masks
            uint32_t born_cols;
            {
                uint32_t w = 0u;
#pragma unroll
                for (int k = 0; k < K; k++) w |= mnext[k] & ~listmask[k];
                born_cols = __ballot_sync(kFull, w != 0u && lane < (uint32_t)NB);
            }
            if (lane < (uint32_t)NB) {
                const uint32_t cn = cb + NB + lane;
#pragma unroll
                for (int k = 0; k < K; k++) mnext[k] = cn < PW ? cmask_row[(size_t)cn * K + k] : 0u;
            }
            __syncwarp();
            uint32_t c = cb;
            while (c < ce) {
                // a segment: no ring wrap, constant has_out / emit
                uint32_t seg_end = min(ce, c + (R - in_slot));
                const bool has_out = c >= span, emit = c >= two_r;
                seg_end = has_out ? min(seg_end, c + (R - out_slot)) : min(seg_end, span);
                if (!emit) seg_end = min(seg_end, two_r);
                const uint8_t* pin = crow + (size_t)in_slot * CS;
                const uint8_t* pout = crow + (size_t)out_slot * CS;
                uint32_t steps = seg_end - c;
                uint32_t rel = c - cb;   // column of the next step inside the batch
                uint32_t x = c - two_r;  // only meaningful when emit
                while (steps > 0u) {
                    const uint32_t births = born_cols >> rel;
                    if (pk_steps >= 2u && steps >= 2u && has_out && emit) {
                        // ---------------- packed fast path: min(S, steps) pixels of a short list at once ----------------
                        const uint32_t sn = min(pk_steps, steps);
                        if ((births & ((1u << sn) - 1u)) == 0u) {
                            const bool act = pk_slot < sn;
                            int32_t d = 0;
                            if (act) d = (int32_t)pin[pk_off] - (int32_t)pout[pk_off];
                            for (uint32_t off = 1u, sh = n; off < sn; off <<= 1, sh <<= 1) {  // running sum over my entry's slots
                                const int32_t t = __shfl_up_sync(kFull, d, sh);
                                if (pk_slot >= off) d += t;
                            }
                            const uint32_t cj = cnt[0] + (uint32_t)d;  // count of my entry after step pk_slot
                            if (!__any_sync(kFull, act && cj == 0u)) {
                                if (act) *dst = make_uint2(hi[0], __float_as_uint(__fmul_rn(__uint2float_rn(cj), inv)));
                                if (lane < sn) hso_row[x + lane] = row_first + rowpos + lane * n;
                                cnt[0] = __shfl_sync(kFull, cj, (sn - 1u) * n + pk_entry);
                                dst += sn * n;
                                rowpos += sn * n;
                                x += sn;
                                rel += sn;
                                steps -= sn;
                                pin += sn * CS;
                                pout += sn * CS;
                                continue;
                            }
                        }
                    } else if (steps >= 4u && has_out && emit) {
                        // ---------------- fast path: four steps without any birth or death ----------------
                        if ((births & 0xFu) == 0u) {
                            uint32_t c1[K], c2[K], c3[K], c4[K];
                            bool bad = false;
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                const uint8_t* gi = pin + id[k];
                                const uint8_t* go = pout + id[k];
                                c1[k] = cnt[k] + gi[0] - go[0];
                                c2[k] = c1[k] + gi[CS] - go[CS];
                                c3[k] = c2[k] + gi[2 * CS] - go[2 * CS];
                                c4[k] = c3[k] + gi[3 * CS] - go[3 * CS];
                                bad |= ((uint32_t)(k * 32) + lane < n) && min(min(c1[k], c2[k]), min(c3[k], c4[k])) == 0u;
                            }
                            if (!__any_sync(kFull, bad)) {
#pragma unroll
                                for (int k = 0; k < K; k++) {
                                    if ((uint32_t)(k * 32) + lane < n) {
                                        uint2* d = dst + k * 32;
                                        d[0] = make_uint2(hi[k], __float_as_uint(__fmul_rn(__uint2float_rn(c1[k]), inv)));
                                        d[n] = make_uint2(hi[k], __float_as_uint(__fmul_rn(__uint2float_rn(c2[k]), inv)));
                                        d[2u * n] = make_uint2(hi[k], __float_as_uint(__fmul_rn(__uint2float_rn(c3[k]), inv)));
                                        d[3u * n] = make_uint2(hi[k], __float_as_uint(__fmul_rn(__uint2float_rn(c4[k]), inv)));
                                    }
                                    cnt[k] = c4[k];
                                }
                                if (lane < 4u) hso_row[x + lane] = row_first + rowpos + lane * n;
                                dst += 4u * n;
                                rowpos += 4u * n;
                                x += 4u;
                                rel += 4u;
                                steps -= 4u;
                                pin += 4 * CS;
                                pout += 4 * CS;
                                continue;
                            }
                        }
                    } else if (steps >= 4u && !has_out && !emit) {
                        // ---------------- fast path while the window fills: counts only grow, no pixel yet ----------------
                        if ((births & 0xFu) == 0u) {
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                const uint8_t* gi = pin + id[k];
                                cnt[k] += (uint32_t)gi[0] + gi[CS] + gi[2 * CS] + gi[3 * CS];
                            }
                            rel += 4u;
                            steps -= 4u;
                            pin += 4 * CS;
                            pout += 4 * CS;
                            continue;
                        }
                    }
                    // ---------------- one step, any case ----------------
                    {
                        bool deadp = false;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            uint32_t v = cnt[k] + pin[id[k]];
                            if (has_out) v -= pout[id[k]];
                            cnt[k] = v;
                            deadp |= ((uint32_t)(k * 32) + lane < n) && v == 0u;
                        }
                        const bool has_birth = (births & 1u) != 0u;
                        if (__any_sync(kFull, deadp) || has_birth) {
                            const long long ts0 = g.dbg ? clock64() : 0;
                            const uint32_t c_now = cb + rel;
                            // (1) drop dead bins, keeping the order of the survivors (SHF.cpp:435-445)
                            {
                                uint32_t keep_base = 0u;
                                __syncwarp();
#pragma unroll
                                for (int k = 0; k < K; k++) {
                                    const bool alive = ((uint32_t)(k * 32) + lane < n) && cnt[k] != 0u;
                                    const unsigned am = __ballot_sync(kFull, alive);
                                    if (alive) {
                                        const uint32_t idx = keep_base + __popc(am & lanemask_lt());
                                        sA[idx] = (hi[k] << 16) | id[k];
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
                                            const uint32_t packed = sA[e];
                                            id[k] = packed & 0xFFFFu;
                                            hi[k] = packed >> 16;
                                            cnt[k] = sB[e];
                                        }
                                    }
                                }
                                __syncwarp();
                            }
                            // (2) append the bins born in this column, ordered by their vertical chain start
                            //     (SHF.cpp:411-415: the column's own bin order is the horizontal pass's insertion order)
                            if (has_birth) {
                                uint32_t pending[K];
                                uint32_t nb = 0u;
#pragma unroll
                                for (int k = 0; k < K; k++) {
                                    pending[k] = mbuf[rel * K + k] & ~listmask[k];
                                    nb += __popc(pending[k]);
                                }
                                // the window rows of the column, bottom-up, 32 rows per block: every occurrence of a
                                // biome inside one window carries the same chain start in vstart. All loads go out
                                // together (2r+1 <= 255 here, so at most 8 blocks).
                                uint32_t cell[8];
#pragma unroll
                                for (int blk = 0; blk < 8; blk++) {
                                    const int32_t off = (int32_t)two_r - (int32_t)(blk * 32 + lane);
                                    cell[blk] = 0xFFFFu;
                                    if (off >= 0) {
                                        const size_t at = (size_t)(y + off) * g.P + c_now;
                                        cell[blk] = ((uint32_t)vs[at] << 16) | (uint32_t)cm[at];
                                    }
                                }
                                uint32_t found = 0u;
#pragma unroll
                                for (int blk = 0; blk < 8; blk++) {
                                    if ((uint32_t)(blk * 32) < span && found < nb) {
                                        const uint32_t sv = cell[blk] & 0xFFFFu;
                                        const bool inpend = sv != 0xFFFFu && test_bit<K>(pending, sv);
                                        const unsigned same = __match_any_sync(kFull, sv);
                                        const bool first = inpend && ((uint32_t)(__ffs(same) - 1) == lane);
                                        const unsigned fm = __ballot_sync(kFull, first);
                                        if (first) {
                                            const uint32_t idx = found + __popc(fm & lanemask_lt());
                                            sA[idx] = ((uint32_t)sdict[sv] << 16) | sv;
                                            sB[idx] = (cell[blk] & 0xFFFF0000u) | (uint32_t)pin[sv];
                                        }
                                        found += __popc(fm);
#pragma unroll
                                        for (int k = 0; k < K; k++) {
                                            const uint32_t mine = (first && (sv >> 5) == (uint32_t)k) ? (1u << (sv & 31u)) : 0u;
                                            pending[k] &= ~__reduce_or_sync(kFull, mine);
                                        }
                                    }
                                }
                                __syncwarp();
                                // rank by chain start row (unique per biome within a column)
                                uint32_t ra[K], rb[K], rank[K];
#pragma unroll
                                for (int k = 0; k < K; k++) {
                                    const uint32_t idx = k * 32 + lane;
                                    rank[k] = idx;
                                    ra[k] = 0u;
                                    rb[k] = 0u;
                                    if (idx < nb) {
                                        ra[k] = sA[idx];
                                        rb[k] = sB[idx];
                                        uint32_t rk = 0u;
                                        for (uint32_t i = 0u; i < nb; i++) rk += (sB[i] >> 16) < (rb[k] >> 16);
                                        rank[k] = rk;
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
                                        const uint32_t packed = sA[e - n];
                                        id[k] = packed & 0xFFFFu;
                                        hi[k] = packed >> 16;
                                        cnt[k] = sB[e - n] & 0xFFFFu;
                                    }
                                }
                                n += nb;
                                __syncwarp();
                            }
                            // (3) membership mask of the list, columns of this batch that still hold a stranger
#pragma unroll
                            for (int kk = 0; kk < K; kk++) {
                                uint32_t mine = 0u;
#pragma unroll
                                for (int k = 0; k < K; k++) {
                                    if ((uint32_t)(k * 32) + lane < n && (id[k] >> 5) == (uint32_t)kk)
                                        mine |= 1u << (id[k] & 31u);
                                }
                                listmask[kk] = __reduce_or_sync(kFull, mine);
                            }
                            {
                                uint32_t w = 0u;
                                if (lane < (uint32_t)NB) {
#pragma unroll
                                    for (int k = 0; k < K; k++) w |= mbuf[lane * K + k] & ~listmask[k];
                                }
                                born_cols = __ballot_sync(kFull, w != 0u);
                            }
                            // (4) packed view of a short list
                            pk_steps = 0u;
                            if (n >= 1u && n <= 16u && !(g.flags & 1u)) {
                                pk_steps = 32u / n;
                                pk_slot = lane / n;
                                pk_entry = lane - pk_slot * n;
                                id[0] = __shfl_sync(kFull, id[0], pk_entry);
                                hi[0] = __shfl_sync(kFull, hi[0], pk_entry);
                                cnt[0] = __shfl_sync(kFull, cnt[0], pk_entry);
                                pk_off = pk_slot * CS + id[0];
                            }
                            if (g.dbg && lane == 0u) {
                                atomicAdd(&g.dbg[5], (unsigned long long)(clock64() - ts0));
                                atomicAdd(&g.dbg[6], 1ull);
                            }
                        }
                        if (emit) {
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                if ((uint32_t)(k * 32) + lane < n)
                                    dst[k * 32] = make_uint2(hi[k], __float_as_uint(__fmul_rn(__uint2float_rn(cnt[k]), inv)));
                            }
                            if (lane == 0u) hso_row[x] = row_first + rowpos;
                            dst += n;
                            rowpos += n;
                            x += 1u;
                        }
                        rel += 1u;
                        steps -= 1u;
                        pin += CS;
                        pout += CS;
                    }
                }
                const uint32_t adv = seg_end - c;
                in_slot += adv;
                if (in_slot >= R) in_slot -= R;
                out_slot += adv;
                if (out_slot >= R) out_slot -= R;
                c = seg_end;
            }
        } else {
            const uint32_t adv = ce - cb;
            in_slot += adv;
            if (in_slot >= R) in_slot -= R;
            out_slot += adv;
            if (out_slot >= R) out_slot -= R;
        }
        if (b + stages < n_batches) {
            __syncwarp();
            if (lane == 0u) mbar_arrive(&empty_bar[s]);
        }
    }
    if (g.dbg && lane == 0u) atomicAdd(&g.dbg[0], (unsigned long long)(clock64() - tw0));
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
