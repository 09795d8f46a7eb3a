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
//                           cvt(c,p)     = (compact id, first row of the vertical presence chain the sample at (c,p)
//                                          belongs to; the reference's vertical accumulator keeps a bin alive across
//                                          gaps <= 2r+1), column-major,
//                           tmask(y,blk) = per value the 32 column bits "occurs in the vertical window of row y",
//                           base(c,t)    = per-value counts of the vertical window of the first row of row-tile t
//   events (shf_events.cuh) : per output row the presence chains of every value along the columns = the life spans of
//                         the reference accumulator's bins, sorted by birth; their pixel spans total the bins per row
//   scans               : bins per row -> first bin of every row and chunk (in events_kernel / bases_kernel; rowscan: wide path)
//   emit (shf_events.cuh)   : a CTA owns TY consecutive rows (one consumer warp per row); producer warps rebuild the
//                         vertical window counts of the columns into a shared-memory ring of (2r+1 + 16*stages)
//                         columns x TY rows x 32K counts; consumers slide them horizontally and write every pixel's
//                         bins (the alive chains in list order) and HistogramStartOffset in place
// Shapes this does not cover (2r+1 > 511, more than 256 distinct values) take shf_generic.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#ifdef SHF_WATCHDOG
#include <cstdio>
#endif

namespace shf {

constexpr int kDictWords = 2048;      // 65536 possible sample values / 32
constexpr int kVscanThreads = 32;     // columns per CTA in vscan (one warp: small CTAs spread single-chunk calls over the SMs)
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint16_t kNoRow = 0xFFFFu;
constexpr uint32_t kTentative = 0xFFFFu;  // chain start not known inside a vscan row segment (rows are < 65535)

struct Geo {
    uint32_t W, H, r, span;        // chunk map size, radius, 2r+1
    uint32_t PW, PH;               // W+2r, H+2r : the halo-extended region
    uint32_t P;                    // row pitch (elements) of cmap / vstart
    uint32_t small_ids;            // the sample values themselves serve as compact ids (all < 32K): no remap pass, the
                                   // kernels read the caller's maps in place
    uint32_t ids_row_stride;       // where vscan / emit read compact ids: cmap (P, PH * P) or the input view
    uint64_t ids_chunk_stride;
    uint32_t n_chunks;
    uint32_t in_row_stride;        // input view: sample(n,p,c) = in[n*in_chunk_stride + p*in_row_stride + c]
    uint64_t in_chunk_stride;
    uint32_t TY, T;                // rows per emit CTA, tiles per chunk
    uint32_t K, Bpad;              // 32-biome sets, entries per count vector (= 32*K)
    uint32_t FW;                   // bits per vertical window count: 8 while 2r+1 <= 255, else 16
    uint32_t R;                    // ring columns = span + 16 * stages
    uint32_t stages;               // batches the emit producers may run ahead of the consumers (2..6)
    uint32_t stages_magic;         // ceil(2^32 / stages): x / stages = umulhi(x, magic) for the batch counts of a launch
    uint32_t producers;            // producer warps of the emit kernel (1..4)
    uint32_t VS;                   // count vector stride in the ring (= Bpad)
    uint32_t cv_pitch, cv_pad;     // column-major (compact id | chain start << 16) map: cvt[(n*PW + c)*cv_pitch + cv_pad + p]
    uint32_t vseg, vseg_rows;      // vscan row segments per column block (1 = none) and output rows per segment (x32)
    uint32_t cseg, cseg_px;        // emit column segments per row tile (1 = none) and pixels per segment (x16)
    uint32_t persist;              // != 0: emit CTAs are persistent and walk the flat tile list (large calls), see
                                   // EmitItem; = 2 * SMs to occupy + (1 if the first tile of every CTA is cut in two)
    uint32_t emit_chunk0, emit_chunks;   // the chunks one emit launch covers (a call whose result goes to the host emits
                                         // in a few ranges, each copied while the next is computed)
    float inv_total;               // 1.0f / float((2r+1)^2), SHF.cpp:495
    unsigned long long bins_cap, pool_cap;  // capacities of the bin buffer / event pool (guards of a speculative call)
};

// ------------------------------------------------------------------------------------------------------------------
// dictionary
// ------------------------------------------------------------------------------------------------------------------
// V = samples per load: 8 (one 16-byte load; needs a 16-byte aligned view and row stride) or 1
// The CTA that finishes a chunk last (arrival counter `done[n]`) turns the chunk's bitmap into the dictionary: exclusive
// prefix of the popcounts over the 2048 words and the number of distinct values, and re-arms the counter.
template <int V>
__global__ void __launch_bounds__(256) presence_kernel(const uint16_t* __restrict__ in, Geo g, uint32_t* __restrict__ bitmap,
                                                       uint32_t* __restrict__ prefix, uint32_t* __restrict__ n_biomes,
                                                       uint32_t* __restrict__ done) {
    __shared__ uint32_t bm[kDictWords];
    __shared__ uint32_t part[256];
    __shared__ uint32_t is_last, top_value;
    const uint32_t n = blockIdx.y;
    for (int i = threadIdx.x; i < kDictWords; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    const uint16_t* src = in + (size_t)n * g.in_chunk_stride;
    // values below 64 (every map of up to 64 biomes numbered from 0) stay in two registers per thread; only larger
    // values go through the shared-memory bitmap
    uint32_t lo0 = 0u, lo1 = 0u;
    bool big = false;
    auto mark = [&](uint32_t s) {
        const uint32_t bit = 1u << (s & 31u);
        if (s < 32u) {
            lo0 |= bit;
        } else if (s < 64u) {
            lo1 |= bit;
        } else {
            big = true;
            if (!(((volatile uint32_t*)bm)[s >> 5] & bit)) atomicOr(&bm[s >> 5], bit);
        }
    };
    // the CTA's rows [p0, p1) form one flat index space of 16-byte vectors, so that every thread has several
    // independent loads in flight whatever the row length
    const uint32_t rows_per = (g.PH + gridDim.x - 1u) / gridDim.x;
    const uint32_t p0 = blockIdx.x * rows_per, p1 = min(g.PH, p0 + rows_per);
    if (p0 < p1) {
        const uint32_t nv = V == 8 ? g.PW / 8u : 0u;
        if (V == 8) {
            const uint32_t total = (p1 - p0) * nv;
            for (uint32_t i0 = threadIdx.x; i0 < total; i0 += 4u * blockDim.x) {
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                    q[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (i < total) q[u] = reinterpret_cast<const uint4*>(src + (size_t)(p0 + i / nv) * g.in_row_stride)[i % nv];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (i0 + (uint32_t)u * blockDim.x < total) {
                        const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            mark(w[t] & 0xFFFFu);
                            mark(w[t] >> 16);
                        }
                    }
                }
            }
        }
        const uint32_t tail0 = nv * 8u, ntail = g.PW - tail0;   // columns the vectors do not cover
        for (uint32_t i = threadIdx.x; i < (p1 - p0) * ntail; i += blockDim.x)
            mark(src[(size_t)(p0 + i / ntail) * g.in_row_stride + tail0 + i % ntail]);
    }
    lo0 = __reduce_or_sync(kFull, lo0);
    lo1 = __reduce_or_sync(kFull, lo1);
    if ((threadIdx.x & 31u) == 0u) {
        if (lo0) atomicOr(&bitmap[(size_t)n * kDictWords], lo0);
        if (lo1) atomicOr(&bitmap[(size_t)n * kDictWords + 1], lo1);
    }
    if (__syncthreads_or(big ? 1 : 0)) {
        for (int i = threadIdx.x; i < kDictWords; i += blockDim.x)
            if (bm[i]) atomicOr(&bitmap[(size_t)n * kDictWords + i], bm[i]);
    }
    // ---- the last CTA of the chunk: dictionary prefix ----
    // (one fence by the thread that signals, after the barrier that makes it see the CTA's writes: fences are cumulative)
    __syncthreads();
    if (threadIdx.x == 0u) {
        __threadfence();
        is_last = atomicAdd(&done[n], 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const uint32_t t = threadIdx.x;
    const uint32_t* words = bitmap + (size_t)n * kDictWords;
    uint32_t local[8], sum = 0u, top = 0u;
    if (t == 0u) top_value = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        local[i] = sum;
        const uint32_t w = __ldcg(words + t * 8 + i);   // (other CTAs' atomics: read past the L1)
        sum += __popc(w);
        if (w) top = (t * 8u + (uint32_t)i) * 32u + 31u - (uint32_t)__clz((int)w);
    }
    part[t] = sum;
    __syncthreads();
    if (top) atomicMax(&top_value, top);
    for (int off = 1; off < 256; off <<= 1) {   // Hillis-Steele over the 256 partial sums
        const uint32_t v = (t >= (uint32_t)off) ? part[t - off] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const uint32_t excl = part[t] - sum;
#pragma unroll
    for (int i = 0; i < 8; i++) prefix[(size_t)n * kDictWords + t * 8 + i] = excl + local[i];
    if (t == 255u) {
        n_biomes[n] = part[255];
        n_biomes[gridDim.y + n] = top_value;   // largest sample value of the chunk (the scan's barriers order the atomicMax)
    }
    if (t == 0u) done[n] = 0u;
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
    // (the clamp only bites when a repeated call runs with the previous call's plan and this map has more distinct
    // values than that plan holds: the result is then discarded and recomputed, but no kernel may index out of range)
    auto rank = [&](uint32_t s) { return min(pf[s >> 5] + __popc(bm[s >> 5] & ((1u << (s & 31u)) - 1u)), dict_stride - 1u); };
    const uint16_t* src = in + (size_t)n * g.in_chunk_stride;
    const uint32_t rows_per = (g.PH + gridDim.x - 1u) / gridDim.x;
    const uint32_t p0 = blockIdx.x * rows_per, p1 = min(g.PH, p0 + rows_per);
    if (p0 < p1) {
        uint16_t* dst = cmap + ((size_t)n * g.PH + p0) * g.P;
        const uint32_t nv = V == 8 ? g.PW / 8u : 0u;
        if (V == 8) {
            const uint32_t total = (p1 - p0) * nv;
            for (uint32_t i0 = threadIdx.x; i0 < total; i0 += 4u * blockDim.x) {
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                    q[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (i < total) q[u] = reinterpret_cast<const uint4*>(src + (size_t)(p0 + i / nv) * g.in_row_stride)[i % nv];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                    if (i < total) {
                        uint4 o;
                        o.x = rank(q[u].x & 0xFFFFu) | (rank(q[u].x >> 16) << 16);
                        o.y = rank(q[u].y & 0xFFFFu) | (rank(q[u].y >> 16) << 16);
                        o.z = rank(q[u].z & 0xFFFFu) | (rank(q[u].z >> 16) << 16);
                        o.w = rank(q[u].w & 0xFFFFu) | (rank(q[u].w >> 16) << 16);
                        reinterpret_cast<uint4*>(dst + (size_t)(i / nv) * g.P)[i % nv] = o;
                    }
                }
            }
        }
        const uint32_t tail0 = nv * 8u, ntail = g.PW - tail0;   // columns the vectors do not cover
        for (uint32_t i = threadIdx.x; i < (p1 - p0) * ntail; i += blockDim.x) {
            const uint32_t pr = i / ntail, c = tail0 + i % ntail;
            dst[(size_t)pr * g.P + c] = (uint16_t)rank(src[(size_t)(p0 + pr) * g.in_row_stride + c]);
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
// One warp = one block of 32 columns, one thread per column walking down the rows.
// Per (compact id, column) state in shared memory, one 32-bit word at st[id * 32 + lane]: window count in the low half,
// chain-start row in the high half. Every lane always touches its own bank, whatever the ids.
// The window holds rows [p - span, p - 1] when row p is about to enter (SHF.cpp:549-550: increment, then decrement), so
// "no occurrence within the last span rows" (the start of a new vertical chain) is simply count == 0.
// Output per (row y, column block): the presence masks TRANSPOSED, tmask(n, y, block)[id] = 32 bits, bit j set iff
// compact id `id` occurs in column 32*block + j, rows [y, y+2r]. They are kept incrementally in shared memory (a count
// reaching / leaving zero flips one bit with a shared-memory atomic), so events_kernel reads them lane = id as they are.
// (free functions taking plain values: a noinline member would force the whole state struct into local memory)
constexpr uint32_t kVscanStageRows = 16;                       // rows staged per flush (a 64-byte line per column)
constexpr uint32_t kVscanStagePitch = kVscanStageRows + 1;     // words per staged column (odd: conflict-free pushes)
__device__ __noinline__ void vscan_flush(const uint32_t* stg, uint32_t* cv_blk, uint32_t cv_pitch, uint32_t ncols,
                                         uint32_t line0, uint32_t lane) {
    __syncwarp();
    const uint32_t row = lane & (kVscanStageRows - 1u);
    uint32_t* dst = cv_blk + line0 + row;
    // two columns per store instruction: lanes 0-15 the even one, lanes 16-31 the odd one
    for (uint32_t col = lane >> 4; col < ncols; col += 2u) dst[(size_t)col * cv_pitch] = stg[col * kVscanStagePitch + row];
    __syncwarp();
}
// a tile's base vector: the window counts (<= 2r+1) by compact id, as bytes (fw = 8) or 16-bit words (fw = 16)
template <int K>
__device__ __noinline__ void vscan_dump(const uint32_t* st, uint8_t* dst8, uint32_t fw) {
    uint4* dst = reinterpret_cast<uint4*>(dst8);
    if (fw == 8u) {
#pragma unroll 2
        for (int q = 0; q < 2 * K; q++) {
            uint32_t v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t* src = st + (size_t)(16 * q + 4 * e) * kVscanThreads;
                v[e] = (src[0] & 0xFFu) | ((src[kVscanThreads] & 0xFFu) << 8) | ((src[2 * kVscanThreads] & 0xFFu) << 16) |
                       (src[3 * kVscanThreads] << 24);
            }
            dst[q] = make_uint4(v[0], v[1], v[2], v[3]);
        }
    } else {
#pragma unroll 2
        for (int q = 0; q < 4 * K; q++) {
            uint32_t v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t* src = st + (size_t)(8 * q + 2 * e) * kVscanThreads;
                v[e] = (src[0] & 0xFFFFu) | (src[kVscanThreads] << 16);
            }
            dst[q] = make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
}

template <int K>
struct VscanState {
    uint32_t* st;        // this thread's column of the state table
    uint32_t* tm;        // the warp's transposed masks, [32K]
    uint32_t lanebit;    // 1 << lane, or 0 for a lane beyond the last column
    uint32_t lane;

    // row p enters: returns the start row of the (possibly new) vertical chain of its sample
    __device__ __forceinline__ uint32_t enter(uint32_t s, uint32_t p) {
        uint32_t w = st[s * kVscanThreads];
        const bool born = (w & 0xFFFFu) == 0u;
        w = born ? p << 16 : w;
        st[s * kVscanThreads] = w + 1u;
        if (born) atomicOr(&tm[s], lanebit);
        return w >> 16;
    }
    // replay of a row segment's first window: counts only, every chain met here began somewhere above
    __device__ __forceinline__ void enter_tentative(uint32_t s) {
        uint32_t w = st[s * kVscanThreads];
        if ((w & 0xFFFFu) == 0u) {
            w = kTentative << 16;
            atomicOr(&tm[s], lanebit);
        }
        st[s * kVscanThreads] = w + 1u;
    }
    __device__ __forceinline__ void leave(uint32_t o) {
        const uint32_t w = st[o * kVscanThreads] - 1u;
        st[o * kVscanThreads] = w;
        if ((w & 0xFFFFu) == 0u) atomicAnd(&tm[o], ~lanebit);
    }
    // row p enters and the row 2r+1 above it leaves, in one go: both state words are loaded together (one shared-memory
    // round trip instead of two). When both rows hold the same value nothing changes: its count is >= 1 already (the
    // leaving row is inside the window), so no chain starts, and +1 -1 cancel.
    __device__ __forceinline__ uint32_t step(uint32_t s, uint32_t o, uint32_t p) {
        uint32_t ws = st[s * kVscanThreads], wo = st[o * kVscanThreads];
        const bool born = (ws & 0xFFFFu) == 0u;
        ws = born ? p << 16 : ws;
        if (s != o) {
            st[s * kVscanThreads] = ws + 1u;
            wo -= 1u;
            st[o * kVscanThreads] = wo;
            if (born) atomicOr(&tm[s], lanebit);
            if ((wo & 0xFFFFu) == 0u) atomicAnd(&tm[o], ~lanebit);
        }
        return ws >> 16;
    }
    __device__ __forceinline__ void store_mask(uint32_t* dst) const {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; k++) dst[k * 32 + lane] = tm[k * 32 + lane];
        __syncwarp();
    }
    // (compact id | chain start << 16) of every cell goes to a column-major map (a column's window is contiguous for
    // events_kernel). 16 rows are staged in shared memory ([column][17] words) and leave as one 64-byte line per column
    // (a smaller stage than a full 128-byte line: 20 instead of 17 of these one-warp CTAs fit an SM, and the kernel is
    // bound by the latency of its per-row chain, i.e. by the warps in flight).
    uint32_t* stg;        // the warp's staging tile
    uint32_t* cv_blk;     // the block's first column in the map, at index 0
    uint32_t cv_pitch, ncols;
    __device__ __forceinline__ void flush(uint32_t line0) const { vscan_flush(stg, cv_blk, cv_pitch, ncols, line0, lane); }
    __device__ __forceinline__ void push(uint32_t idx, uint32_t value) const {
        stg[lane * kVscanStagePitch + (idx & (kVscanStageRows - 1u))] = value;
    }

    uint32_t fw;
    __device__ __forceinline__ void dump(uint8_t* dst8) const { vscan_dump<K>(st, dst8, fw); }
};

template <int K, bool ALIGNED>   // ALIGNED: rows per tile (TY) is a multiple of 8
__global__ void __launch_bounds__(kVscanThreads) vscan_kernel(Geo g, const uint16_t* __restrict__ cmap,
                                                              uint32_t* __restrict__ cvt, uint8_t* __restrict__ base,
                                                              uint32_t* __restrict__ tmask, uint32_t* __restrict__ vexit) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int Bpad = 32 * K;
    constexpr int T = kVscanThreads;
    static_assert(T == 32, "one warp per column block");
    const uint32_t n = blockIdx.y, lane = threadIdx.x;
    const bool valid = blockIdx.x * T + lane < g.PW;
    const uint32_t c = min(blockIdx.x * T + lane, g.PW - 1u);  // lanes beyond the last column shadow it, without output
    VscanState<K> vs;
    vs.st = reinterpret_cast<uint32_t*>(smem) + lane;
    vs.tm = reinterpret_cast<uint32_t*>(smem) + Bpad * T;
    vs.stg = vs.tm + Bpad;
    vs.cv_blk = cvt + ((size_t)n * g.PW + blockIdx.x * T) * g.cv_pitch;
    vs.cv_pitch = g.cv_pitch;
    vs.ncols = min((uint32_t)T, g.PW - blockIdx.x * T);
    vs.lanebit = valid ? 1u << lane : 0u;
    vs.lane = lane;
    vs.fw = g.FW;
#pragma unroll 8
    for (int i = 0; i < Bpad; i++) vs.st[i * T] = 0u;
#pragma unroll
    for (int k = 0; k < K; k++) vs.tm[k * 32 + lane] = 0u;
    __syncwarp();
    const size_t P = g.ids_row_stride;
    // (ids are clamped into the state table: only a call running ahead with a plan that does not fit its input -- its
    // result is discarded -- can meet larger ones, when it reads the caller's raw samples as ids)
    const uint32_t id_max = (uint32_t)Bpad - 1u;
    const uint16_t* in = cmap + (size_t)n * g.ids_chunk_stride + c;      // row p (entering)
    uint32_t ci = g.cv_pad;   // index of row p in the column-major map: cv_pad + p (the 8-row groups start on a line)
    const uint32_t nblk = gridDim.x;
    uint32_t* mout = tmask + ((size_t)n * g.H * nblk + blockIdx.x) * Bpad;   // row y
    const size_t mstep = (size_t)nblk * Bpad;
    const size_t vbytes = (size_t)Bpad * g.FW / 8u;               // bytes of one base vector
    uint8_t* bout = base + ((size_t)n * g.T * g.PW + c) * vbytes;  // tile 0
    const size_t bstep = (size_t)g.PW * vbytes;
    const uint32_t two_r = 2u * g.r;

    // Row segments (blockIdx.z; small calls only, see vseg in the plan): segment k > 0 starts at output row
    // Y = 1 + k * vseg_rows. It rebuilds the window state of row Y - 1 by replaying that window's 2r+1 rows -- the
    // counts come out exact, the chain starts cannot (a chain may have begun anywhere above): they are marked
    // kTentative, travel into cvt as such while the chain lasts, and events_kernel, the only reader, looks the few it
    // needs up in the exit state of the segment above (written below; resolve_start in shf_events.cuh).
    const uint32_t seg = blockIdx.z, n_seg = g.vseg;
    const uint32_t y_first = seg == 0u ? 0u : 1u + seg * g.vseg_rows;
    const uint32_t y_end = seg + 1u < n_seg ? 1u + (seg + 1u) * g.vseg_rows : g.H;
    uint32_t next_dump = g.TY;   // the next tile's first row
    if (seg == 0u) {
        // ---- rows 0 .. 2r: the first window fills up, nothing leaves ----
        uint32_t p = 0u;
        for (; p + 8u <= two_r + 1u; p += 8u) {
            uint32_t s_in[8];
#pragma unroll
            for (int j = 0; j < 8; j++) s_in[j] = min((uint32_t)in[j * P], id_max);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                vs.push(ci, s_in[j] | (vs.enter(s_in[j], p + j) << 16));
                if ((++ci & (kVscanStageRows - 1u)) == 0u) vs.flush(ci - kVscanStageRows);
            }
            in += 8 * P;
        }
        for (; p <= two_r; p++) {
            const uint32_t sv = min((uint32_t)*in, id_max);
            vs.push(ci, sv | (vs.enter(sv, p) << 16));
            if ((++ci & (kVscanStageRows - 1u)) == 0u) vs.flush(ci - kVscanStageRows);
            in += P;
        }
        vs.store_mask(mout);
        mout += mstep;
        if (valid) vs.dump(bout);
        bout += bstep;
    } else {
        // ---- replay of the window of output row y_first - 1: rows y_first - 1 .. y_first + 2r - 1 ----
        in += (size_t)(y_first - 1u) * P;
        for (uint32_t q = 0u; q <= two_r; q++) {
            vs.enter_tentative(min((uint32_t)*in, id_max));
            in += P;
        }
        ci = g.cv_pad + y_first + two_r;                       // a multiple of 32 by construction
        mout += (size_t)y_first * mstep;
        const uint32_t t0 = (y_first + g.TY - 1u) / g.TY;       // first tile starting inside this segment
        bout += (size_t)t0 * bstep;
        next_dump = t0 * g.TY;
    }

    // ---- output rows 1 .. H-1: row y + 2r enters, row y - 1 leaves ----
    uint32_t y = seg == 0u ? 1u : y_first;
    const uint16_t* out = cmap + (size_t)n * g.ids_chunk_stride + c + (size_t)(y - 1u) * P;     // row y - 1 (leaving)
    // groups of 8 rows; the samples of a group are loaded while the group before it is processed (one DRAM round trip
    // hidden per group). With TY a multiple of 8 a tile's first row is always the last row of a group (ALIGNED); other
    // plans check every row.
    {
        uint32_t n_in[8], n_out[8];
        if (y + 8u <= y_end) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                n_in[j] = in[j * P];
                n_out[j] = out[j * P];
            }
        }
        for (; y + 8u <= y_end; y += 8u) {
            uint32_t s_in[8], s_out[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                s_in[j] = min(n_in[j], id_max);
                s_out[j] = min(n_out[j], id_max);
            }
            if (y + 16u <= y_end) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    n_in[j] = in[(8 + j) * P];
                    n_out[j] = out[(8 + j) * P];
                }
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                vs.push(ci + j, s_in[j] | (vs.step(s_in[j], s_out[j], two_r + y + j) << 16));
                vs.store_mask(mout + j * mstep);
                if (!ALIGNED && y + j == next_dump) {
                    if (valid) vs.dump(bout);
                    bout += bstep;
                    next_dump += g.TY;
                }
            }
            ci += 8u;
            if ((ci & (kVscanStageRows - 1u)) == 0u) vs.flush(ci - kVscanStageRows);
            in += 8 * P;
            out += 8 * P;
            mout += 8 * mstep;
            if (ALIGNED && y + 7u == next_dump) {
                if (valid) vs.dump(bout);
                bout += bstep;
                next_dump += g.TY;
            }
        }
    }
    for (; y < y_end; y++) {
        const uint32_t sv = min((uint32_t)*in, id_max);
        vs.push(ci, sv | (vs.step(sv, min((uint32_t)*out, id_max), two_r + y) << 16));
        if ((++ci & (kVscanStageRows - 1u)) == 0u) vs.flush(ci - kVscanStageRows);
        vs.store_mask(mout);
        in += P;
        out += P;
        mout += mstep;
        if (y == next_dump) {
            if (valid) vs.dump(bout);
            bout += bstep;
            next_dump += g.TY;
        }
    }
    if (ci & (kVscanStageRows - 1u)) vs.flush(ci & ~(kVscanStageRows - 1u));
    // exit state for the segment below: (count, start) of every value in this column
    if (seg + 1u < n_seg) {
        uint32_t* ex = vexit + ((((size_t)n * (n_seg - 1u) + seg) * gridDim.x + blockIdx.x) * Bpad) * T + lane;
#pragma unroll 8
        for (int i = 0; i < Bpad; i++) ex[i * T] = vs.st[i * T];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// helpers of the emit kernel (shf_events.cuh): lane masks, shared-memory mbarriers
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag = 0u) {
    uint32_t done;
#ifdef SHF_WATCHDOG   // debug builds: a wait that never ends reports itself and stops the kernel
    unsigned long long spins = 0ull;
#endif
    (void)tag;
    do {
#ifdef SHF_WATCHDOG
        if (++spins == (1ull << 24)) {
            if ((threadIdx.x & 31u) == 0u)
                printf("mbar_wait stuck: block %u warp %u tag %u parity %u\n", blockIdx.x, threadIdx.x >> 5, tag, parity);
            __trap();
        }
#endif
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// columns per batch handed from the producer warps to the consumer warps of emit_kernel
constexpr int kMarchNB = 16;

// ------------------------------------------------------------------------------------------------------------------
// first-bin index of every row (exclusive scan of the bins-per-row), chunk totals, the closing offset entry
// ------------------------------------------------------------------------------------------------------------------
// Exclusive scan of `count` 32- or 64-bit values by one CTA of `T` threads (T <= 1024, a multiple of 32), 64-bit sums:
// out[i] = sum of in[0..i) (+ nothing else), returns the total to every thread. `warp_part` = 32 shared 64-bit words.
template <typename TIn, typename TOut>
__device__ __forceinline__ unsigned long long cta_exclusive_scan(const TIn* in, TOut* out, uint32_t count,
                                                                 unsigned long long* warp_part) {
    const uint32_t t = threadIdx.x, T = blockDim.x, lane = t & 31u, warp = t >> 5, n_warps = T >> 5;
    unsigned long long carry = 0ull;
    for (uint32_t i0 = 0u; i0 < count; i0 += T) {
        const uint32_t i = i0 + t;
        const unsigned long long v = i < count ? (unsigned long long)__ldcg(in + i) : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long u = __shfl_up_sync(kFull, incl, off);
            if (lane >= (uint32_t)off) incl += u;
        }
        __syncthreads();   // warp_part of the previous round has been read
        if (lane == 31u) warp_part[warp] = incl;
        __syncthreads();
        unsigned long long before = 0ull, round = 0ull;
        for (uint32_t w = 0u; w < n_warps; w++) {
            const unsigned long long u = warp_part[w];
            if (w < warp) before += u;
            round += u;
        }
        if (i < count) out[i] = (TOut)(carry + before + incl - v);
        carry += round;
    }
    return carry;
}

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
