// shf_events.cuh -- the event-list formulation of the horizontal pass (the default path for up to 256 distinct values).
//
// The reference's horizontal accumulator (SHF.cpp:608-679) keeps a bin for a sample value from the column where the
// value enters the sliding window until the step in which its count returns to zero (SHF.cpp:435-445); a value that
// re-enters later gets a NEW bin at the end of the list (SHF.cpp:411-415). So along one output row every sample value
// owns a sequence of "presence chains": maximal runs of columns holding the value in their vertical window with gaps of
// at most 2r+1 columns. Each chain is one EVENT (value, birth column c_b, last presence column c_l):
//   * the chain's bin exists for the pixels x with  max(c_b - 2r, 0) <= x <= min(c_l, W - 1)
//   * a pixel lists the bins of the events alive at x, ordered by birth: (c_b, vertical chain start of the value in
//     column c_b) -- the second key is the order of the column's own intermediate histogram (SHF.cpp:522-563).
// events_kernel builds the sorted event list of every row from the transposed presence masks of vscan_kernel and
// totals the bins per row (which replaces a counting march). emit_kernel then needs no sequential list maintenance at
// all: a pixel's bins are the alive events in list order (warp ballot + popc gives the position), the counts come from
// a dense per-row horizontal sliding sum kept next to the shared-memory ring of vertical window counts.
#pragma once
#include "shf_kernels.cuh"

namespace shf {

constexpr int kEventWarps = 4;          // rows per CTA of events_kernel
constexpr uint32_t kEventStage = 192;   // events of one row staged in shared memory before their pool slot is known
constexpr uint32_t kNoEvent = 0xFFFFFFFFu;
constexpr int32_t kNeverSeen = -0x40000000;   // "last column" of a value not met yet: further back than any window

// Event record, 8 bytes: x = compact id | sample value << 16, y = first pixel | (last pixel + 1) << 16
// 32-column mask blocks staged per cp.async group (two groups = 4 KB in flight per warp)
__host__ __device__ constexpr int event_group(int K) { return K <= 2 ? 8 : K == 4 ? 4 : 2; }

// The transposed masks of `event_group(K)` blocks of 32 columns, copied asynchronously into shared memory: the walk
// below consumes a block every few dozen instructions, so without many bytes in flight per warp it would be bound by
// DRAM latency.
template <int K>
__device__ __forceinline__ void stage_masks(uint32_t* sdst, const uint32_t* __restrict__ tmask_row, uint32_t w0,
                                            uint32_t n_words, uint32_t lane) {
    constexpr int kEventGroup = event_group(K);
#pragma unroll
    for (int i = 0; i < kEventGroup; i++) {
        if (w0 + (uint32_t)i < n_words) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t* src = tmask_row + ((size_t)(w0 + i) * K + k) * 32u + lane;
                const uint32_t dst = smem_addr(sdst + (i * K + k) * 32 + lane);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// One pass over the row's presence masks. Writes at most `cap` records to `dst` (shared staging or the row's pool
// slot); returns the number of events and adds the row's bins to `bins_total` (per-lane partial sums).
// Chain start of a value whose cells in the window of output row y, column c were all written as kTentative by
// vscan_kernel: they entered in row segments >= 1 of the vertical scan while the chain was alive since before those
// segments began. The chain is the one the value had when the segment above row y's ended (every cell of the window had
// entered by then or belongs to row y's own segment), so its start is what that segment ended with for the value -- or,
// if that is tentative as well, what the one above it ended with (vexit = exit states of the segments).
__device__ __forceinline__ uint32_t resolve_start(const Geo& g, const uint32_t* __restrict__ vexit_chunk, uint32_t c,
                                                  uint32_t y, uint32_t id) {
    const uint32_t seg = min((max(y, 1u) - 1u) / g.vseg_rows, g.vseg - 1u);
    const uint32_t nblk = (g.PW + 31u) / 32u;
    uint32_t start = kTentative;
    for (int k = (int)seg - 1; k >= 0 && start == kTentative; k--)
        start = vexit_chunk[(((size_t)k * nblk + c / 32u) * g.Bpad + id) * 32u + (c & 31u)] >> 16;
    return start;
}

template <int K>
__device__ __forceinline__ uint32_t walk_row(const Geo& g, uint32_t y, uint32_t lane, const uint32_t* __restrict__ cmask_row,
                                             const uint32_t* __restrict__ cv, const uint32_t* __restrict__ vexit_chunk,
                                             const uint32_t (&item)[K], uint16_t* tab, uint32_t* mstage, uint2* dst,
                                             uint32_t cap, uint32_t& bins_total) {
    const uint32_t PW = g.PW, span = g.span, two_r = 2u * g.r, W = g.W;
    int32_t last[K];
    uint32_t open[K], openxb[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        last[k] = kNeverSeen;
        open[k] = kNoEvent;
        openxb[k] = 0u;
    }
    uint32_t E = 0u;
    const uint32_t n_words = (PW + 31u) / 32u;
    constexpr int kEventGroup = event_group(K);
    constexpr uint32_t GW = kEventGroup * 32 * K;  // words of one staged group
    __syncwarp();
    stage_masks<K>(mstage, cmask_row, 0u, n_words, lane);
    stage_masks<K>(mstage + GW, cmask_row, (uint32_t)kEventGroup, n_words, lane);
    for (uint32_t w = 0u; w < n_words; w++) {
        const uint32_t c0 = w * 32u;
        const uint32_t gi = w / kEventGroup, wi = w % kEventGroup;
        if (wi == 0u) {
            if (w) {  // the buffer of the group just finished is free: stage the group after the next into it
                __syncwarp();
                stage_masks<K>(mstage + ((gi + 1u) & 1u) * GW, cmask_row, (gi + 1u) * kEventGroup, n_words, lane);
            }
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
        }
        uint32_t rem[K];  // lane = compact id 32k + lane, bit j = column c0 + j
#pragma unroll
        for (int k = 0; k < K; k++) rem[k] = mstage[(gi & 1u) * GW + (wi * K + k) * 32u + lane];
        if (span >= 32u) {
            // Most words hold no birth at all (on dense maps every value is present in every column: nothing is born
            // after column 0): when no lane's first column in the word opens a chain -- with 2r+1 >= 32 no later one
            // of the word can either -- the chains just carry on to their last column in the word.
            uint32_t noisy = 0u;   // (branch-free: ffs(0) = 0 gives a column that is masked out again)
#pragma unroll
            for (int k = 0; k < K; k++)
                noisy |= (rem[k] != 0u) & ((int32_t)(c0 + (uint32_t)__ffs((int)rem[k]) - 1u) - last[k] > (int32_t)span);
            if (!__any_sync(kFull, noisy)) {
#pragma unroll
                for (int k = 0; k < K; k++)
                    if (rem[k]) last[k] = (int32_t)(c0 + 31u - (uint32_t)__clz((int)rem[k]));
                continue;
            }
        }
        for (;;) {
            // every lane advances to its next birth inside this word
            uint32_t bj[K], jmin = 32u;
#pragma unroll
            for (int k = 0; k < K; k++) {
                bj[k] = 32u;
                while (rem[k]) {
                    const uint32_t j = (uint32_t)__ffs((int)rem[k]) - 1u;
                    const int32_t c = (int32_t)(c0 + j);
                    if (last[k] < 0 || c - last[k] > (int32_t)span) {
                        bj[k] = j;
                        break;
                    }
                    if (span >= 32u) {  // the rest of the word cannot hold a gap wider than the window
                        last[k] = (int32_t)(c0 + 31u - (uint32_t)__clz((int)rem[k]));
                        rem[k] = 0u;
                    } else {
                        last[k] = c;
                        rem[k] &= rem[k] - 1u;
                    }
                }
                jmin = min(jmin, bj[k]);
            }
            jmin = __reduce_min_sync(kFull, jmin);
            if (jmin == 32u) break;
            const uint32_t c = c0 + jmin;
            // the values born in column c
            bool mem[K];
            unsigned mb[K];
            uint32_t gsz = 0u;
#pragma unroll
            for (int k = 0; k < K; k++) {
                mem[k] = bj[k] == jmin;
                mb[k] = __ballot_sync(kFull, mem[k]);
                gsz += (uint32_t)__popc(mb[k]);
            }
            uint32_t rank[K];
#pragma unroll
            for (int k = 0; k < K; k++) rank[k] = 0u;
            if (gsz > 1u) {
                // order them by the start row of their vertical chain: every cell of the column's window publishes the
                // chain start of its value (all cells of one value inside a window agree)
                // (2r+1 <= 511 on this path: two rounds of 8 cells per lane; the loads of a round go out before its stores)
                const uint32_t* colw = cv + (size_t)c * g.cv_pitch + g.cv_pad + y;  // the column's window, contiguous
                if (g.vseg > 1u) {   // (cells may carry tentative starts: a value keeps 0xFFFF unless a cell knows better)
#pragma unroll
                    for (int k = 0; k < K; k++) tab[k * 32 + lane] = (uint16_t)kTentative;
                    __syncwarp();
                }
                for (uint32_t o0 = 0u; o0 < span; o0 += 256u) {
                    uint32_t cell[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t o = o0 + lane + 32u * (uint32_t)i;
                        if (o < span) cell[i] = colw[o];
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t o = o0 + lane + 32u * (uint32_t)i;
                        if (o < span && (cell[i] >> 16) != kTentative) tab[cell[i] & 0xFFFFu] = (uint16_t)(cell[i] >> 16);
                    }
                }
                __syncwarp();
                uint32_t vk[K];
#pragma unroll
                for (int k = 0; k < K; k++) {
                    vk[k] = mem[k] ? (uint32_t)tab[k * 32 + lane] : 0u;
                    if (mem[k] && vk[k] == kTentative) vk[k] = resolve_start(g, vexit_chunk, c, y, (uint32_t)(k * 32) + lane);
                }
                // rank = how many of the column's newborn have an earlier chain start (starts are distinct: one sample per
                // cell). Few newborn: one shuffle each. Many (the first column of a dense map gives birth to every value):
                // their starts go back into the table, 0xFFFF for the others, and every lane counts through it with
                // 16-byte broadcast loads -- a quarter of the instructions of 64 shuffle rounds.
                constexpr uint32_t kTableRank = K == 1 ? 4u : K == 2 ? 12u : K == 4 ? 42u : 156u;
                if (gsz > kTableRank) {
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < K; k++) tab[k * 32 + lane] = mem[k] ? (uint16_t)vk[k] : (uint16_t)0xFFFFu;
                    __syncwarp();
                    const uint4* t4 = reinterpret_cast<const uint4*>(tab);
#pragma unroll 2
                    for (int q = 0; q < 4 * K; q++) {
                        const uint4 w4 = t4[q];
                        const uint32_t h[8] = {w4.x & 0xFFFFu, w4.x >> 16, w4.y & 0xFFFFu, w4.y >> 16,
                                               w4.z & 0xFFFFu, w4.z >> 16, w4.w & 0xFFFFu, w4.w >> 16};
#pragma unroll
                        for (int e = 0; e < 8; e++)
#pragma unroll
                            for (int k = 0; k < K; k++) rank[k] += (h[e] < vk[k]) ? 1u : 0u;
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < K; kk++) {
                        unsigned m2 = mb[kk];
                        while (m2) {
                            const int l = __ffs((int)m2) - 1;
                            const uint32_t v = __shfl_sync(kFull, vk[kk], l);
#pragma unroll
                            for (int k = 0; k < K; k++) rank[k] += (v < vk[k]) ? 1u : 0u;
                            m2 &= m2 - 1u;
                        }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (mem[k]) {
                    if (open[k] != kNoEvent) {  // the previous chain of this value ended at column last[k]
                        const uint32_t xd = min((uint32_t)last[k] + 1u, W);
                        if (open[k] < cap) dst[open[k]].y = openxb[k] | (xd << 16);
                        bins_total += xd - openxb[k];
                    }
                    const uint32_t idx = E + rank[k];
                    const uint32_t xb = c > two_r ? c - two_r : 0u;
                    if (idx < cap) dst[idx] = make_uint2((uint32_t)(k * 32) + lane | (item[k] << 16), xb);
                    open[k] = idx;
                    openxb[k] = xb;
                    if (span >= 32u) {
                        last[k] = (int32_t)(c0 + 31u - (uint32_t)__clz((int)rem[k]));
                        rem[k] = 0u;
                    } else {
                        last[k] = (int32_t)c;
                        rem[k] &= rem[k] - 1u;
                    }
                }
            }
            E += gsz;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (open[k] != kNoEvent) {
            const uint32_t xd = min((uint32_t)last[k] + 1u, W);
            if (open[k] < cap) dst[open[k]].y = openxb[k] | (xd << 16);
            bins_total += xd - openxb[k];
        }
    }
    return E;
}

// the event list of one output row (one warp): rowinfo(n, y) = (first record of the row in the pool, number of records).
// counter[0] = records handed out so far; a row whose slot would end beyond `pool_cap` writes nothing (the host grows
// the pool and runs the kernel again).
template <int K>
__device__ __forceinline__ void events_row(const Geo& g, uint32_t n, uint32_t y, uint32_t lane,
                                           const uint32_t* __restrict__ colmask, const uint32_t* __restrict__ cvt,
                                           const uint16_t* __restrict__ dict, uint32_t dict_stride,
                                           uint2* __restrict__ pool, unsigned long long pool_cap,
                                           unsigned long long* __restrict__ counter, uint2* __restrict__ rowinfo,
                                           uint32_t* __restrict__ rowtotal, const uint32_t* __restrict__ vexit,
                                           uint2* stage, uint16_t* tab, uint32_t* mstage) {
    // exit states of the vertical scan's row segments of this chunk (small calls only, g.vseg > 1)
    const uint32_t* vexit_chunk = vexit + (size_t)n * (g.vseg - 1u) * ((g.PW + 31u) / 32u) * g.Bpad * 32u;
    const uint32_t* cmask_row = colmask + ((size_t)n * g.H + y) * ((g.PW + 31u) / 32u) * 32u * K;  // transposed masks
    const uint32_t* cv = cvt + (size_t)n * g.PW * g.cv_pitch;
    uint32_t item[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t id = k * 32 + lane;
        item[k] = g.small_ids ? id : id < dict_stride ? (uint32_t)dict[(size_t)n * dict_stride + id] : 0u;
    }
    uint32_t bins = 0u;
    const uint32_t E = walk_row<K>(g, y, lane, cmask_row, cv, vexit_chunk, item, tab, mstage, stage, kEventStage, bins);
    bins = __reduce_add_sync(kFull, bins);
    unsigned long long first = 0ull;
    if (lane == 0u) first = atomicAdd(counter, (unsigned long long)E);
    first = __shfl_sync(kFull, first, 0);
    if (lane == 0u) {
        rowinfo[(size_t)n * g.H + y] = make_uint2((uint32_t)first, E);
        rowtotal[(size_t)n * g.H + y] = bins;
    }
    if (first + E > pool_cap) return;
    __syncwarp();
    if (E <= kEventStage) {
        for (uint32_t i = lane; i < E; i += 32u) pool[first + i] = stage[i];
    } else {
        uint32_t again = 0u;
        (void)walk_row<K>(g, y, lane, cmask_row, cv, vexit_chunk, item, tab, mstage, pool + first, E, again);
    }
}

// One warp per output row.
// FOLD (small calls, where every launch counts): the CTA that finishes a chunk last (sync[1 + n]) scans the chunk's bins
// per row into the first-bin index of every row (rowbase), the chunk total and the closing offset entry; the CTA that
// finishes the last chunk (sync[0]) scans the chunk totals into the first-bin index of every chunk and hands the event
// count to the host (chunktotal[n_chunks]). Both re-arm their counters. Large calls leave that to bases_kernel: the
// fence every CTA pays before it signals costs more there (+10 % on a 256-chunk batch) than the launch it saves.
template <int K, bool FOLD>
__global__ void __launch_bounds__(kEventWarps * 32, K == 1 ? 6 : K == 2 ? 8 : K == 4 ? 4 : 3)
    events_kernel(Geo g, const uint32_t* __restrict__ colmask, const uint32_t* __restrict__ cvt,
                  const uint32_t* __restrict__ vexit, const uint16_t* __restrict__ dict, uint32_t dict_stride,
                  uint2* __restrict__ pool, unsigned long long pool_cap, unsigned long long* __restrict__ counter,
                  uint2* __restrict__ rowinfo, uint32_t* __restrict__ rowtotal, uint32_t* __restrict__ rowbase,
                  unsigned long long* __restrict__ chunktotal, unsigned long long* __restrict__ chunkbase,
                  uint32_t* __restrict__ hso, uint32_t* __restrict__ sync) {
    __shared__ uint2 stage_all[kEventWarps][kEventStage];
    __shared__ __align__(16) uint16_t tab_all[kEventWarps][32 * K];
    __shared__ __align__(16) uint32_t mstage_all[kEventWarps][2 * event_group(K) * 32 * K];
    __shared__ unsigned long long warp_part[32];
    __shared__ uint32_t last_flag;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n = blockIdx.y, y = blockIdx.x * kEventWarps + warp;
    if (y < g.H) events_row<K>(g, n, y, lane, colmask, cvt, dict, dict_stride, pool, pool_cap, counter, rowinfo, rowtotal,
                               vexit, stage_all[warp], tab_all[warp], mstage_all[warp]);
    if (!FOLD) return;
    // ---- last CTA of the chunk: first-bin index of every row ----
    // (one fence by the thread that signals, after the barrier that makes it see the CTA's writes: fences are cumulative)
    __syncthreads();
    if (threadIdx.x == 0u) {
        __threadfence();
        last_flag = atomicAdd(&sync[1u + n], 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!last_flag) return;
    __threadfence();
    // (u32 row bases: garbage if the chunk total overflows 32 bits; the host checks the total)
    const unsigned long long total = cta_exclusive_scan(rowtotal + (size_t)n * g.H, rowbase + (size_t)n * g.H, g.H, warp_part);
    __syncthreads();
    if (threadIdx.x == 0u) {
        chunktotal[n] = total;
        hso[(size_t)n * ((size_t)g.W * g.H + 1u) + (size_t)g.W * g.H] = (uint32_t)total;
        sync[1u + n] = 0u;
        __threadfence();
        last_flag = atomicAdd(&sync[0], 1u) == gridDim.y - 1u;
    }
    __syncthreads();
    if (!last_flag) return;
    // ---- last chunk of the call: first-bin index of every chunk ----
    __threadfence();
    const unsigned long long all = cta_exclusive_scan(chunktotal, chunkbase, g.n_chunks, warp_part);
    if (threadIdx.x == 0u) {
        chunkbase[g.n_chunks] = all;
        chunktotal[g.n_chunks] = *counter;   // records asked for by this call (read by the host)
        *counter = 0ull;
        sync[0] = 0u;
    }
}

// The same scans for large calls behind events_kernel<K, false>: one WARP per chunk (a chunk has a few hundred rows: two
// passes of 32 coalesced loads and a shuffle scan each), eight chunks per CTA; the CTA that finishes last scans the chunk
// totals of the call.
__global__ void __launch_bounds__(256) bases_kernel(Geo g, const uint32_t* __restrict__ rowtotal, uint32_t* __restrict__ rowbase,
                                                    unsigned long long* __restrict__ chunktotal,
                                                    unsigned long long* __restrict__ chunkbase, uint32_t* __restrict__ hso,
                                                    unsigned long long* __restrict__ counter, uint32_t* __restrict__ sync) {
    __shared__ unsigned long long warp_part[32];
    __shared__ uint32_t last_flag;
    const uint32_t lane = threadIdx.x & 31u, n = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (n < g.n_chunks) {
        const uint32_t* in = rowtotal + (size_t)n * g.H;
        uint32_t* out = rowbase + (size_t)n * g.H;
        unsigned long long carry = 0ull;
        for (uint32_t y0 = 0u; y0 < g.H; y0 += 128u) {   // four rows per lane and round
            unsigned long long v[4], sum = 0ull;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t y = y0 + lane * 4u + (uint32_t)j;
                v[j] = y < g.H ? (unsigned long long)in[y] : 0ull;
                sum += v[j];
            }
            unsigned long long incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long u = __shfl_up_sync(kFull, incl, off);
                if (lane >= (uint32_t)off) incl += u;
            }
            unsigned long long at = carry + incl - sum;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t y = y0 + lane * 4u + (uint32_t)j;
                if (y < g.H) out[y] = (uint32_t)at;   // (garbage if the chunk total overflows 32 bits; the host checks it)
                at += v[j];
            }
            carry += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0u) {
            chunktotal[n] = carry;
            hso[(size_t)n * ((size_t)g.W * g.H + 1u) + (size_t)g.W * g.H] = (uint32_t)carry;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0u) {
        __threadfence();
        last_flag = atomicAdd(&sync[0], 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!last_flag) return;
    __threadfence();
    const unsigned long long all = cta_exclusive_scan(chunktotal, chunkbase, g.n_chunks, warp_part);
    if (threadIdx.x == 0u) {
        chunkbase[g.n_chunks] = all;
        chunktotal[g.n_chunks] = *counter;
        *counter = 0ull;
        sync[0] = 0u;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// emit: producer warps (vertical window counts into a shared-memory ring), consumers = one warp per row
// ------------------------------------------------------------------------------------------------------------------
// FW = bits per vertical window count in the ring: 8 while 2r+1 <= 255, else 16 (2r+1 <= 511; the horizontal window
// counts then need 32 bits: (2r+1)^2 exceeds 65535 from r = 128 on).
__host__ __device__ constexpr uint32_t emit_sbuf_stride(int K, int FW = 8) {
    return FW == 8 ? 64u * (uint32_t)K + (K <= 2 ? 4u : 2u * (uint32_t)K) : 128u * (uint32_t)K + 4u * (uint32_t)K;
}
__host__ __device__ constexpr uint32_t emit_act_cap(int K) { return 32u * (uint32_t)K + 32u; }
__host__ __device__ constexpr int emit_acc_regs(int K, int FW) { return FW == 8 ? (K + 1) / 2 : K; }

// The K counts of one count vector a lane owns (compact ids lane*K .. lane*K+K-1): 8-bit counts widened to 16-bit
// pairs, 16-bit counts to one 32-bit word each
template <int K, int FW>
__device__ __forceinline__ void load_counts(const uint8_t* p, uint32_t (&v)[emit_acc_regs(K, FW)]) {
    if (FW == 8) {
        if (K == 1) {
            v[0] = *p;
        } else if (K == 2) {
            const uint32_t t = *reinterpret_cast<const uint16_t*>(p);
            v[0] = __byte_perm(t, 0u, 0x4140);
        } else if (K == 4) {
            const uint32_t t = *reinterpret_cast<const uint32_t*>(p);
            v[0] = __byte_perm(t, 0u, 0x4140);
            v[1 % emit_acc_regs(K, FW)] = __byte_perm(t, 0u, 0x4342);
        } else {
            const uint2 t = *reinterpret_cast<const uint2*>(p);
            v[0] = __byte_perm(t.x, 0u, 0x4140);
            v[1 % emit_acc_regs(K, FW)] = __byte_perm(t.x, 0u, 0x4342);
            v[2 % emit_acc_regs(K, FW)] = __byte_perm(t.y, 0u, 0x4140);
            v[3 % emit_acc_regs(K, FW)] = __byte_perm(t.y, 0u, 0x4342);
        }
    } else {
        if (K == 1) {
            v[0] = *reinterpret_cast<const uint16_t*>(p);
        } else if (K == 2) {
            const uint32_t t = *reinterpret_cast<const uint32_t*>(p);
            v[0] = t & 0xFFFFu;
            v[1 % K] = t >> 16;
        } else {
#pragma unroll
            for (int q = 0; q < K / 4; q++) {
                const uint2 t = reinterpret_cast<const uint2*>(p)[q];
                v[(4 * q) % K] = t.x & 0xFFFFu;
                v[(4 * q + 1) % K] = t.x >> 16;
                v[(4 * q + 2) % K] = t.y & 0xFFFFu;
                v[(4 * q + 3) % K] = t.y >> 16;
            }
        }
    }
}

template <int K, int FW>
__device__ __forceinline__ void store_sums(uint8_t* p, const uint32_t (&v)[emit_acc_regs(K, FW)]) {
    constexpr int N = emit_acc_regs(K, FW);
    if (N == 1) {
        if (FW == 8 && K == 1) *reinterpret_cast<uint16_t*>(p) = (uint16_t)v[0];
        else *reinterpret_cast<uint32_t*>(p) = v[0];
    } else if (N == 2) {
        *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1 % N]);
    } else {
#pragma unroll
        for (int q = 0; q < N / 4; q++)
            reinterpret_cast<uint4*>(p)[q] = make_uint4(v[(4 * q) % N], v[(4 * q + 1) % N], v[(4 * q + 2) % N], v[(4 * q + 3) % N]);
    }
}

// n columns of the dense horizontal slide: srun += counts(column c) - counts(column c - span); the window counts of a
// pixel (c >= 2r) go to the staging rows. OUT: the leaving column exists; ST: store.
template <int K, int FW, bool OUT, bool ST>
__device__ __forceinline__ void slide_cols(const uint8_t* pin, const uint8_t* pout, uint8_t* ps, uint32_t n,
                                           uint32_t (&srun)[emit_acc_regs(K, FW)]) {
    constexpr int CS = 32 * K * FW / 8, SR = emit_acc_regs(K, FW);
    constexpr uint32_t SS = emit_sbuf_stride(K, FW);
    constexpr int UJ = SR <= 4 ? 4 : 2;
    while (n >= (uint32_t)UJ) {
        uint32_t a[UJ][SR], o[UJ][SR];
#pragma unroll
        for (int j = 0; j < UJ; j++) {
            load_counts<K, FW>(pin + j * CS, a[j]);
            if (OUT) load_counts<K, FW>(pout + j * CS, o[j]);
        }
#pragma unroll
        for (int j = 0; j < UJ; j++) {
#pragma unroll
            for (int i = 0; i < SR; i++) srun[i] = OUT ? srun[i] + a[j][i] - o[j][i] : srun[i] + a[j][i];
            if (ST) store_sums<K, FW>(ps + j * SS, srun);
        }
        pin += UJ * CS;
        pout += UJ * CS;
        ps += UJ * SS;
        n -= (uint32_t)UJ;
    }
    while (n) {
        uint32_t a[SR], o[SR];
        load_counts<K, FW>(pin, a);
        if (OUT) load_counts<K, FW>(pout, o);
#pragma unroll
        for (int i = 0; i < SR; i++) srun[i] = OUT ? srun[i] + a[i] - o[i] : srun[i] + a[i];
        if (ST) store_sums<K, FW>(ps, srun);
        pin += CS;
        pout += CS;
        ps += SS;
        n--;
    }
}

// window count of compact id `id` in a staging row
template <int FW>
__device__ __forceinline__ uint32_t staged_count(uint32_t row_addr, uint32_t id) {
    uint32_t c;
    if (FW == 8) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(c) : "r"(row_addr + id * 2u));
    else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(row_addr + id * 4u));
    return c;
}

// Work items of an emit CTA. An item = the pixels [X0, X1) of the TY rows of one row tile of one chunk.
//  * one item per CTA (g.persist == 0): blockIdx = (tile, chunk, column segment), as planned on the host;
//  * persistent CTAs (g.persist == 1, one or a few CTAs per SM, grid.x = G): CTA i takes the tiles i, i + G, i + 2G, ...
//    of the flat tile list (static: every warp of the CTA derives the same sequence on its own, no hand-over). Its FIRST
//    tile is cut at a column that grows with i -- pixels [cut, W) are emitted first, pixels [0, cut) last -- so that the
//    CTAs of a launch run out of phase for its whole duration: a tile starts with 2r columns that only fill the window
//    and write nothing, and CTAs marching in lockstep would all leave the HBM idle at the same moments (measured on a
//    32-chunk batch: 7 waves x ~10 us of 790 us). Every CTA still does the same amount of work.
struct EmitItem {
    uint32_t n_chunk, tile, X0, X1;
};
__device__ __forceinline__ uint32_t emit_cut(const Geo& g) {   // first-tile cut of this CTA (a multiple of 16 pixels)
    if (!(g.persist & 1u) || gridDim.x < 2u) return 0u;
    return ((blockIdx.x * (g.W / 16u)) / gridDim.x) * 16u;
}
__device__ __forceinline__ uint32_t emit_item_count(const Geo& g) {
    if (!g.persist) return 1u;
    const uint32_t n_tiles = g.T * g.emit_chunks;
    return (n_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x + (emit_cut(g) ? 1u : 0u);
}
__device__ __forceinline__ EmitItem emit_item(const Geo& g, uint32_t k, uint32_t n_items) {
    EmitItem it;
    if (!g.persist) {
        it.n_chunk = g.emit_chunk0 + blockIdx.y;
        it.tile = blockIdx.x;
        it.X0 = blockIdx.z * g.cseg_px;
        it.X1 = min(g.W, it.X0 + g.cseg_px);
        return it;
    }
    const uint32_t cut = emit_cut(g);
    const bool tail = cut && k + 1u == n_items;          // the rest of the first tile comes last
    const uint32_t idx = blockIdx.x + (tail ? 0u : k) * gridDim.x;
    it.n_chunk = idx / g.T;
    it.tile = idx - it.n_chunk * g.T;
    it.n_chunk += g.emit_chunk0;
    it.X0 = (k == 0u) ? cut : 0u;
    it.X1 = tail ? cut : g.W;
    return it;
}

// shared memory of an emit CTA:
//   cring[TY][R][32K]  u8 / u16  vertical window counts (ring of R = 2r+1 + 16*stages columns per row), by compact id
//   sbuf[TY][16][SS]   u16 / u32 horizontal window counts of the batch being emitted, SS bytes per pixel
//   act[TY][32K+32]    8 B  the events overlapping the pixels being emitted, in list order

// (K = 1 is compiled for 768 threads, i.e. at most 80 registers: its small-call plans (8 rows + 4 producer warps = 384
// threads, < 76 KB of shared memory) then fit two CTAs per SM; at 86 registers a single 1024 x 1024 chunk lost 38 %)
template <int K, int FW>
__global__ void __launch_bounds__(K == 1 ? 768 : 640, 1)
    emit_kernel(Geo g, const uint16_t* __restrict__ cmap, const uint8_t* __restrict__ base,
                const uint2* __restrict__ pool, const uint2* __restrict__ rowinfo, const uint32_t* __restrict__ rowbase,
                const uint64_t* __restrict__ chunkbase, uint2* __restrict__ bins, uint32_t* __restrict__ hso) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int CS = 32 * K * FW / 8;            // bytes of one count vector
    constexpr int NB = kMarchNB;
    constexpr int KE = K + 1;                      // register sets of the active list
    constexpr uint32_t SS = emit_sbuf_stride(K, FW);
    constexpr uint32_t ACAP = emit_act_cap(K);
    constexpr int SR = emit_acc_regs(K, FW);
    // An item (see EmitItem) emits the pixels [X0, X1) of its rows. Its columns are the item's own window span
    // [X0, X1 + 2r): the first 2r of them only fill the window, exactly like the first 2r columns of a whole row; PW and
    // column indices below are local to the item. Batches and ring slots are counted through all items of the CTA
    // (gb0, ring0), so the barriers and the ring simply carry on from one item to the next: no CTA-wide barrier
    // between items, the producers start on the next item while the consumers finish the current one.
    const uint32_t TY = g.TY, R = g.R, span = g.span, two_r = 2u * g.r, stages = g.stages;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t NP = g.producers;
    const uint32_t n_my_items = emit_item_count(g);
    uint32_t total_batches = 0u;   // of all items of this CTA
    for (uint32_t k = 0u; k < n_my_items; k++) {
        const EmitItem it = emit_item(g, k, n_my_items);
        total_batches += (it.X1 - it.X0 + two_r + NB - 1u) / NB;
    }
    uint32_t gb0 = 0u, ring0 = 0u;   // batches before the current item; ring slot of the current item's column 0

    uint8_t* cring = smem;
    uint8_t* sbuf_all = smem + (size_t)TY * R * CS;
    uint2* act_all = reinterpret_cast<uint2*>(sbuf_all + (size_t)TY * NB * SS);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(act_all + (size_t)TY * ACAP);
    uint64_t* empty_bar = full_bar + 8;
    uint8_t* pscr_all = reinterpret_cast<uint8_t*>(empty_bar + 8);  // [producers][2][CPP][16] u16
    if (threadIdx.x < stages) {
        mbar_init(&full_bar[threadIdx.x], (uint32_t)(NB / (32 / (CS / 16))));  // producer passes per batch
        mbar_init(&empty_bar[threadIdx.x], TY);
    }
    __syncthreads();

    if (warp >= TY) {
        // =============================== producer warps ===============================
        // Work item t = (batch b, pass q): CPP columns x the tile's rows. The samples entering / leaving the vertical
        // window (rows y0+i+span and y0+i) are fetched as one row segment per lane (lanes 0-15 entering, 16-31
        // leaving), a whole item ahead of their use, and handed to the column lanes through a transposed scratch.
        constexpr int LPC = CS / 16;     // lanes per column, 16 bytes of counters each
        constexpr int CPP = 32 / LPC;    // columns per pass
        constexpr int PPB = NB / CPP;    // passes per batch
        constexpr int SW = CPP > 1 ? CPP / 2 : 1;  // 32-bit words of one row segment (one 16-bit sample when CPP = 1)
        const uint32_t part = lane % LPC, colq = lane / LPC;
        const uint32_t bit0 = part * 128u;
        const uint32_t half = lane >> 4, lrow = lane & 15u;
        uint16_t* scr = reinterpret_cast<uint16_t*>(pscr_all) + (size_t)(warp - TY) * (2 * CPP * 16);
        for (uint32_t item_k = 0u; item_k < n_my_items; item_k++) {
        const EmitItem item = emit_item(g, item_k, n_my_items);
        const uint32_t X0 = item.X0, PW = (item.X1 - item.X0) + two_r, n_chunk = item.n_chunk, tile = item.tile;
        const uint32_t y0 = tile * TY, n_batches = (PW + NB - 1u) / NB;
        const uint16_t* cm = cmap + (size_t)n_chunk * g.ids_chunk_stride;
        const uint32_t tile_rows = min(TY, g.H - y0);
        const uint16_t* seg_row = cm + (size_t)min(half ? y0 + lrow : y0 + lrow + span, g.PH - 1u) * g.ids_row_stride + X0;
        const uint8_t* base_tile = base + (((size_t)n_chunk * g.T + tile) * g.PW + X0) * CS + part * 16u;
        const uint32_t n_items = n_batches * PPB;
        uint32_t seg_n[SW];
        uint4 v_n = make_uint4(0u, 0u, 0u, 0u);
        auto fetch = [&](uint32_t t) {
            const uint32_t c0 = (t / PPB) * NB + (t % PPB) * CPP;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(seg_row + c0);
            if (SW == 8) {
                const uint4 a = reinterpret_cast<const uint4*>(src)[0], b2 = reinterpret_cast<const uint4*>(src)[1];
                seg_n[0] = a.x, seg_n[1] = a.y, seg_n[2] = a.z, seg_n[3] = a.w;
                seg_n[4 % SW] = b2.x, seg_n[5 % SW] = b2.y, seg_n[6 % SW] = b2.z, seg_n[7 % SW] = b2.w;
            } else if (SW == 4) {
                const uint4 a = reinterpret_cast<const uint4*>(src)[0];
                seg_n[0] = a.x, seg_n[1 % SW] = a.y, seg_n[2 % SW] = a.z, seg_n[3 % SW] = a.w;
            } else if (SW == 2) {
                const uint2 a = reinterpret_cast<const uint2*>(src)[0];
                seg_n[0] = a.x, seg_n[1 % SW] = a.y;
            } else if (CPP == 2) {
                seg_n[0] = src[0];
            } else {
                seg_n[0] = seg_row[c0];
            }
            const uint32_t c = min(c0 + colq, PW - 1u);
            v_n = *reinterpret_cast<const uint4*>(base_tile + (size_t)c * CS);
        };
        // Passes go to the producer warps round robin by their GLOBAL index (counted through all items of the CTA), not
        // by their index inside the item: a warp's consecutive batches then stay NP / PPB <= stages apart. The parity
        // wait below can only tell "the batch `stages` back is consumed" from "not yet"; a warp that jumped further
        // ahead (items whose batch count is no multiple of the warp count would let it) would see the parity of an
        // older phase, take it for its own and overwrite ring columns still in use.
        uint32_t t = (warp - TY + NP - (gb0 * PPB) % NP) % NP;
        if (t < n_items) fetch(t);
        for (; t < n_items; t += NP) {
            const uint32_t b = t / PPB, pass = t % PPB;
            // (batch -> ring stage and barrier phase without a division: stages is a run-time value; q = gb / stages)
            const uint32_t gb = gb0 + b, q = __umulhi(gb, g.stages_magic), s = gb - q * stages, cb = b * NB;
            const uint32_t cu = pass * CPP + colq;
            const bool live = cb + cu < PW;
            uint32_t slot = (ring0 + cb) % R + cu;
            if (slot >= R) slot -= R;
            uint4 v = v_n;
            // transposed scratch: scr[half][column][row]
#pragma unroll
            for (int j = 0; j < CPP; j++)
                scr[(half * CPP + j) * 16 + lrow] = (uint16_t)(seg_n[j / 2] >> (16 * (j & 1)));
            if (t + NP < n_items) fetch(t + NP);
            __syncwarp();
            uint32_t wa[8], wo[8];
            {
                const uint4* pa = reinterpret_cast<const uint4*>(scr + colq * 16);
                const uint4* po = reinterpret_cast<const uint4*>(scr + (CPP + colq) * 16);
                const uint4 a0 = pa[0], a1 = pa[1], o0 = po[0], o1 = po[1];
                wa[0] = a0.x, wa[1] = a0.y, wa[2] = a0.z, wa[3] = a0.w, wa[4] = a1.x, wa[5] = a1.y, wa[6] = a1.z, wa[7] = a1.w;
                wo[0] = o0.x, wo[1] = o0.y, wo[2] = o0.z, wo[3] = o0.w, wo[4] = o1.x, wo[5] = o1.y, wo[6] = o1.z, wo[7] = o1.w;
            }
            if (gb >= stages) mbar_wait(&empty_bar[s], (q - 1u) & 1u, 0x10000000u | (item_k << 16) | gb);  // batch gb - stages is consumed
            uint8_t* out = cring + (size_t)slot * CS + part * 16u;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if ((uint32_t)i < tile_rows) {
                    if (live) *reinterpret_cast<uint4*>(out) = v;
                    out += (size_t)R * CS;
                    // counter of compact id a sits at bit FW*a of the column vector; a shift by >= 32 (or "negative",
                    // i.e. huge) yields 0 with PTX shl, so every word only sees its own counters
                    const uint32_t sa = (i & 1) ? wa[i / 2] >> 16 : wa[i / 2] & 0xFFFFu;
                    const uint32_t so = (i & 1) ? wo[i / 2] >> 16 : wo[i / 2] & 0xFFFFu;
                    const uint32_t ba = sa * (uint32_t)FW - bit0, bo = so * (uint32_t)FW - bit0;
                    v.x += shl_clamp(1u, ba) - shl_clamp(1u, bo);
                    v.y += shl_clamp(1u, ba - 32u) - shl_clamp(1u, bo - 32u);
                    v.z += shl_clamp(1u, ba - 64u) - shl_clamp(1u, bo - 64u);
                    v.w += shl_clamp(1u, ba - 96u) - shl_clamp(1u, bo - 96u);
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0u) mbar_arrive(&full_bar[s]);
        }
        gb0 += n_batches;
        ring0 = (ring0 + n_batches * NB) % R;
        }
        return;
    }

    // =============================== consumer warps ===============================
    uint8_t* crow = cring + (size_t)warp * R * CS;
    uint8_t* sb = sbuf_all + (size_t)warp * NB * SS;
    uint2* act = act_all + (size_t)warp * ACAP;
    for (uint32_t item_k = 0u; item_k < n_my_items; item_k++) {
    const EmitItem item = emit_item(g, item_k, n_my_items);
    const uint32_t X0 = item.X0, PW = (item.X1 - item.X0) + two_r, n_chunk = item.n_chunk;
    const uint32_t n_batches = (PW + NB - 1u) / NB;
    const uint32_t y = item.tile * TY + warp;
    // a call that runs ahead of the host's size checks must stay inside the buffers it was given (the host then sees the
    // totals, grows the buffers and runs the call again): such a chunk is walked through without output
    const bool row_active = y < g.H && chunkbase[n_chunk + 1u] <= g.bins_cap;
    const uint2 ri = row_active ? rowinfo[(size_t)n_chunk * g.H + y] : make_uint2(0u, 0u);
    const uint2* ev = pool + ri.x;
    const uint32_t E = (unsigned long long)ri.x + ri.y <= g.pool_cap ? ri.y : 0u;
    uint32_t off = row_active ? rowbase[(size_t)n_chunk * g.H + y] : 0u;  // offset of the next pixel's first bin
    if (X0 != 0u && row_active) {
        // bins of the row's pixels before this segment: every event contributes its overlap with [0, X0)
        uint32_t before = 0u;
        for (uint32_t i = lane; i < E; i += 32u) {
            const uint32_t span_px = ev[i].y;
            before += min(span_px >> 16, X0) - min(span_px & 0xFFFFu, X0);
        }
        off += __reduce_add_sync(kFull, before);
    }
    uint2* dst = bins + (row_active ? (size_t)chunkbase[n_chunk] + off : (size_t)0);
    uint32_t* hso_row = hso + (size_t)n_chunk * ((size_t)g.W * g.H + 1u) + (size_t)(row_active ? y : 0u) * g.W;
    const float inv = g.inv_total;

    uint32_t srun[SR];
#pragma unroll
    for (int i = 0; i < SR; i++) srun[i] = 0u;
    uint32_t rx[KE], ry[KE];   // the active list: record k*32 + lane, or the packed view in set 0
#pragma unroll
    for (int k = 0; k < KE; k++) {
        rx[k] = 0u;
        ry[k] = 0xFFFFu;
    }
    uint32_t Ea = 0u, e_lo = 0u, valid_until = 0u;
    bool clean = false;
    uint32_t pk_n = 0u, pk_slot = 0u, pk_entry = 0u;  // packed view (Ea <= 16): lane = slot * Ea + entry

    uint32_t in_slot = ring0, out_slot = (ring0 + R - span % R) % R;
    for (uint32_t b = 0u; b < n_batches; b++) {
        const uint32_t gb = gb0 + b, q = __umulhi(gb, g.stages_magic), s = gb - q * stages, cb = b * NB;
        const uint32_t ce = min(cb + (uint32_t)NB, PW);
        mbar_wait(&full_bar[s], q & 1u, (item_k << 16) | gb);
        if (row_active) {
            // ---- horizontal window counts of the batch's columns, all values at once (lane = K consecutive ids) ----
            if (cb >= span && ce - cb == (uint32_t)NB && in_slot + NB <= R && out_slot + NB <= R) {
                // the steady state: a whole batch, no ring wrap, every column has a leaving partner and a pixel
                slide_cols<K, FW, true, true>(crow + in_slot * CS + lane * (K * FW / 8), crow + out_slot * CS + lane * (K * FW / 8),
                                              sb + lane * (FW == 8 ? 2 * K : 4 * K), (uint32_t)NB, srun);
                in_slot += NB;
                if (in_slot >= R) in_slot -= R;
                out_slot += NB;
                if (out_slot >= R) out_slot -= R;
            } else
            for (uint32_t c = cb; c < ce;) {
                // a segment: no ring wrap, the leaving column exists or not, pixels are due or not
                const bool has_out = c >= span, st = c >= two_r;
                uint32_t seg_end = min(ce, c + (R - in_slot));
                seg_end = has_out ? min(seg_end, c + (R - out_slot)) : min(seg_end, span);
                if (!st) seg_end = min(seg_end, two_r);
                const uint32_t n = seg_end - c;
                const uint8_t* pin = crow + in_slot * CS + lane * (K * FW / 8);
                const uint8_t* pout = crow + out_slot * CS + lane * (K * FW / 8);
                uint8_t* ps = sb + (c - cb) * SS + lane * (FW == 8 ? 2 * K : 4 * K);
                if (has_out) slide_cols<K, FW, true, true>(pin, pout, ps, n, srun);
                else if (st) slide_cols<K, FW, false, true>(pin, pout, ps, n, srun);
                else slide_cols<K, FW, false, false>(pin, pout, ps, n, srun);
                in_slot += n;
                if (in_slot >= R) in_slot -= R;
                out_slot += n;
                if (out_slot >= R) out_slot -= R;
                c = seg_end;
            }
            __syncwarp();
            // ---- the pixels of this batch: x = c - 2r ----
            if (ce > two_r) {
                const uint32_t xa0 = X0 + (cb > two_r ? cb - two_r : 0u), xe0 = X0 + ce - two_r;   // pixels of the row
                uint32_t xa = xa0, xe = xe0;
                bool single = false;
                while (xa < xe0) {
                    if (xe > valid_until || single) {
                        // ---- rebuild the active list for the pixels [xa, xe) ----
                        uint32_t nact = 0u, vmin = 0xFFFFFFFFu;
                        bool dirty = false;
                        __syncwarp();
                        for (uint32_t eb = e_lo; eb < E; eb += 32u) {
                            const uint32_t idx = eb + lane;
                            uint2 rec = make_uint2(0u, 0xFFFFu);
                            if (idx < E) rec = ev[idx];
                            const uint32_t xb = rec.y & 0xFFFFu, xd = rec.y >> 16;
                            const bool in_range = idx < E;
                            const bool fut = in_range && xb >= xe;
                            const bool ov = in_range && xb < xe && xd > xa;
                            const unsigned bm = __ballot_sync(kFull, ov);
                            if (ov) {
                                const uint32_t at = nact + (uint32_t)__popc(bm & lanemask_lt());
                                if (at < ACAP) act[at] = rec;
                                dirty |= xb > xa || xd < xe;
                                vmin = min(vmin, xd);
                            }
                            if (fut) vmin = min(vmin, xb);
                            nact += (uint32_t)__popc(bm);
                            const unsigned deadm = __ballot_sync(kFull, in_range && xd <= xa);
                            if (eb == e_lo && deadm == kFull) e_lo += 32u;
                            if (__ballot_sync(kFull, fut || !in_range) == kFull) break;
                        }
                        __syncwarp();
                        if (nact > ACAP && xe - xa > 1u) {  // too many for one pass: pixel by pixel (<= one bin per value)
                            single = true;
                            xe = xa + 1u;
                            continue;
                        }
                        Ea = nact;
                        clean = !__any_sync(kFull, dirty);
                        valid_until = clean ? __reduce_min_sync(kFull, vmin) : xe;
                        if (single) valid_until = xe;
                        pk_n = 0u;
                        if (Ea >= 1u && Ea <= 16u) {
                            pk_n = 32u / Ea;
                            pk_slot = lane / Ea;
                            pk_entry = lane - pk_slot * Ea;
                            const uint2 rec = act[pk_entry];
                            rx[0] = rec.x;
                            ry[0] = pk_slot < pk_n ? rec.y : 0xFFFFu;
                        } else {
#pragma unroll
                            for (int k = 0; k < KE; k++) {
                                const uint32_t idx = k * 32 + lane;
                                uint2 rec = make_uint2(0u, 0xFFFFu);
                                if (idx < Ea) rec = act[idx];
                                rx[k] = rec.x;
                                ry[k] = rec.y;
                            }
                        }
                    }
                    // ---- emit the pixels [xa, xe) ----
                    const uint8_t* sp = sb + (size_t)(xa - X0 + two_r - cb) * SS;  // window counts of pixel xa
                    if (pk_n) {
                        // lane = slot * Ea + entry: pk_n pixels per pass; positions by one ballot + popc
                        const uint32_t xb = ry[0] & 0xFFFFu, lim = min(ry[0] >> 16, xe), itm = rx[0] >> 16;
                        const uint32_t len = lim > xb ? lim - xb : 0u;             // alive <=> (x - xb) < len
                        const uint32_t hlim = (pk_entry == 0u && pk_slot < pk_n) ? xe : 0u;  // this lane stores the offset of x
                        uint32_t x = xa + pk_slot;
                        uint32_t sa = smem_addr(sp) + pk_slot * SS;
                        const uint32_t vid = rx[0] & 0xFFFFu;
                        const uint32_t sa_step = pk_n * SS;
                        uint32_t* hp = hso_row + x;
                        uint32_t xp = xa;
                        if (clean) {
                            // every listed event is alive for every pixel of the range: a full pass writes pk_n * Ea
                            // bins, lane = position, no vote needed
                            const uint32_t full = pk_n * Ea;
                            const bool writer = lane < full, marks = pk_entry == 0u && pk_slot < pk_n;
                            const uint32_t hoff = pk_slot * Ea;
                            for (; xp + pk_n <= xe; xp += pk_n) {
                                if (writer) {
                                    const uint32_t cnt = staged_count<FW>(sa, vid);
                                    dst[lane] = make_uint2(itm, __float_as_uint(__fmul_rn(__uint2float_rn(cnt), inv)));
                                }
                                if (marks) *hp = off + hoff;
                                dst += full;
                                off += full;
                                x += pk_n;
                                sa += sa_step;
                                hp += pk_n;
                            }
                        }
                        for (; xp < xe; xp += pk_n) {
                            const bool alive = x - xb < len;
                            const unsigned bm = __ballot_sync(kFull, alive);
                            const uint32_t pos = (uint32_t)__popc(bm & lanemask_lt());
                            if (alive) {
                                const uint32_t cnt = staged_count<FW>(sa, vid);
                                dst[pos] = make_uint2(itm, __float_as_uint(__fmul_rn(__uint2float_rn(cnt), inv)));
                            }
                            if (x < hlim) *hp = off + pos;
                            const uint32_t nb = (uint32_t)__popc(bm);
                            dst += nb;
                            off += nb;
                            x += pk_n;
                            sa += sa_step;
                            hp += pk_n;
                        }
                    } else if (clean) {
#pragma unroll 4
                        for (uint32_t x = xa; x < xe; x++) {
#pragma unroll
                            for (int k = 0; k < KE; k++) {
                                if ((uint32_t)(k * 32) < Ea) {
                                    const uint32_t idx = k * 32 + lane;
                                    if (idx < Ea) {
                                        const uint32_t cnt = staged_count<FW>(smem_addr(sp), rx[k] & 0xFFFFu);
                                        dst[idx] = make_uint2(rx[k] >> 16, __float_as_uint(__fmul_rn(__uint2float_rn(cnt), inv)));
                                    }
                                }
                            }
                            if (lane == 0u) hso_row[x] = off;
                            dst += Ea;
                            off += Ea;
                            sp += SS;
                        }
                    } else {
                        for (uint32_t x = xa; x < xe; x++) {
                            uint32_t at = 0u;
#pragma unroll
                            for (int k = 0; k < KE; k++) {
                                if ((uint32_t)(k * 32) < Ea) {
                                    const uint32_t xb = ry[k] & 0xFFFFu, xd = ry[k] >> 16;
                                    const bool alive = xb <= x && x < xd;
                                    const unsigned bm = __ballot_sync(kFull, alive);
                                    if (alive) {
                                        const uint32_t cnt = staged_count<FW>(smem_addr(sp), rx[k] & 0xFFFFu);
                                        dst[at + (uint32_t)__popc(bm & lanemask_lt())] =
                                            make_uint2(rx[k] >> 16, __float_as_uint(__fmul_rn(__uint2float_rn(cnt), inv)));
                                    }
                                    at += (uint32_t)__popc(bm);
                                }
                            }
                            if (lane == 0u) hso_row[x] = off;
                            dst += at;
                            off += at;
                            sp += SS;
                        }
                    }
                    xa = xe;
                    xe = single ? min(xa + 1u, xe0) : xe0;
                }
            }
        } else {
            const uint32_t adv = ce - cb;
            in_slot += adv;
            if (in_slot >= R) in_slot -= R;
            out_slot += adv;
            if (out_slot >= R) out_slot -= R;
        }
        __syncwarp();  // the batch's window counts and the ring columns it read are free again
        if (gb + stages < total_batches && lane == 0u) mbar_arrive(&empty_bar[s]);
    }
    gb0 += n_batches;
    ring0 = (ring0 + n_batches * NB) % R;
    }
}

}  // namespace shf
