/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of SuperTerrain+'s single histogram filter.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library,
 * and only as the checker or the timed CPU baseline. The product path (superterrainplus_b200/csrc) never links it.
 *
 * Parity status: PINNED. tests/test_oracle.py checks this restatement against
 *   (1) the reference's own golden vector (SuperTest+/SuperAlgorithm+/STPTestHistogram.cpp:44-94, pixels 0/8/15),
 *   (2) the reference's own filter compiled from /root/reference into oracle/_ref/libshf_ref.so (bit-exact on
 *       items, offsets and weight bits over randomised and adversarial maps), with the resulting vectors committed
 *       under tests/golden/ so the check also runs where /root/reference is absent.
 *
 * Written from the behaviour of /root/reference/SuperTerrain+/SuperAlgorithm+/Host/Private/STPSingleHistogramFilter.cpp
 * (cited as SHF.cpp below); it shares no code with it. Structure here: a fixed 65536-entry slot table and one ordered
 * run of (item,count) pairs per window, first a column sweep into a column-major scratch volume, then a row sweep.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint16_t item;
    float weight;
} shf_oracle_bin; /* layout of STPSingleHistogram::STPBin, SuperAlgorithm+Host/STPSingleHistogram.hpp:23-37 */

#define SLOT_NONE (-1)

/* An insertion-ordered multiset of samples: the behaviour of STPAccumulator (SHF.cpp:375-475). */
typedef struct {
    int32_t* slot_of;   /* sample value -> position in item[]/count[], or SLOT_NONE */
    uint16_t* item;
    uint32_t* count;
    uint32_t used;
    uint32_t cap;
} ordered_tally;

static int tally_init(ordered_tally* t) {
    t->slot_of = (int32_t*)malloc(65536u * sizeof(int32_t));
    t->cap = 64u;
    t->item = (uint16_t*)malloc(t->cap * sizeof(uint16_t));
    t->count = (uint32_t*)malloc(t->cap * sizeof(uint32_t));
    t->used = 0u;
    if (!t->slot_of || !t->item || !t->count) return -1;
    for (uint32_t i = 0u; i < 65536u; i++) t->slot_of[i] = SLOT_NONE;
    return 0;
}

static void tally_free(ordered_tally* t) {
    free(t->slot_of);
    free(t->item);
    free(t->count);
}

/* forget everything; O(used) because only live samples have a slot */
static void tally_reset(ordered_tally* t) {
    for (uint32_t i = 0u; i < t->used; i++) t->slot_of[t->item[i]] = SLOT_NONE;
    t->used = 0u;
}

/* SHF.cpp:400-420: a sample seen for the first time goes to the END of the run, otherwise its count grows */
static int tally_add(ordered_tally* t, uint16_t sample, uint32_t amount) {
    const int32_t pos = t->slot_of[sample];
    if (pos != SLOT_NONE) {
        t->count[pos] += amount;
        return 0;
    }
    if (t->used == t->cap) {
        t->cap *= 2u;
        t->item = (uint16_t*)realloc(t->item, t->cap * sizeof(uint16_t));
        t->count = (uint32_t*)realloc(t->count, t->cap * sizeof(uint32_t));
        if (!t->item || !t->count) return -1;
    }
    t->item[t->used] = sample;
    t->count[t->used] = amount;
    t->slot_of[sample] = (int32_t)t->used;
    t->used++;
    return 0;
}

/* SHF.cpp:430-450: when the count would drop to zero (count <= amount) the entry is removed and every later
 * entry moves one place towards the front, keeping relative order */
static void tally_sub(ordered_tally* t, uint16_t sample, uint32_t amount) {
    const int32_t pos = t->slot_of[sample];
    if (t->count[pos] > amount) {
        t->count[pos] -= amount;
        return;
    }
    for (uint32_t i = (uint32_t)pos + 1u; i < t->used; i++) {
        t->item[i - 1u] = t->item[i];
        t->count[i - 1u] = t->count[i];
        t->slot_of[t->item[i]] = (int32_t)(i - 1u);
    }
    t->slot_of[sample] = SLOT_NONE;
    t->used--;
}

/* growable (item,count) stream with one start offset per emitted window */
typedef struct {
    uint16_t* item;
    uint32_t* count;
    uint64_t* start;
    uint64_t n_pairs, cap_pairs;
    uint64_t n_starts, cap_starts;
} pair_stream;

static int stream_init(pair_stream* s, uint64_t windows) {
    s->cap_pairs = windows * 4u + 16u;
    s->cap_starts = windows + 1u;
    s->item = (uint16_t*)malloc(s->cap_pairs * sizeof(uint16_t));
    s->count = (uint32_t*)malloc(s->cap_pairs * sizeof(uint32_t));
    s->start = (uint64_t*)malloc(s->cap_starts * sizeof(uint64_t));
    s->n_pairs = s->n_starts = 0u;
    return (s->item && s->count && s->start) ? 0 : -1;
}

static void stream_free(pair_stream* s) {
    free(s->item);
    free(s->count);
    free(s->start);
}

/* append the whole current run as one window (SHF.cpp:483-506, un-normalised flavour) */
static int stream_emit(pair_stream* s, const ordered_tally* t) {
    if (s->n_pairs + t->used > s->cap_pairs) {
        while (s->n_pairs + t->used > s->cap_pairs) s->cap_pairs *= 2u;
        s->item = (uint16_t*)realloc(s->item, s->cap_pairs * sizeof(uint16_t));
        s->count = (uint32_t*)realloc(s->count, s->cap_pairs * sizeof(uint32_t));
        if (!s->item || !s->count) return -1;
    }
    s->start[s->n_starts++] = s->n_pairs;
    memcpy(s->item + s->n_pairs, t->item, t->used * sizeof(uint16_t));
    memcpy(s->count + s->n_pairs, t->count, t->used * sizeof(uint32_t));
    s->n_pairs += t->used;
    return 0;
}

/* status codes shared with the reference driver (oracle/ref_driver.cpp) */
enum { ORACLE_OK = 0, ORACLE_NUMERIC_DOMAIN = 1, ORACLE_NO_MEMORY = 5, ORACLE_OFFSET_OVERFLOW = 6 };

/*
 * One filter execution. `map` is the merged nearest-neighbour sample map, row-major with stride total_x.
 * On success *bins_out / *offsets_out are malloc'd (caller frees with shf_oracle_free) and hold the same values the
 * reference leaves in its STPFilterBuffer: offsets has W*H+1 entries, the last one the bin total (SHF.cpp:688-690).
 */
int shf_oracle_run(const uint16_t* map, uint32_t map_w, uint32_t map_h, uint32_t nn_x, uint32_t nn_y, uint32_t total_x,
                   uint32_t radius, shf_oracle_bin** bins_out, uint32_t** offsets_out, uint64_t* n_bins_out) {
    *bins_out = NULL;
    *offsets_out = NULL;
    *n_bins_out = 0u;
    /* SHF.cpp:874-880: radius positive, even, and not wider than the neighbour ring on either axis */
    if (radius == 0u || (radius & 1u) != 0u) return ORACLE_NUMERIC_DOMAIN;
    const uint32_t origin_x = map_w * (nn_x / 2u), origin_y = map_h * (nn_y / 2u);
    if (radius > origin_x || radius > origin_y) return ORACLE_NUMERIC_DOMAIN;

    const uint32_t span = 2u * radius + 1u;
    const uint32_t strip_w = map_w + 2u * radius; /* columns the row sweep needs: SHF.cpp:885-886 */
    const uint32_t first_col = origin_x - radius;
    int status = ORACLE_NO_MEMORY;

    ordered_tally tally;
    pair_stream columns; /* column-major: window (c, y) sits at c * map_h + y, like the reference's scratch */
    memset(&columns, 0, sizeof(columns));
    memset(&tally, 0, sizeof(tally));
    shf_oracle_bin* bins = NULL;
    uint32_t* offsets = NULL;
    if (tally_init(&tally) != 0) goto done;
    if (stream_init(&columns, (uint64_t)strip_w * map_h) != 0) goto done;

    /* column sweep (SHF.cpp:522-563): per column load rows [origin_y-r, origin_y+r] top to bottom, then for each
     * further centre row FIRST take in the row below, THEN drop the row above */
    for (uint32_t c = 0u; c < strip_w; c++) {
        const uint16_t* col = map + (size_t)(first_col + c);
        const size_t top = (size_t)(origin_y - radius);
        for (uint32_t k = 0u; k < span; k++)
            if (tally_add(&tally, col[(top + k) * total_x], 1u) != 0) goto done;
        if (stream_emit(&columns, &tally) != 0) goto done;
        for (uint32_t y = 1u; y < map_h; y++) {
            if (tally_add(&tally, col[(top + y + span - 1u) * total_x], 1u) != 0) goto done;
            tally_sub(&tally, col[(top + y - 1u) * total_x], 1u);
            if (stream_emit(&columns, &tally) != 0) goto done;
        }
        tally_reset(&tally);
    }
    columns.start[columns.n_starts] = columns.n_pairs; /* sentinel, SHF.cpp:688-690 */

    /* row sweep (SHF.cpp:608-679) over the column windows, emitting normalised bins row-major */
    {
        uint64_t cap = (uint64_t)map_w * map_h * 4u + 16u, used = 0u;
        bins = (shf_oracle_bin*)calloc(cap, sizeof(shf_oracle_bin));
        offsets = (uint32_t*)malloc(((size_t)map_w * map_h + 1u) * sizeof(uint32_t));
        if (!bins || !offsets) goto done;
        for (uint32_t y = 0u; y < map_h; y++) {
            for (uint32_t x = 0u; x < map_w; x++) {
                if (x == 0u) {
                    for (uint32_t c = 0u; c < span; c++) {
                        const uint64_t w = (uint64_t)c * map_h + y;
                        for (uint64_t i = columns.start[w]; i < columns.start[w + 1u]; i++)
                            if (tally_add(&tally, columns.item[i], columns.count[i]) != 0) goto done;
                    }
                } else {
                    const uint64_t in = (uint64_t)(x + span - 1u) * map_h + y, out = (uint64_t)(x - 1u) * map_h + y;
                    for (uint64_t i = columns.start[in]; i < columns.start[in + 1u]; i++)
                        if (tally_add(&tally, columns.item[i], columns.count[i]) != 0) goto done;
                    for (uint64_t i = columns.start[out]; i < columns.start[out + 1u]; i++)
                        tally_sub(&tally, columns.item[i], columns.count[i]);
                }
                if (used > 0xFFFFFFFFull) {
                    status = ORACLE_OFFSET_OVERFLOW;
                    goto done;
                }
                offsets[(size_t)y * map_w + x] = (uint32_t)used;
                if (used + tally.used > cap) {
                    while (used + tally.used > cap) cap *= 2u;
                    bins = (shf_oracle_bin*)realloc(bins, cap * sizeof(shf_oracle_bin));
                    if (!bins) goto done;
                }
                /* SHF.cpp:493-499: weight = count * (1.0f / float(sum of counts)), both in binary32 */
                uint32_t sum = 0u;
                for (uint32_t i = 0u; i < tally.used; i++) sum += tally.count[i];
                const float scale = 1.0f / (float)sum;
                for (uint32_t i = 0u; i < tally.used; i++) {
                    memset(&bins[used + i], 0, sizeof(shf_oracle_bin));
                    bins[used + i].item = tally.item[i];
                    bins[used + i].weight = (float)tally.count[i] * scale;
                }
                used += tally.used;
            }
            tally_reset(&tally);
        }
        if (used > 0xFFFFFFFFull) {
            status = ORACLE_OFFSET_OVERFLOW;
            goto done;
        }
        offsets[(size_t)map_w * map_h] = (uint32_t)used;
        *n_bins_out = used;
    }
    *bins_out = bins;
    *offsets_out = offsets;
    bins = NULL;
    offsets = NULL;
    status = ORACLE_OK;

done:
    free(bins);
    free(offsets);
    stream_free(&columns);
    tally_free(&tally);
    return status;
}

void shf_oracle_free(void* p) { free(p); }

unsigned shf_oracle_bin_stride(void) { return (unsigned)sizeof(shf_oracle_bin); }
