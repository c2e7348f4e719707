// TEST INFRASTRUCTURE (see vm_oracle.h): CPU restatement of the series assembly step of the query path.
//   mergeSortBlocks   app/vmselect/netstorage/netstorage.go:566-616 (+ sortBlocksHeap :684-730, binarySearchTimestamps :646,
//                     equalSamplesPrefix :622-644) on top of Go's container/heap (Init / Fix / Pop: sift-down from n/2-1,
//                     "down" prefers the right child only when it is strictly less, Fix = down else up)
//   DeduplicateSamples / needsDedup   lib/storage/dedup.go:30-100, :158-175
#include <cstring>
#include <vector>

#include "vm_oracle.h"

namespace {

struct SortBlock {
    const int64_t* ts;
    const double* vals;
    size_t n, next;
};

struct Heap {  // sortBlocksHeap over container/heap
    std::vector<SortBlock*> sbs;
    bool less(size_t i, size_t j) const { return sbs[i]->ts[sbs[i]->next] < sbs[j]->ts[sbs[j]->next]; }
    bool down(size_t i0, size_t n) {
        size_t i = i0;
        for (;;) {
            size_t j1 = 2 * i + 1;
            if (j1 >= n) break;
            size_t j = j1;
            size_t j2 = j1 + 1;
            if (j2 < n && less(j2, j1)) j = j2;
            if (!less(j, i)) break;
            std::swap(sbs[i], sbs[j]);
            i = j;
        }
        return i > i0;
    }
    void up(size_t j) {
        for (;;) {
            if (j == 0) break;
            size_t i = (j - 1) / 2;
            if (i == j || !less(j, i)) break;
            std::swap(sbs[i], sbs[j]);
            j = i;
        }
    }
    void init() {
        size_t n = sbs.size();
        for (size_t i = n / 2; i-- > 0;) down(i, n);
    }
    void fix(size_t i) {
        if (!down(i, sbs.size())) up(i);
    }
    void pop() {
        size_t n = sbs.size() - 1;
        std::swap(sbs[0], sbs[n]);
        down(0, n);
        sbs.pop_back();
    }
    SortBlock* next_block() {  // getNextBlock netstorage.go:689
        if (sbs.size() < 2) return nullptr;
        if (sbs.size() < 3) return sbs[1];
        SortBlock *a = sbs[1], *b = sbs[2];
        return a->ts[a->next] <= b->ts[b->next] ? a : b;
    }
};

size_t binary_search_timestamps(const int64_t* ts, size_t n, int64_t t) {  // netstorage.go:646
    if (n > 0 && ts[n - 1] <= t) return n;
    size_t i = 0, j = n;
    while (i < j) {
        size_t h = (i + j) >> 1;
        if (ts[h] <= t) i = h + 1;
        else j = h;
    }
    return i;
}

size_t equal_samples_prefix(const SortBlock* a, const SortBlock* b) {  // netstorage.go:622
    size_t na = a->n - a->next, nb = b->n - b->next, n = 0;
    while (n < na && n < nb && a->ts[a->next + n] == b->ts[b->next + n]) n++;
    if (n == 0) return 0;
    size_t m = 0;
    while (m < n) {
        uint64_t x, y;
        memcpy(&x, &a->vals[a->next + m], 8);
        memcpy(&y, &b->vals[b->next + m], 8);
        if (x != y) break;
        m++;
    }
    return m;
}

inline bool is_stale_nan(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return b == 0x7ff0000000000002ULL;
}

}  // namespace

extern "C" int vmo_needs_dedup(const int64_t* ts, size_t n, int64_t interval) {  // dedup.go:158
    if (n < 2 || interval <= 0) return 0;
    int64_t tsNext = ts[0] + interval - 1;
    tsNext -= tsNext % interval;
    for (size_t i = 1; i < n; i++) {
        if (ts[i] <= tsNext) return 1;
        tsNext += interval;
        if (tsNext < ts[i]) {
            tsNext = ts[i] + interval - 1;
            tsNext -= tsNext % interval;
        }
    }
    return 0;
}

extern "C" size_t vmo_deduplicate_samples(int64_t* ts, double* vals, size_t n, int64_t interval) {  // dedup.go:30
    if (!vmo_needs_dedup(ts, n, interval)) return n;
    int64_t tsNext = ts[0] + interval - 1;
    tsNext -= tsNext % interval;
    size_t o = 0;
    auto pick = [&](size_t j) {  // maximum among equal timestamps, never a staleness marker when something else exists
        int64_t tsPrev = ts[j];
        double vPrev = vals[j];
        while (j > 0 && ts[j - 1] == tsPrev) {
            j--;
            if (is_stale_nan(vals[j])) continue;
            if (is_stale_nan(vPrev)) {
                vPrev = vals[j];
                continue;
            }
            if (vals[j] > vPrev) vPrev = vals[j];
        }
        ts[o] = tsPrev;  // o <= j: never overwrites rows that are still to be read
        vals[o] = vPrev;
        o++;
    };
    for (size_t i = 1; i < n; i++) {
        int64_t t = ts[i];
        if (t <= tsNext) continue;
        pick(i - 1);
        tsNext += interval;
        if (tsNext < t) {
            tsNext = t + interval - 1;
            tsNext -= tsNext % interval;
        }
    }
    pick(n - 1);
    return o;
}

// blocks b = 0..nblocks-1 hold rows [offsets[b], offsets[b+1]) of ts/vals (each block sorted by timestamp, already trimmed
// to the query time range); out_* must hold offsets[nblocks] rows.  -> number of rows written
extern "C" size_t vmo_merge_sort_blocks(const int64_t* ts, const double* vals, const uint64_t* offsets, size_t nblocks,
                                        int64_t dedup_interval, int64_t* out_ts, double* out_vals) {
    std::vector<SortBlock> blocks(nblocks);
    Heap h;
    for (size_t b = 0; b < nblocks; b++) {
        blocks[b] = SortBlock{ts + offsets[b], vals + offsets[b], (size_t)(offsets[b + 1] - offsets[b]), 0};
        if (blocks[b].n) h.sbs.push_back(&blocks[b]);  // empty blocks are skipped :568-575
    }
    if (h.sbs.empty()) return 0;
    h.init();
    size_t o = 0;
    for (;;) {
        SortBlock* top = h.sbs[0];
        if (h.sbs.size() == 1) {
            size_t m = top->n - top->next;
            memcpy(out_ts + o, top->ts + top->next, m * 8);
            memcpy(out_vals + o, top->vals + top->next, m * 8);
            o += m;
            break;
        }
        SortBlock* nx = h.next_block();
        int64_t tsNext = nx->ts[nx->next];
        size_t idx = top->next;
        size_t n = equal_samples_prefix(top, nx);
        if (n > 0 && dedup_interval > 0) {
            top->next = idx + n;
        } else {
            top->next = idx + binary_search_timestamps(top->ts + idx, top->n - idx, tsNext);
            size_t m = top->next - idx;
            memcpy(out_ts + o, top->ts + idx, m * 8);
            memcpy(out_vals + o, top->vals + idx, m * 8);
            o += m;
        }
        if (top->next < top->n) h.fix(0);
        else h.pop();
    }
    return vmo_deduplicate_samples(out_ts, out_vals, o, dedup_interval);
}
