// ORACLE (test infrastructure, not product code): the reference's per-series query loop on CPU threads, used as the
// timed CPU baseline (bench.py cpu_baseline / --impl reference) and as the whole-path checker.
//
// Mirrors Results.RunParallel (app/vmselect/netstorage/netstorage.go:221: one series per task, workers pull tasks) with
// the per-series closure of evalRollupNoIncrementalAggregate (app/vmselect/promql/eval.go:1855): Unpack ->
// Block.UnmarshalData -> AppendRowsWithTimeRangeFilter -> dropStaleNaNs -> preFunc (removeCounterResets) ->
// rollupConfig.Do.  zstd goes through the reference's own libzstd (oracle/_ref) when present, else the oracle decoder.
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "vm_oracle.h"

extern "C" {

typedef struct {  // same 64-byte layout as the product's vmb_block_desc (include/vmb200.h) so tests can share buffers
    int64_t first_value, min_ts, max_ts;
    uint64_t ts_off, val_off;
    uint32_t ts_size, val_size, rows, series_idx;
    int16_t scale;
    uint8_t ts_mt, val_mt, precision_bits, _pad[3];
} vmo_block_desc;

// ---- persistent worker pool (the Go runtime keeps its worker goroutines / Ps alive between queries: netstorage.go:221
// RunParallel hands series to already running workers).  bench.py creates the pool once, OUTSIDE every timed region.
struct VmoPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<void(int)> job;  // job(thread index)
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int)> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                j = job;
            }
            j(idx);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    void run(const std::function<void(int)>& j) {
        std::unique_lock<std::mutex> lk(mu);
        job = j;
        pending = (int)th.size();
        gen++;
        cv_job.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

void* vmo_pool_create(int nthreads) {
    if (nthreads < 1) nthreads = 1;
    VmoPool* p = new VmoPool();
    for (int i = 0; i < nthreads; i++) p->th.emplace_back([p, i] { p->loop(i); });
    return p;
}
void vmo_pool_destroy(void* h) {
    VmoPool* p = (VmoPool*)h;
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto& t : p->th) t.join();
    delete p;
}
int vmo_pool_threads(void* h) { return h ? (int)((VmoPool*)h)->th.size() : 0; }

// returns 0 or the first error; out: [nblocks x P]; one block == one series (all BASELINE configs)
static int cpu_eval_rollup_impl(VmoPool* pool, const vmo_block_desc* descs, size_t nblocks, const uint8_t* payload, int64_t tr_min,
                                int64_t tr_max, const vmo_rollup_cfg* cfg, int remove_counter_resets, int drop_stale_nans, double* out,
                                uint64_t* samples_scanned, int nthreads, int use_ref_zstd) {
    const int64_t P = vmo_rollup_points(cfg->start, cfg->end, cfg->step);
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    std::atomic<uint64_t> scanned{0};
    const int64_t max_stale = cfg->lookback_delta != 0 ? cfg->lookback_delta + cfg->window : 0;
    const size_t grab = 8;  // series per task pull (RunParallel hands out one series at a time; a few per pull keeps the
                            // shared counter off the profile at 128 threads)
    auto worker = [&](int) {
        static thread_local std::vector<int64_t> ts(16384), iv(16384);
        static thread_local std::vector<double> fv(16384);
        static thread_local std::vector<uint8_t> zt, zv;
        uint64_t local_scanned = 0;
        for (;;) {
            const size_t b_first = next.fetch_add(grab);
            if (b_first >= nblocks) break;
            const size_t b_last = b_first + grab < nblocks ? b_first + grab : nblocks;
            for (size_t b = b_first; b < b_last; b++) {
            const vmo_block_desc& d = descs[b];
            vmo_block_header bh;
            memset(&bh, 0, sizeof(bh));
            bh.min_ts = d.min_ts;
            bh.max_ts = d.max_ts;
            bh.first_value = d.first_value;
            bh.ts_size = d.ts_size;
            bh.val_size = d.val_size;
            bh.rows = d.rows;
            bh.scale = d.scale;
            bh.ts_mt = d.ts_mt;
            bh.val_mt = d.val_mt;
            bh.precision_bits = d.precision_bits;
            const uint8_t* tsd = payload + d.ts_off;
            const uint8_t* vald = payload + d.val_off;
            // the reference decompresses with libzstd; do the same here (then hand plain varints to the restatement)
            if (use_ref_zstd && vmo_zstd_ref_available()) {
                for (int which = 0; which < 2; which++) {
                    uint8_t& mt = which ? bh.val_mt : bh.ts_mt;
                    if (mt != 1 && mt != 4) continue;
                    const uint8_t* src = which ? vald : tsd;
                    uint32_t len = which ? d.val_size : d.ts_size;
                    std::vector<uint8_t>& z = which ? zv : zt;
                    int64_t cs = vmo_zstd_content_size(src, len);
                    if (cs < 0) { err = VMO_ERR_ZSTD; return; }
                    z.resize((size_t)cs + 16);
                    int64_t r = vmo_zstd_ref_decompress(z.data(), (size_t)cs, src, len);
                    if (r < 0) { err = VMO_ERR_ZSTD; return; }
                    if (which) { vald = z.data(); bh.val_size = (uint32_t)r; }
                    else { tsd = z.data(); bh.ts_size = (uint32_t)r; }
                    mt = mt == 1 ? 5 : 6;
                }
                // NeedsValidation() is false for the zstd types (encoding.go:46): keep it false after the swap
                if ((d.ts_mt == 1 || d.ts_mt == 4) && bh.precision_bits == 64) {
                    bh.min_ts = INT64_MIN;  // disables the bounds check that the original type would have skipped
                    bh.max_ts = INT64_MAX;
                }
            }
            int64_t first_ts_save = d.min_ts;
            // UnmarshalTimestamps needs the real first timestamp even when the bounds check is disabled
            int64_t n;
            if (bh.min_ts == INT64_MIN) {
                // decode timestamps / values separately (same calls Block.UnmarshalData makes)
                int rc = vmo_unmarshal_int64_array(ts.data(), tsd, bh.ts_size, bh.ts_mt, first_ts_save, d.rows);
                if (!rc) rc = vmo_unmarshal_int64_array(iv.data(), vald, bh.val_size, bh.val_mt, d.first_value, d.rows);
                if (rc) { err = rc; return; }
                size_t i = 0, j = d.rows;
                while (i < j && ts[i] < tr_min) i++;
                while (j > i && ts[j - 1] > tr_max) j--;
                n = (int64_t)(j - i);
                if (i) memmove(ts.data(), ts.data() + i, (size_t)n * 8);
                vmo_decimal_to_float(fv.data(), iv.data() + i, (size_t)n, d.scale);
            } else {
                n = vmo_block_unmarshal(ts.data(), fv.data(), iv.data(), &bh, tsd, vald, tr_min, tr_max);
                if (n < 0) { err = (int)n; return; }
            }
            size_t m = (size_t)n;
            if (drop_stale_nans) m = vmo_drop_stale_nans(fv.data(), ts.data(), m);
            if (remove_counter_resets) vmo_remove_counter_resets(fv.data(), ts.data(), m, max_stale);
            local_scanned += vmo_rollup_do(cfg, out + b * (size_t)P, fv.data(), ts.data(), m);
            }
        }
        scanned += local_scanned;
    };
    if (pool) pool->run(worker);
    else if (nthreads <= 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    if (samples_scanned) *samples_scanned = scanned.load();
    return err.load();
}

// returns 0 or the first error; out: [nblocks x P]; one block == one series (all BASELINE configs).  Spawns its threads.
int vmo_cpu_eval_rollup(const vmo_block_desc* descs, size_t nblocks, const uint8_t* payload, int64_t tr_min, int64_t tr_max,
                        const vmo_rollup_cfg* cfg, int remove_counter_resets, int drop_stale_nans, double* out,
                        uint64_t* samples_scanned, int nthreads, int use_ref_zstd) {
    return cpu_eval_rollup_impl(nullptr, descs, nblocks, payload, tr_min, tr_max, cfg, remove_counter_resets, drop_stale_nans, out,
                                samples_scanned, nthreads, use_ref_zstd);
}
// the same on an already running pool (bench.py: the timed region holds no thread creation)
int vmo_pool_eval_rollup(void* pool, const vmo_block_desc* descs, size_t nblocks, const uint8_t* payload, int64_t tr_min,
                         int64_t tr_max, const vmo_rollup_cfg* cfg, int remove_counter_resets, int drop_stale_nans, double* out,
                         uint64_t* samples_scanned, int use_ref_zstd) {
    if (!pool) return VMO_ERR_BUG;
    return cpu_eval_rollup_impl((VmoPool*)pool, descs, nblocks, payload, tr_min, tr_max, cfg, remove_counter_resets, drop_stale_nans,
                                out, samples_scanned, 0, use_ref_zstd);
}

// Block.MarshalData (lib/storage/block.go:192) for ncols equal-length int64 columns with the REFERENCE encoder
// (vmo_marshal_int64_array = marshalInt64Array encoding.go:119, zstd through the reference's libzstd at getCompressLevel):
// what a vmstorage part holds.  Payloads back to back in dst; offs[ncols + 1].  Used to build test / bench inputs.
int64_t vmo_pool_marshal_columns(void* pool, uint8_t* dst, size_t cap, uint64_t* offs, uint8_t* mts, int64_t* firsts,
                                 const int64_t* vals, size_t ncols, size_t rows, uint8_t precision_bits) {
    if (!pool || !dst || !offs || !mts || !firsts || !vals || rows == 0) return VMO_ERR_BUG;
    std::vector<std::vector<uint8_t>> outs(ncols);
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    auto worker = [&](int) {
        std::vector<uint8_t> tmp(rows * 10 + 1024);
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= ncols) break;
            int mt = 0;
            int64_t first = 0;
            const int64_t r = vmo_marshal_int64_array(tmp.data(), tmp.size(), vals + c * rows, rows, precision_bits, &mt, &first);
            if (r < 0) { err = (int)r; continue; }
            outs[c].assign(tmp.begin(), tmp.begin() + r);
            mts[c] = (uint8_t)mt;
            firsts[c] = first;
        }
    };
    ((VmoPool*)pool)->run(worker);
    if (err.load()) return err.load();
    uint64_t o = 0;
    for (size_t c = 0; c < ncols; c++) {
        offs[c] = o;
        if (o + outs[c].size() > cap) return VMO_ERR_CAP;
        if (!outs[c].empty()) memcpy(dst + o, outs[c].data(), outs[c].size());
        o += outs[c].size();
    }
    offs[ncols] = o;
    return (int64_t)o;
}


// ---- synthetic input of bench.py / the full-size tests, generated and marshaled series by series on the pool so
// that the [series x rows] matrix is never materialised (config 3: 1 M series x 8192 rows would be 65 GB of int64).
// Every series is marshaled with the REFERENCE encoder (vmo_marshal_int64_array: type detection, nearest-delta(2), libzstd
// 1.5.7 at getCompressLevel) -- the bytes a vmstorage part would hold.
//   kind: 0 counter (node_cpu_seconds_total-like: increments U[0,1500] at scale -2, reset to 0 with p = 1e-4 per sample),
//         1 gauge round(N(5000, 300)), 2 mixed by series index mod 10 (0-3 counter, 4-6 gauge, 7-8 const, 9 delta-const)
//   ts_kind: 0 regular t0 + dt * i (one shared MarshalTypeDeltaConst payload at offset 0), 1 per-series jitter of +-50 ms
// descs: [nblocks] (series_idx = block index).  payload: caller buffer of `cap` bytes.  stats[4] = {series, series with a value
// drop, rows, rows from the 128-row group of the first value drop on}.  Returns the payload length or < 0.
namespace {
struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// values of synthetic series `b` (see vmo_pool_gen_blocks); *rng is left positioned after the values (the jitter of the
// timestamps continues the same stream)
static void gen_series_values(int kind, size_t b, size_t rows, uint64_t seed, int64_t* v, SplitMix* rng) {
    // per-series stream: the state is a hash of (seed, series) -- consecutive series must not be shifted copies of one stream
    SplitMix h{seed ^ (0xd1b54a32d192ed03ull * (uint64_t)(b + 1))};
    h.next();
    SplitMix r{h.next() ^ (h.next() << 1)};
    int k = kind == 2 ? (int)(b % 10) : (kind == 0 ? 0 : 4);
    if (k < 4) {  // counter
        int64_t x = (int64_t)r.below(1000000000ull);
        for (size_t i = 0; i < rows; i++) {
            x += (int64_t)r.below(1501);
            if (i > 0 && r.unit() < 1e-4) x = 0;
            v[i] = x;
        }
    } else if (k < 7) {  // gauge: Box-Muller
        for (size_t i = 0; i < rows; i += 2) {
            const double u1 = 1.0 - r.unit(), u2 = r.unit();
            const double m = std::sqrt(-2.0 * std::log(u1));
            v[i] = (int64_t)std::llround(5000.0 + 300.0 * m * std::cos(6.283185307179586 * u2));
            if (i + 1 < rows) v[i + 1] = (int64_t)std::llround(5000.0 + 300.0 * m * std::sin(6.283185307179586 * u2));
        }
    } else if (k < 9) {
        const int64_t c = (int64_t)r.below(1000000);
        for (size_t i = 0; i < rows; i++) v[i] = c;
    } else {
        const int64_t c = (int64_t)r.below(1000000), d = 1 + (int64_t)r.below(99);
        for (size_t i = 0; i < rows; i++) v[i] = c + d * (int64_t)i;
    }
    *rng = r;
}
}  // namespace

// the raw int64 mantissas of series [b0, b0 + nb) into out[nb x rows] (same values as vmo_pool_gen_blocks marshals): lets a
// second encoder (the product's own, through its C ABI) marshal the very same data
int vmo_pool_gen_values(void* pool, int kind, size_t b0, size_t nb, size_t rows, uint64_t seed, int64_t* out) {
    if (!pool || !out || rows == 0) return VMO_ERR_BUG;
    std::atomic<size_t> next{0};
    auto worker = [&](int) {
        SplitMix r{0};
        for (;;) {
            const size_t i = next.fetch_add(16);
            if (i >= nb) break;
            for (size_t k = i; k < i + 16 && k < nb; k++) gen_series_values(kind, b0 + k, rows, seed, out + k * rows, &r);
        }
    };
    ((VmoPool*)pool)->run(worker);
    return 0;
}

int64_t vmo_pool_gen_blocks(void* pool, int kind, int ts_kind, size_t nblocks, size_t rows, uint64_t seed, int64_t t0, int64_t dt,
                            int16_t scale, vmo_block_desc* descs, uint8_t* payload, size_t cap, uint64_t* stats) {
    if (!pool || !descs || !payload || rows == 0 || rows > 16384) return VMO_ERR_BUG;
    VmoPool* P = (VmoPool*)pool;
    const int nt = (int)P->th.size();
    std::vector<std::vector<uint8_t>> bufs(nt);
    struct Loc { uint32_t thread; uint64_t off; };
    std::vector<Loc> vloc(nblocks), tloc(ts_kind ? nblocks : 0);
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    std::atomic<uint64_t> st_drop{0}, st_rows_from_drop{0};
    // shared timestamps payload of the regular case
    std::vector<int64_t> ts0(rows);
    for (size_t i = 0; i < rows; i++) ts0[i] = t0 + dt * (int64_t)i;
    uint8_t tshared[16];
    int tmt0 = 0;
    int64_t tfirst0 = 0;
    const int64_t tlen0 = vmo_marshal_int64_array(tshared, sizeof(tshared), ts0.data(), rows, 64, &tmt0, &tfirst0);
    if (tlen0 < 0) return tlen0;
    auto worker = [&](int ti) {
        std::vector<int64_t> v(rows), tj(rows);
        std::vector<uint8_t> tmp(rows * 10 + 1024);
        std::vector<uint8_t>& out = bufs[ti];
        uint64_t ndrop = 0, nrows_from = 0;
        for (;;) {
            const size_t b0 = next.fetch_add(16);
            if (b0 >= nblocks) break;
            for (size_t b = b0; b < b0 + 16 && b < nblocks; b++) {
                SplitMix r{0};
                gen_series_values(kind, b, rows, seed, v.data(), &r);
                size_t fd = rows;
                for (size_t i = 1; i < rows; i++)
                    if (v[i] < v[i - 1]) { fd = i; break; }
                if (fd < rows) {
                    ndrop++;
                    nrows_from += rows - (fd & ~(size_t)127);
                }
                vmo_block_desc& d = descs[b];
                memset(&d, 0, sizeof(d));
                int mt = 0;
                int64_t first = 0;
                int64_t len = vmo_marshal_int64_array(tmp.data(), tmp.size(), v.data(), rows, 64, &mt, &first);
                if (len < 0) { err = (int)len; return; }
                vloc[b] = {(uint32_t)ti, (uint64_t)out.size()};
                out.insert(out.end(), tmp.begin(), tmp.begin() + len);
                d.first_value = first;
                d.val_size = (uint32_t)len;
                d.val_mt = (uint8_t)mt;
                d.rows = (uint32_t)rows;
                d.series_idx = (uint32_t)b;
                d.scale = scale;
                d.precision_bits = 64;
                if (ts_kind) {
                    for (size_t i = 0; i < rows; i++) tj[i] = ts0[i] + (int64_t)r.below(101) - 50;
                    len = vmo_marshal_int64_array(tmp.data(), tmp.size(), tj.data(), rows, 64, &mt, &first);
                    if (len < 0) { err = (int)len; return; }
                    tloc[b] = {(uint32_t)ti, (uint64_t)out.size()};
                    out.insert(out.end(), tmp.begin(), tmp.begin() + len);
                    d.ts_size = (uint32_t)len;
                    d.ts_mt = (uint8_t)mt;
                    d.min_ts = first;
                    d.max_ts = tj[rows - 1];
                } else {
                    d.ts_off = 0;
                    d.ts_size = (uint32_t)tlen0;
                    d.ts_mt = (uint8_t)tmt0;
                    d.min_ts = tfirst0;
                    d.max_ts = ts0[rows - 1];
                }
            }
        }
        st_drop += ndrop;
        st_rows_from_drop += nrows_from;
    };
    P->run(worker);
    if (err.load()) return err.load();
    // final layout in block order: [shared timestamps payload] then per block [timestamps (jitter only)] [values]
    uint64_t o = 0;
    if (!ts_kind) {
        if ((size_t)tlen0 > cap) return VMO_ERR_CAP;
        memcpy(payload, tshared, (size_t)tlen0);
        o = (uint64_t)tlen0;
    }
    for (size_t b = 0; b < nblocks; b++) {
        if (ts_kind) {
            descs[b].ts_off = o;
            o += descs[b].ts_size;
        }
        descs[b].val_off = o;
        o += descs[b].val_size;
    }
    if (o > cap) return VMO_ERR_CAP;
    next = 0;
    auto copier = [&](int) {
        for (;;) {
            const size_t b0 = next.fetch_add(256);
            if (b0 >= nblocks) break;
            for (size_t b = b0; b < b0 + 256 && b < nblocks; b++) {
                if (ts_kind) memcpy(payload + descs[b].ts_off, bufs[tloc[b].thread].data() + tloc[b].off, descs[b].ts_size);
                memcpy(payload + descs[b].val_off, bufs[vloc[b].thread].data() + vloc[b].off, descs[b].val_size);
            }
        }
    };
    P->run(copier);
    if (stats) {
        stats[0] = nblocks;
        stats[1] = st_drop.load();
        stats[2] = (uint64_t)nblocks * rows;
        stats[3] = st_rows_from_drop.load();
    }
    return (int64_t)o;
}

}  // extern "C"
