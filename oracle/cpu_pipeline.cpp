// ORACLE (test infrastructure, not product code): the reference's per-series query loop on CPU threads, used as the
// timed CPU baseline (bench.py cpu_baseline / --impl reference) and as the whole-path checker.
//
// Mirrors Results.RunParallel (app/vmselect/netstorage/netstorage.go:221: one series per task, workers pull tasks) with
// the per-series closure of evalRollupNoIncrementalAggregate (app/vmselect/promql/eval.go:1855): Unpack ->
// Block.UnmarshalData -> AppendRowsWithTimeRangeFilter -> dropStaleNaNs -> preFunc (removeCounterResets) ->
// rollupConfig.Do.  zstd goes through the reference's own libzstd (oracle/_ref) when present, else the oracle decoder.
#include <atomic>
#include <cstring>
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

// returns 0 or the first error; out: [nblocks x P]; one block == one series (all BASELINE configs)
int vmo_cpu_eval_rollup(const vmo_block_desc* descs, size_t nblocks, const uint8_t* payload, int64_t tr_min, int64_t tr_max,
                        const vmo_rollup_cfg* cfg, int remove_counter_resets, int drop_stale_nans, double* out,
                        uint64_t* samples_scanned, int nthreads, int use_ref_zstd) {
    const int64_t P = vmo_rollup_points(cfg->start, cfg->end, cfg->step);
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    std::atomic<uint64_t> scanned{0};
    const int64_t max_stale = cfg->lookback_delta != 0 ? cfg->lookback_delta + cfg->window : 0;
    auto worker = [&]() {
        std::vector<int64_t> ts(16384), iv(16384);
        std::vector<double> fv(16384);
        std::vector<uint8_t> zt, zv;
        uint64_t local_scanned = 0;
        for (;;) {
            size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
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
        scanned += local_scanned;
    };
    if (nthreads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (samples_scanned) *samples_scanned = scanned.load();
    return err.load();
}

}  // extern "C"
