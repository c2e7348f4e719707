/*
 * vmb200.h -- C ABI of libvmb200: a B200-native (sm_100a) implementation of VictoriaMetrics' data-parallel
 * hot path: the per-series block codec (lib/encoding + lib/decimal) and the range-vector rollup executor
 * (app/vmselect/promql), as scoped by SURVEY.md section 8.
 *
 * The reference has no FFI for this path (it is plain Go calls); each entry point below names the Go function
 * (file:line under the reference checkout) it replaces.  The Go-side cgo binding is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 (VMB_OK) or a negative VMB_ERR_* code; vmb_last_error() gives thread-local text.
 *   - plain pointers and sizes only.  Host buffers are caller-owned and never retained after return.
 *   - device memory is library-owned behind opaque handles (vmb_blocks, vmb_series), or caller-owned raw device
 *     pointers where an argument is documented as "device pointer" (e.g. a torch tensor's data_ptr()).
 *   - there is NO CPU fallback: every compute entry point launches CUDA kernels and fails with
 *     VMB_ERR_CUDA if no sm_100-class device is usable.
 *   - all kernels of one vmb_ctx are issued on the ctx's stream (vmb_ctx_set_stream), so the caller can
 *     bracket them with its own CUDA events.
 */
#ifndef VMB200_H
#define VMB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMB_OK 0
/* per-column decode errors: same numbering as the reference error sites they stand for */
#define VMB_ERR_SHORT_SRC (-1)        /* lib/encoding/int.go:183,199,211 */
#define VMB_ERR_VARINT_TOO_BIG (-2)   /* int.go:272 */
#define VMB_ERR_VARINT_TOO_LONG (-3)  /* int.go:277 */
#define VMB_ERR_TAIL (-4)             /* nearest_delta.go:65, encoding.go:238 */
#define VMB_ERR_MARSHAL_TYPE (-5)     /* encoding.go:248 */
#define VMB_ERR_ZSTD (-6)             /* encoding.go:181,193 */
#define VMB_ERR_CONST_TAIL (-7)       /* encoding.go:217 */
#define VMB_ERR_DELTA_CONST (-8)      /* encoding.go:235 */
#define VMB_ERR_TS_BOUNDS (-9)        /* lib/storage/block.go:298 checkTimestampsBounds */
#define VMB_ERR_ROWS (-10)            /* block.go:263 RowsCount must be > 0; block_header.go:233 <= 16384 */
#define VMB_ERR_BLOCK_ORDER (-11)     /* internal consistency check of the series assembly (a hole between time-disjoint blocks) */
/* API-level errors */
#define VMB_ERR_INVALID_ARG (-50)     /* the Go code would logger.Panicf("BUG: ...") */
#define VMB_ERR_CUDA (-51)
#define VMB_ERR_NOMEM (-52)
#define VMB_ERR_BLOCK_FAILED (-53)    /* at least one block failed to decode; see the per-block status array */
#define VMB_ERR_CAP (-54)
#define VMB_ERR_COMM (-55)            /* NCCL could not be loaded, or a collective failed (vmb_last_error has the NCCL text) */

typedef struct vmb_ctx vmb_ctx;
typedef struct vmb_blocks vmb_blocks;  /* compressed blocks resident in HBM (descriptors + payload arena) */
typedef struct vmb_series vmb_series;  /* decoded columns resident in HBM: int64 timestamps + f64 values per series */

/* ---- context ------------------------------------------------------------------------------------------- */
int vmb_ctx_create(int device, vmb_ctx** out);
void vmb_ctx_destroy(vmb_ctx* ctx);
/* stream = a cudaStream_t (0 = legacy default stream). */
int vmb_ctx_set_stream(vmb_ctx* ctx, void* stream);
/* storage.SetDedupInterval lib/storage/dedup.go:15 (-dedup.minScrapeInterval), in ms; 0 (default) = off.  Applied by every
 * block-decoding entry point after the series are assembled (netstorage.go:611 DeduplicateSamples). */
int vmb_ctx_set_dedup_interval(vmb_ctx* ctx, int64_t interval_ms);
/* The one-call device paths (vmb_eval_rollup_device, vmb_eval_rollup_aggr_device) run series that qualify -- one block,
 * MarshalTypeDeltaConst timestamps at precisionBits 64 -- through one fused kernel per batch: decode into shared memory,
 * removeCounterResets and rollupConfig.Do without materialising the decoded columns (the per-series shape of eval.go:1855-1866).
 * Everything else takes the kernel-per-stage pipeline.  enable = 0 forces that pipeline for every series (default: 1; the
 * environment variable VMB_NO_FUSED sets the default to 0).  Results are bit-identical either way. */
int vmb_ctx_set_fused(vmb_ctx* ctx, int enable);
int vmb_ctx_synchronize(vmb_ctx* ctx);
const char* vmb_last_error(void);
int vmb_version(void);
/* number of kernel launches issued by this ctx since creation (for bench.py's gpu_launches) */
uint64_t vmb_ctx_launch_count(const vmb_ctx* ctx);

/* ---- block descriptor  ==  lib/storage/block_header.go:19-82 blockHeader (the 81-byte wire form is parsed by
 * vmb_block_desc_from_header) --------------------------------------------------------------------------- */
typedef struct {
    int64_t first_value;     /* FirstValue */
    int64_t min_ts;          /* MinTimestamp (== firstTimestamp passed to UnmarshalTimestamps) */
    int64_t max_ts;          /* MaxTimestamp */
    uint64_t ts_off;         /* byte offset of the timestamps payload inside the payload arena */
    uint64_t val_off;        /* byte offset of the values payload inside the payload arena */
    uint32_t ts_size;        /* TimestampsBlockSize */
    uint32_t val_size;       /* ValuesBlockSize */
    uint32_t rows;           /* RowsCount, 1..16384 */
    uint32_t series_idx;     /* dense series index; the blocks of one series are consecutive, in any order (may overlap) */
    int16_t scale;           /* Scale: value = mantissa * 10^scale */
    uint8_t ts_mt;           /* TimestampsMarshalType 1..6 (lib/encoding/encoding.go:20-43) */
    uint8_t val_mt;          /* ValuesMarshalType */
    uint8_t precision_bits;  /* PrecisionBits 1..64 */
    uint8_t _pad[3];
} vmb_block_desc; /* 64 bytes */

/* blockHeader.Unmarshal block_header.go:122 (81 bytes, big endian); ts_off/val_off are taken from the header
 * (file offsets) -- the caller rebases them onto its payload arena.  series_idx is left 0. */
int vmb_block_desc_from_header(vmb_block_desc* out, const uint8_t header[81], uint8_t tsid_out[24]);

/* blockHeader.Marshal block_header.go:104 (inverse of the above; tsid may be NULL = zeros) */
int vmb_block_header_marshal(uint8_t header[81], const vmb_block_desc* d, const uint8_t tsid[24]);
/* unmarshalBlockHeaders block_header.go:261 (part_search.go:247): an uncompressed index block = `count` 81-byte headers
 * sorted by TSID.  out[count]; tsids ([count * 24], may be NULL) receives the TSIDs.  Errors: wrong length
 * (VMB_ERR_SHORT_SRC / VMB_ERR_ROWS), a header that fails blockHeader.validate (:230), unsorted TSIDs (VMB_ERR_INVALID_ARG). */
int vmb_index_block_unmarshal(vmb_block_desc* out, uint8_t* tsids, size_t count, const uint8_t* data, size_t len);
/* metaindexRow metaindex_row.go:12 (56-byte big-endian wire form :61) */
typedef struct {
    uint8_t tsid[24];
    int64_t min_ts, max_ts;
    uint64_t index_block_offset;
    uint32_t block_headers_count, index_block_size;
} vmb_metaindex_row; /* 56 bytes */
/* unmarshalMetaindexRows metaindex_row.go:129 on the decompressed metaindex.bin (vmb_zstd_decompress_batch below);
 * *n receives the number of rows (also when cap is too small: VMB_ERR_CAP).  Same checks as the Go code: at least one row,
 * BlockHeadersCount > 0, IndexBlockSize <= 2*maxBlockSize, rows sorted by TSID. */
int vmb_metaindex_rows_unmarshal(vmb_metaindex_row* out, size_t cap, size_t* n, const uint8_t* data, size_t len);
int vmb_metaindex_row_marshal(uint8_t out[56], const vmb_metaindex_row* row); /* metaindexRow.Marshal :61 */
/* encoding.DecompressZSTD (lib/encoding/compress.go:27) for n frames at once on the GPU -- index blocks (part_search.go:238)
 * and metaindex.bin (metaindex_row.go:134) go through the same kernels as the block payloads.  frames + offs[n+1] = the
 * compressed frames back to back; every frame must declare its content size (libzstd/gozstd always do; sanity cap 128 MiB).
 * Frame i is written to dst + dst_offs[i] (16-byte aligned), dst_lens[i] bytes; vmb_zstd_decompress_bound gives the dst size.
 * statuses ([n], may be NULL): 0 or VMB_ERR_ZSTD per frame; returns VMB_ERR_ZSTD if any frame failed. */
int vmb_zstd_decompress_bound(const uint8_t* frames, const uint64_t* offs, size_t n, uint64_t* out_bytes);
int vmb_zstd_decompress_batch(vmb_ctx* ctx, const uint8_t* frames, const uint64_t* offs, size_t n, uint8_t* dst,
                              size_t dst_cap, uint64_t* dst_offs, uint32_t* dst_lens, int32_t* statuses);

/* ---- per-call drop-ins (single column; host buffers; run on the GPU) ------------------------------------- */
/* encoding.UnmarshalValues / UnmarshalTimestamps  encoding.go:111 / :90 (unmarshalInt64Array :173) */
int vmb_unmarshal_int64(vmb_ctx* ctx, int64_t* dst, size_t items_count, const uint8_t* src, size_t src_len, int mt,
                        int64_t first_value);
/* decimal.AppendDecimalToFloat  lib/decimal/decimal.go:100 */
int vmb_decimal_to_float(vmb_ctx* ctx, double* dst, const int64_t* va, size_t n, int16_t e);
/* encoding.MarshalValues / MarshalTimestamps  encoding.go:103 / :82 (marshalInt64Array :119).
 * Host-side encoder (the write path stays on the host this round); zstd frames are produced by
 * the library's own Huffman-literals compressor: valid zstd that libzstd/klauspost decode, not byte-identical to
 * libzstd's output (compressed bytes are unpinned by the reference's tests). */
int vmb_marshal_int64(uint8_t* dst, size_t cap, size_t* out_len, int* out_mt, int64_t* out_first, const int64_t* vals,
                      size_t n, uint8_t precision_bits);
/* the library's zstd writer on its own (exposed for tests) */
int vmb_zstd_compress(uint8_t* dst, size_t cap, size_t* out_len, const uint8_t* src, size_t n);
/* Block.MarshalData (block.go:192) for ncols equal-length int64 columns on `nthreads` host threads: payloads are written
 * back to back into dst, offs[ncols+1] receives their offsets, mts/firsts the MarshalType and first value of each. */
int vmb_marshal_columns(uint8_t* dst, size_t cap, uint64_t* offs, uint8_t* mts, int64_t* firsts, const int64_t* vals,
                        size_t ncols, size_t rows, uint8_t precision_bits, int nthreads);
/* The same with the int64 work on the GPU (csrc/encode.cu): isConst / isDeltaConst / isGauge (encoding.go:289-369) as one pass of warp
 * reductions per column, nearest-delta / delta2 (lossless, and the lossy precisionBits < 64 state machine of nearest_delta.go:83),
 * zig-zag varint packing by warp scans; the zstd stage of streams >= 128 bytes and the 0.9 rule (encoding.go:152-167) follow on
 * `nthreads` host threads.  Byte-identical to vmb_marshal_columns.  vals: HOST [ncols x rows]. */
int vmb_marshal_columns_gpu(vmb_ctx* ctx, uint8_t* dst, size_t cap, uint64_t* offs, uint8_t* mts, int64_t* firsts,
                            const int64_t* vals, size_t ncols, size_t rows, uint8_t precision_bits, int nthreads);
/* decimal.AppendFloatToDecimal decimal.go:173 (host-side, write path) */
int vmb_float_to_decimal(int64_t* dst, int16_t* out_scale, const double* src, size_t n);
/* ... for ncols equal-length columns on the GPU (one warp per column: FromFloat per value, min-exponent / overflow reductions,
 * rescale); dst HOST [ncols x rows], scales HOST [ncols], src HOST [ncols x rows].  Same results as the call above per column. */
int vmb_float_to_decimal_columns(vmb_ctx* ctx, int64_t* dst, int16_t* scales, const double* src, size_t ncols, size_t rows);
/* decimal.CalibrateScale decimal.go:13 (host-side; block merge path lib/storage/merge.go): a and b are rescaled in place to
 * the common exponent returned in *out_e. */
int vmb_calibrate_scale(int64_t* a, size_t na, int16_t ae, int64_t* b, size_t nb, int16_t be, int16_t* out_e);

/* ---- batched block decode  ==  Block.UnmarshalData (block.go:250) + AppendRowsWithTimeRangeFilter (:324)
 * for every block of a query at once (replaces netstorage.go:425 packedTimeseries.Unpack fan-out) ------------- */
/* copies descriptors + payload host->device (the one H2D of the path). payload_len bytes are copied. */
int vmb_blocks_upload(vmb_ctx* ctx, const vmb_block_desc* descs, size_t nblocks, const uint8_t* payload,
                      size_t payload_len, vmb_blocks** out);
/* The same from a part on disk: `headers` = nblocks marshaled blockHeaders (81 bytes each) -- what vmselect keeps per block in its
 * tmpBlocksFile (tmp_blocks_file.go:110 WriteBlockRefData, parsed back by BlockRef.Init lib/storage/search.go:38) --, in series
 * order (consecutive headers with the same TSID form one series; netstorage.go groups BlockRefs that way); timestamps_bin /
 * values_bin = the part's two data files (mmap'ed or read), addressed by TimestampsBlockOffset / ValuesBlockOffset like
 * BlockRef.MustReadBlock search.go:73.  Only the referenced byte ranges are copied.  Errors: a header that fails
 * blockHeader.validate, offsets outside the files (VMB_ERR_SHORT_SRC). */
int vmb_blocks_upload_part(vmb_ctx* ctx, const uint8_t* headers, size_t nblocks, const uint8_t* timestamps_bin, size_t ts_len,
                           const uint8_t* values_bin, size_t val_len, vmb_blocks** out);
void vmb_blocks_free(vmb_blocks* b);
size_t vmb_blocks_count(const vmb_blocks* b);
uint64_t vmb_blocks_rows(const vmb_blocks* b);           /* sum of RowsCount */
uint64_t vmb_blocks_compressed_bytes(const vmb_blocks* b); /* sum of ts_size + val_size */

#define VMB_DECODE_VALUES_AS_INT64 1u /* leave values as int64 mantissas (no decimal->float) */
/* decodes all blocks; rows outside [tr_min, tr_max] are trimmed like filterTimestamps (block.go:331).
 * block_status (host, nblocks entries, may be NULL) receives 0 or a VMB_ERR_* per block.
 * Returns VMB_ERR_BLOCK_FAILED if any block failed (the batch is still returned; failed blocks have 0 rows). */
int vmb_decode_blocks(vmb_ctx* ctx, const vmb_blocks* blocks, int64_t tr_min, int64_t tr_max, uint32_t flags,
                      int32_t* block_status, vmb_series** out);

/* builds a device batch from already-decoded host columns (the `values, timestamps` arguments of
 * rollupConfig.Do): series s owns rows [offsets[s], offsets[s+1]) */
int vmb_series_from_host(vmb_ctx* ctx, const int64_t* timestamps, const double* values, const uint64_t* offsets,
                         size_t nseries, vmb_series** out);
/* the feed of evalRollupFuncWithSubquery (eval.go:910): the inner expression's result, a DEVICE matrix [nseries x points] on the grid
 * start, start + step, ..., becomes a batch of series -- removeNanValues (eval.go:1027) drops the NaN points of every row together
 * with their timestamps -- for vmb_rollup with the outer function's config (its preFunc flags apply as usual; dropStaleNaNs does not,
 * there is nothing stale left).  The matrix is not modified and not retained. */
int vmb_series_from_matrix(vmb_ctx* ctx, const double* d_matrix, size_t nseries, size_t points, int64_t start, int64_t step,
                           vmb_series** out);
void vmb_series_free(vmb_series* s);
size_t vmb_series_count(const vmb_series* s);
uint64_t vmb_series_rows(const vmb_series* s); /* allocated rows (before trimming / stale-NaN drop) */
/* per-series current (start,row count) -> host arrays of nseries entries */
int vmb_series_layout(vmb_ctx* ctx, const vmb_series* s, uint64_t* starts, uint32_t* counts);
/* copies the dense decoded columns (vmb_series_rows entries each) device->host; either pointer may be NULL */
int vmb_series_download(vmb_ctx* ctx, const vmb_series* s, int64_t* timestamps, double* values);

/* ---- rollup  ==  rollupConfig.Do (rollup.go:688) for every series at once, preceded by the per-series preamble of
 * eval.go:1855 (dropStaleNaNs eval.go:1985, preFunc = removeCounterResets rollup.go:921) ------------------------- */
enum vmb_rollup_func { /* rollup.go:24-108; names in comments are the MetricsQL names */
    VMB_RF_DEFAULT_ROLLUP = 0, VMB_RF_RATE /* rate, deriv_fast */, VMB_RF_DELTA /* delta, increase */, VMB_RF_AVG,
    VMB_RF_MIN, VMB_RF_MAX, VMB_RF_SUM, VMB_RF_COUNT, VMB_RF_QUANTILE, VMB_RF_FIRST, VMB_RF_LAST, VMB_RF_RANGE,
    VMB_RF_SUM2, VMB_RF_STDDEV, VMB_RF_STDVAR, VMB_RF_IDERIV /* ideriv, irate */, VMB_RF_IDELTA, VMB_RF_DERIV,
    VMB_RF_INCREASE_PURE, VMB_RF_CHANGES, VMB_RF_CHANGES_PROMETHEUS, VMB_RF_RESETS /* resets, decreases_over_time */,
    VMB_RF_INCREASES, VMB_RF_INTEGRATE, VMB_RF_LAG, VMB_RF_LIFETIME, VMB_RF_SCRAPE_INTERVAL, VMB_RF_TMIN, VMB_RF_TMAX,
    VMB_RF_TFIRST, VMB_RF_TLAST /* tlast_over_time, timestamp */, VMB_RF_TLAST_CHANGE, VMB_RF_MODE, VMB_RF_MAD,
    VMB_RF_OUTLIER_IQR, VMB_RF_ZSCORE, VMB_RF_ASCENT, VMB_RF_DESCENT, VMB_RF_DISTINCT, VMB_RF_GEOMEAN,
    VMB_RF_PREDICT_LINEAR, VMB_RF_HOLT_WINTERS, VMB_RF_HOEFFDING_LOWER, VMB_RF_HOEFFDING_UPPER, VMB_RF_DURATION,
    VMB_RF_COUNT_LE, VMB_RF_COUNT_GT, VMB_RF_COUNT_EQ, VMB_RF_COUNT_NE, VMB_RF_SHARE_LE, VMB_RF_SHARE_GT,
    VMB_RF_SHARE_EQ, VMB_RF_SUM_LE, VMB_RF_SUM_GT, VMB_RF_SUM_EQ, VMB_RF_PRESENT, VMB_RF_ABSENT, VMB_RF_STALE_SAMPLES,
    VMB_RF_MEDIAN, VMB_RF_RATE_OVER_SUM, VMB_RF_DELTA_PROMETHEUS /* delta_prometheus, increase_prometheus */,
    VMB_RF_RATE_PROMETHEUS, VMB_RF_OPEN, VMB_RF_CLOSE, VMB_RF_HIGH, VMB_RF_LOW, VMB_RF__COUNT
};

#define VMB_RC_MAY_ADJUST_WINDOW 1u     /* rollupFuncsCanAdjustWindow rollup.go:199 */
#define VMB_RC_IS_DEFAULT_ROLLUP 2u     /* funcName == "default_rollup" rollup.go:408 */
#define VMB_RC_REMOVE_COUNTER_RESETS 4u /* rollupFuncsRemoveCounterResets rollup.go:223: preFunc */
#define VMB_RC_DROP_STALE_NANS 8u       /* eval.go:1985 (not for default_rollup / stale_samples_over_time) */
/* value preFuncs of the multi-output rollups (getRollupConfigs rollup.go:440-476), applied after removeCounterResets,
 * once per batch like it */
#define VMB_RC_PRE_DELTA_VALUES 16u     /* rollup_increase / rollup_delta: deltaValues rollup.go:960 */
#define VMB_RC_PRE_DERIV_VALUES 32u     /* rollup_rate / rollup_deriv: derivValues rollup.go:976 */
#define VMB_RC_PRE_SCRAPE_INTERVAL 64u  /* rollup_scrape_interval: seconds between samples rollup.go:462-474 */
#define VMB_RC_PRE_MASK 112u

typedef struct { /* == rollupConfig rollup.go:574 + what getRollupConfigs (rollup.go:374) derives from the func name */
    int32_t func_id;          /* enum vmb_rollup_func */
    uint32_t flags;           /* VMB_RC_* */
    int64_t start, end, step; /* ms; output grid = start, start+step, ... <= end (eval.go:230 getTimestamps) */
    int64_t window;           /* ms; 0 = not set */
    int64_t lookback_delta;   /* ms; rollupConfig.LookbackDelta */
    int64_t min_staleness_ms; /* -search.minStalenessInterval rollup.go:20 */
    int32_t samples_scanned_per_call; /* rollupFuncsSamplesScannedPerCall rollup.go:238; 0 = len(window) */
    int32_t _pad;
    const double* args;       /* host, P entries or NULL: per-point scalar arg (phi / limit / secs / sf) */
    const double* args2;      /* host, P entries or NULL: second per-point arg (holt_winters tf) */
} vmb_rollup_cfg;

/* number of output points: 1 + (end-start)/step  (eval.go:243) */
int64_t vmb_rollup_points(const vmb_rollup_cfg* cfg);

/* Runs the per-series preamble (in place on the batch: call once per batch) and the rollup.
 * out: [nseries x P] row-major doubles; out_is_device != 0 => out is a device pointer, else host.
 * samples_scanned (host, may be NULL) = sum over series of rollupConfig.Do's second result. */
int vmb_rollup(vmb_ctx* ctx, vmb_series* series, const vmb_rollup_cfg* cfg, double* out, int out_is_device,
               uint64_t* samples_scanned);

/* ---- aggr(rollup(...)) by (...)  ==  evalRollupWithIncrementalAggregate eval.go:1804 + aggr_incremental.go ------ */
enum vmb_aggr_func { VMB_AGGR_SUM = 0, VMB_AGGR_MIN, VMB_AGGR_MAX, VMB_AGGR_AVG, VMB_AGGR_COUNT, VMB_AGGR_SUM2,
                     VMB_AGGR_GEOMEAN, VMB_AGGR_ANY, VMB_AGGR_GROUP };
/* Per-GPU partial state (the per-worker incrementalAggrContext, aggr_incremental.go:184): folds every series of the
 * batch into d_values/d_counts ([ngroups x P] DEVICE pointers, overwritten). group_ids: host, nseries entries, dense
 * ids assigned by the host from the group-by label set (identically on all ranks).  Within a group the series are
 * folded in ascending series order (deterministic).  d_rollup_scratch: device pointer to [nseries x P] doubles or
 * NULL to let the library allocate it. */
int vmb_rollup_aggr_partial(vmb_ctx* ctx, vmb_series* series, const vmb_rollup_cfg* cfg, int aggr_id,
                            const uint32_t* group_ids, uint32_t ngroups, double* d_values, double* d_counts,
                            double* d_rollup_scratch, uint64_t* samples_scanned);
/* mergeAggr* (aggr_incremental.go:218...): dst <- merge(dst, src), all DEVICE pointers, n = ngroups*P.
 * (Multi-GPU runs replace this by one NCCL all-reduce of values and counts for sum/avg/count/sum2, see DESIGN.md.) */
int vmb_aggr_merge(vmb_ctx* ctx, int aggr_id, double* d_dst_values, double* d_dst_counts, const double* d_src_values,
                   const double* d_src_counts, size_t n);
/* makes partial state all-reduce-able with ncclSum (values of empty cells zeroed) or ncclMin/ncclMax (+-Inf) */
int vmb_aggr_prepare_allreduce(vmb_ctx* ctx, int aggr_id, double* d_values, const double* d_counts, size_t n);
/* finalizeAggr* (aggr_incremental.go:189,:368,:400...): in place on DEVICE pointers, then optional copy to out_host */
int vmb_aggr_finalize(vmb_ctx* ctx, int aggr_id, double* d_values, const double* d_counts, size_t n, double* out_host);

/* ---- topk(k, q) / bottomk(k, q)  ==  newAggrFuncTopK aggr.go:646 on a DEVICE matrix d_vals[nseries x P] (e.g. the output
 * of vmb_rollup / vmb_eval_rollup_device): per group and point only the k best values survive, the others become NaN
 * (fillNaNsAtIdx aggr.go:786); rows left without a value are reported so that the host drops them (removeEmptySeries).
 * Exactly k values survive per (group, point); equal values rank by ascending GLOBAL series id (series_id_base + row), one of
 * the outcomes of the reference's unstable sort.Slice, identical on every run and rank.
 *   1. vmb_topk_candidates: d_cand[ngroups x P x kmax] entries of 16 bytes {f64 value, f64 global series id} <- the kmax best
 *      non-NaN values of this process per (group, point), best first, NaN padded; kmax = the largest k of the query, <= 64.
 *      reverse != 0: bottomk.  series_id_base: global id of this process's row 0 (0 on a single GPU).
 *   2. several processes: all-gather the candidate arrays (vmb_topk_allgather: count = cells * kmax * 2 doubles), then
 *      vmb_topk_merge([nparts x cells x kmax] entries, cells = ngroups*P).
 *   3. vmb_topk_apply: ks = one k per point (HOST, getIntK aggr.go:793: NaN / negative -> 0, capped by the group size);
 *      group_sizes = series per group over ALL processes (HOST); row_nonempty = HOST array, 1 byte per series. */
int vmb_topk_candidates(vmb_ctx* ctx, const double* d_vals, size_t nseries, size_t points, const uint32_t* group_ids,
                        uint32_t ngroups, uint32_t kmax, int reverse, uint64_t series_id_base, double* d_cand);
int vmb_topk_merge(vmb_ctx* ctx, const double* d_parts, uint32_t nparts, size_t cells, uint32_t kmax, int reverse, double* d_cand);
int vmb_topk_apply(vmb_ctx* ctx, double* d_vals, size_t nseries, size_t points, const uint32_t* group_ids, uint32_t ngroups,
                   const uint32_t* group_sizes, const double* d_cand, uint32_t kmax, const double* ks, int reverse,
                   uint64_t series_id_base, unsigned char* row_nonempty);

/* ---- whole path in one call with HOST buffers (what a patched evalRollupNoIncrementalAggregate, eval.go:1845,
 * would call): H2D of descriptors+payload, decode, preamble, rollup, D2H of the [nseries x P] result; processed in
 * chunks so copies overlap the kernels.  out_host: [nseries x P]. */
int vmb_eval_rollup_host(vmb_ctx* ctx, const vmb_block_desc* descs, size_t nblocks, const uint8_t* payload,
                         size_t payload_len, int64_t tr_min, int64_t tr_max, const vmb_rollup_cfg* cfg, double* out_host,
                         int32_t* block_status, uint64_t* samples_scanned);

/* device-resident variant used for kernel-only timing: decode + preamble + rollup, result left in d_out (device) */
int vmb_eval_rollup_device(vmb_ctx* ctx, const vmb_blocks* blocks, int64_t tr_min, int64_t tr_max,
                           const vmb_rollup_cfg* cfg, double* d_out, uint64_t* samples_scanned);

/* aggregate variant of the host path (evalRollupWithIncrementalAggregate eval.go:1804 end to end): the same chunked
 * pipeline as vmb_eval_rollup_host, but every chunk's [series x P] matrix is folded on the GPU into {values, counts}[G x P]
 * and only the finalized [ngroups x P] result (host pointer) travels back.  group_ids: one dense id per series of the batch. */
int vmb_eval_rollup_aggr_host(vmb_ctx* ctx, const vmb_block_desc* descs, size_t nblocks, const uint8_t* payload,
                              size_t payload_len, int64_t tr_min, int64_t tr_max, const vmb_rollup_cfg* cfg, int aggr_id,
                              const uint32_t* group_ids, uint32_t ngroups, double* out_host, int32_t* block_status,
                              uint64_t* samples_scanned);

/* ... without the final step: the partial state of this batch is left in the caller's DEVICE buffers {values, counts}
 * [ngroups x P] (overwritten), for vmb_aggr_merge or vmb_aggr_prepare_allreduce + all-reduce + vmb_aggr_finalize (multi-GPU) */
int vmb_eval_rollup_aggr_host_partial(vmb_ctx* ctx, const vmb_block_desc* descs, size_t nblocks, const uint8_t* payload,
                                      size_t payload_len, int64_t tr_min, int64_t tr_max, const vmb_rollup_cfg* cfg, int aggr_id,
                                      const uint32_t* group_ids, uint32_t ngroups, double* d_values, double* d_counts,
                                      int32_t* block_status, uint64_t* samples_scanned);

/* aggregate variant of the device-resident path: decode + preamble + rollup + vmb_rollup_aggr_partial in one call, decoded
 * columns cached in the library (per-rank step of `aggr(rollup(m[d])) by (...)`, eval.go:1804) */
int vmb_eval_rollup_aggr_device(vmb_ctx* ctx, const vmb_blocks* blocks, int64_t tr_min, int64_t tr_max,
                                const vmb_rollup_cfg* cfg, int aggr_id, const uint32_t* group_ids, uint32_t ngroups,
                                double* d_values, double* d_counts, uint64_t* samples_scanned);

/* ---- post-rollup operations on DEVICE matrices [series x points] (keep a query's intermediate results in HBM across the expression
 * tree) ------------------------------------------------------------------------------------------------------------------ */
enum vmb_binop { /* binary_op.go:15-43; the element functions of vendor/.../metricsql/binaryop/funcs.go */
    VMB_BO_PLUS = 0, VMB_BO_MINUS, VMB_BO_MUL, VMB_BO_DIV, VMB_BO_MOD, VMB_BO_POW, VMB_BO_ATAN2, VMB_BO_EQ, VMB_BO_NEQ, VMB_BO_GT,
    VMB_BO_LT, VMB_BO_GTE, VMB_BO_LTE, VMB_BO_DEFAULT, VMB_BO_IF, VMB_BO_IFNOT
};
/* newBinaryOpFunc binary_op.go:155-203: d_dst[i][j] = op(d_left[left_rows[i]][j], d_right[right_rows[i]][j]) for the npairs series
 * pairs the host matched by tag set (adjustBinaryOpTags :205).  left_rows / right_rows: HOST row indices or NULL = row i; a scalar
 * operand is a one-row matrix with all indices 0.  Comparisons: without is_bool they filter (left or NaN), with it they give 1 / 0
 * (NaN for a NaN left) -- newBinaryOpCmpFunc :132.  d_dst may alias d_left when left_rows is NULL. */
int vmb_binary_op(vmb_ctx* ctx, int op, int is_bool, const double* d_left, const uint32_t* left_rows, const double* d_right,
                  const uint32_t* right_rows, size_t npairs, size_t points, double* d_dst);
/* The set operators on top of it.  Their right-hand side is first reduced per tag-set group (createTimeseriesMapByTagSet binary_op.go:657
 * puts several series under one key): d_out[g][j] = the first non-NaN value among the rows with group_ids[row] == g at point j, in row
 * order, NaN when there is none -- all that addRightNaNsToLeft (:444), addLeftNaNsIfNoRightNaNs (:625) and fillLeftNaNsWithRightValues
 * (:516) read from tssRight.  Then, with right_rows[i] = the group of left row i:
 *   `and` (binaryOpAnd :430) / `if` = vmb_binary_op(VMB_BO_IF),  `unless` (:610) / `ifnot` = VMB_BO_IFNOT,  `default` = VMB_BO_DEFAULT;
 * left rows whose key has no right group are dropped (`and`) or kept as they are (`unless`, `default`) by the host's tag matching, which
 * also owns removeEmptySeries (exec.go:193).  `or` (:483) is a union of row sets plus the same fill: left rows, then the right rows of
 * keys the left side lacks. */
int vmb_group_first_value(vmb_ctx* ctx, const double* d_vals, size_t nseries, size_t points, const uint32_t* group_ids, uint32_t ngroups,
                          double* d_out);

/* Transform functions that only look at values (app/vmselect/promql/transform.go), in place on a DEVICE matrix [nrows x points]; labels,
 * sorting and the choice of series stay with the host.  Element functions (newTransformFuncOneArg :180 and friends) are applied to
 * every value, NaN included, like doTransformValues :195; row functions walk each series in point order like the reference (float
 * addition order is part of the result).  arg1 / arg2: HOST arrays of `points` values = getScalar of the scalar arguments:
 *   clamp(q, min, max): arg1 = min, arg2 = max;  clamp_min / clamp_max: arg1;  round(q, nearest): arg1 = nearest, arg2 =
 *   math.Pow10(-e) with (_, e) = decimal.FromFloat(nearest) (transform.go:2341; vmb_float_to_decimal gives e).  Others: NULL.
 * exp / ln / log2 / log10 / trigonometric / hyperbolic functions are the CUDA math library's (<= 2 ulp from Go's); everything else is
 * bit-exact. */
enum vmb_transform_func {
    VMB_TF_ABS = 0, VMB_TF_CEIL, VMB_TF_FLOOR, VMB_TF_SQRT, VMB_TF_EXP, VMB_TF_LN, VMB_TF_LOG2, VMB_TF_LOG10, VMB_TF_SIN, VMB_TF_COS,
    VMB_TF_TAN, VMB_TF_ASIN, VMB_TF_ACOS, VMB_TF_ATAN, VMB_TF_SINH, VMB_TF_COSH, VMB_TF_TANH, VMB_TF_ASINH, VMB_TF_ACOSH, VMB_TF_ATANH,
    VMB_TF_DEG, VMB_TF_RAD, VMB_TF_SGN, VMB_TF_CLAMP, VMB_TF_CLAMP_MIN, VMB_TF_CLAMP_MAX, VMB_TF_ROUND,
    /* row functions: running_* (:1308), range_* = running + setLastValues (:1335, :1650), range_first / range_last (:1620, :1640),
     * keep_last_value / keep_next_value (:1214, :1237), remove_resets = removeCounterResetsMaybeNaNs (:2906) */
    VMB_TF_RUNNING_SUM = 32, VMB_TF_RUNNING_MIN, VMB_TF_RUNNING_MAX, VMB_TF_RUNNING_AVG, VMB_TF_RANGE_SUM, VMB_TF_RANGE_MIN,
    VMB_TF_RANGE_MAX, VMB_TF_RANGE_AVG, VMB_TF_RANGE_FIRST, VMB_TF_RANGE_LAST, VMB_TF_KEEP_LAST_VALUE, VMB_TF_KEEP_NEXT_VALUE,
    VMB_TF_REMOVE_RESETS, VMB_TF_INTERPOLATE /* :1261 */
};
int vmb_transform(vmb_ctx* ctx, int func, double* d_matrix, size_t nrows, size_t points, const double* arg1, const double* arg2);
/* mergeSeries rollup_result_cache.go:618: d_dst[nrows x (pa + pb)], row i = d_a[a_rows[i]] ++ d_b[b_rows[i]]; a negative index
 * stands for a series missing on that side (NaNs, :677-690).  a_rows / b_rows: HOST, matched by metric name by the caller. */
int vmb_matrix_merge_rows(vmb_ctx* ctx, const double* d_a, const int64_t* a_rows, size_t pa, const double* d_b, const int64_t* b_rows,
                          size_t pb, size_t nrows, double* d_dst);
/* quantile(phi, q) by (...) / median(q) by (...)  aggr.go:1217-1240 newAggrQuantileFunc: d_out[ngroups x P], phis = one phi per
 * point (HOST).  Per (group, point) the NaNs are dropped and aggr.go:870 quantile applies; rank selection, quadratic in the group
 * size: groups of more than 2048 series return VMB_ERR_CAP. */
int vmb_aggr_quantile(vmb_ctx* ctx, const double* d_vals, size_t nseries, size_t points, const uint32_t* group_ids, uint32_t ngroups,
                      const double* phis, double* d_out);

/* ---- multi-GPU: one process per GPU, the ONE exchange step of the path inside the library ------------------
 * aggr(rollup(m[d])) by (...): every rank folds its shard of the series into {values, counts}[G x P] (the per-worker
 * incrementalAggrContext, aggr_incremental.go:184), the partial states are merged by one ncclAllReduce per array -- the GPU
 * counterpart of the merge loop in finalizeTimeseries (aggr_incremental.go:141-168) -- and finalized on every rank.
 * NCCL is dlopen()ed ("libnccl.so.2", or $VMB_NCCL_LIB): no link-time dependency; in a process that already holds an NCCL it is
 * the same library instance.  Bootstrap: rank 0 calls vmb_comm_get_unique_id and hands the 128 bytes to the other ranks by any
 * means (the Go host: over vmselect's own RPC), then every rank calls vmb_ctx_comm_init; or vmb_ctx_comm_attach with an
 * ncclComm_t the host created itself.  Group ids must be assigned identically on all ranks. */
int vmb_comm_get_unique_id(uint8_t id[128]);                                  /* ncclGetUniqueId */
int vmb_ctx_comm_init(vmb_ctx* ctx, const uint8_t id[128], int nranks, int rank); /* ncclCommInitRank on the ctx's device */
int vmb_ctx_comm_attach(vmb_ctx* ctx, void* nccl_comm, int nranks, int rank);  /* the host's own ncclComm_t (not destroyed by the ctx) */
int vmb_ctx_comm_destroy(vmb_ctx* ctx);
int vmb_ctx_comm_size(const vmb_ctx* ctx);                                     /* 1 without a communicator */
int vmb_ctx_comm_rank(const vmb_ctx* ctx);
/* exchange step on DEVICE partial states, in place, on the ctx stream: identity into empty cells, all-reduce of the values with
 * the aggregate's operator (sum / min / max / prod) and of the counts with sum.  No-op without a communicator. */
int vmb_aggr_allreduce(vmb_ctx* ctx, int aggr_id, double* d_values, double* d_counts, size_t n);
/* topk(): candidate lists of all ranks side by side, d_parts[nranks x count] (input of vmb_topk_merge) -- ncclAllGather */
int vmb_topk_allgather(vmb_ctx* ctx, const double* d_cand, size_t count, double* d_parts);
/* the whole query step on every rank in one call: fold this rank's device-resident blocks, all-reduce, finalize;
 * out_host [ngroups x P] (may be NULL) receives the (identical on all ranks) result */
int vmb_eval_rollup_aggr_dist(vmb_ctx* ctx, const vmb_blocks* blocks, int64_t tr_min, int64_t tr_max, const vmb_rollup_cfg* cfg,
                              int aggr_id, const uint32_t* group_ids, uint32_t ngroups, double* out_host, uint64_t* samples_scanned);

/* pinned host memory helpers (cudaHostAlloc) for callers that want full PCIe speed */
void* vmb_host_alloc(size_t bytes);
void vmb_host_free(void* p);

/* timing hooks: elapsed device time (ms) of the named stage during the last batched call on this ctx;
 * stage: 0 = zstd, 1 = column decode, 2 = series preamble, 3 = rollup, 4 = aggregate, 5 = fused decode+rollup kernel
 * (with the fused kernel on, stage 1 is the whole un-fused sub-batch of the series it did not take, stages 2-3 are 0) */
float vmb_ctx_last_stage_ms(const vmb_ctx* ctx, int stage);
int vmb_ctx_enable_stage_timing(vmb_ctx* ctx, int enable);

#ifdef __cplusplus
}
#endif
#endif /* VMB200_H */
