// vmb200.hpp -- C++ host-side mirror of the reference's Go API for the hot path, on top of the C ABI (vmb200.h).
//
// The reference host code is Go; there is no Go toolchain in this build environment, so the host mirror is written in
// C++ (and in Python/ctypes for the tests: victoriametrics_b200/*.py).  Names, argument meaning and error behaviour follow
// the Go functions cited at each declaration; "append to dst" slices become std::vector&.  Header-only; link -lvmb200.
#ifndef VMB200_HPP
#define VMB200_HPP
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "vmb200.h"

namespace vmb {

struct Error : std::runtime_error {  // the Go functions return `error`; Panicf("BUG: ...") sites become VMB_ERR_INVALID_ARG
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what + ": " + vmb_last_error()), code(c) {}
};

class Ctx {  // one per process per GPU
   public:
    explicit Ctx(int device = 0) {
        int rc = vmb_ctx_create(device, &p_);
        if (rc) throw Error(rc, "vmb_ctx_create");
    }
    ~Ctx() { vmb_ctx_destroy(p_); }
    Ctx(const Ctx&) = delete;
    Ctx& operator=(const Ctx&) = delete;
    vmb_ctx* get() const { return p_; }

   private:
    vmb_ctx* p_ = nullptr;
};

namespace encoding {  // lib/encoding/encoding.go

enum MarshalType : uint8_t {  // encoding.go:20-43
    MarshalTypeZSTDNearestDelta2 = 1,
    MarshalTypeDeltaConst = 2,
    MarshalTypeConst = 3,
    MarshalTypeZSTDNearestDelta = 4,
    MarshalTypeNearestDelta2 = 5,
    MarshalTypeNearestDelta = 6,
};

struct Marshaled {
    MarshalType mt;
    int64_t firstValue;
};

// MarshalValues encoding.go:103 (MarshalTimestamps :82 has the same body): appends the marshaled bytes to dst
inline Marshaled MarshalValues(std::vector<uint8_t>& dst, const std::vector<int64_t>& values, uint8_t precisionBits) {
    if (values.empty()) throw Error(VMB_ERR_INVALID_ARG, "BUG: a must contain at least one item");  // encoding.go:121
    size_t off = dst.size(), len = 0;
    dst.resize(off + values.size() * 10 + 1024);
    int mt = 0;
    int64_t first = 0;
    int rc = vmb_marshal_int64(dst.data() + off, dst.size() - off, &len, &mt, &first, values.data(), values.size(), precisionBits);
    if (rc) throw Error(rc, "MarshalValues");
    dst.resize(off + len);
    return {(MarshalType)mt, first};
}
inline Marshaled MarshalTimestamps(std::vector<uint8_t>& dst, const std::vector<int64_t>& ts, uint8_t precisionBits) {
    return MarshalValues(dst, ts, precisionBits);
}

// UnmarshalValues encoding.go:111 (UnmarshalTimestamps :90): appends itemsCount values to dst; returns 0 or the error code
// (the Go code returns (nil, err)); dst is left untouched on error.
inline int UnmarshalValues(Ctx& ctx, std::vector<int64_t>& dst, const uint8_t* src, size_t srcLen, MarshalType mt,
                           int64_t firstValue, int itemsCount) {
    size_t off = dst.size();
    dst.resize(off + (size_t)itemsCount);
    int rc = vmb_unmarshal_int64(ctx.get(), dst.data() + off, (size_t)itemsCount, src, srcLen, (int)mt, firstValue);
    if (rc) dst.resize(off);
    return rc;
}
inline int UnmarshalTimestamps(Ctx& ctx, std::vector<int64_t>& dst, const uint8_t* src, size_t srcLen, MarshalType mt,
                               int64_t firstTimestamp, int itemsCount) {
    return UnmarshalValues(ctx, dst, src, srcLen, mt, firstTimestamp, itemsCount);
}

}  // namespace encoding

namespace decimal {  // lib/decimal/decimal.go

// AppendDecimalToFloat decimal.go:100
inline void AppendDecimalToFloat(Ctx& ctx, std::vector<double>& dst, const std::vector<int64_t>& va, int16_t e) {
    size_t off = dst.size();
    dst.resize(off + va.size());
    int rc = vmb_decimal_to_float(ctx.get(), dst.data() + off, va.data(), va.size(), e);
    if (rc) throw Error(rc, "AppendDecimalToFloat");
}
// AppendFloatToDecimal decimal.go:173 -> exponent
inline int16_t AppendFloatToDecimal(std::vector<int64_t>& dst, const std::vector<double>& src) {
    size_t off = dst.size();
    dst.resize(off + src.size());
    int16_t e = 0;
    int rc = vmb_float_to_decimal(dst.data() + off, &e, src.data(), src.size());
    if (rc) throw Error(rc, "AppendFloatToDecimal");
    return e;
}

}  // namespace decimal

namespace promql {  // app/vmselect/promql

// rollupFuncs rollup.go:24-108 -> enum vmb_rollup_func
inline int rollupFuncId(const std::string& name) {
    static const std::unordered_map<std::string, int> m = {
        {"default_rollup", VMB_RF_DEFAULT_ROLLUP}, {"rate", VMB_RF_RATE}, {"deriv_fast", VMB_RF_RATE}, {"delta", VMB_RF_DELTA},
        {"increase", VMB_RF_DELTA}, {"avg_over_time", VMB_RF_AVG}, {"min_over_time", VMB_RF_MIN}, {"max_over_time", VMB_RF_MAX},
        {"sum_over_time", VMB_RF_SUM}, {"count_over_time", VMB_RF_COUNT}, {"quantile_over_time", VMB_RF_QUANTILE},
        {"first_over_time", VMB_RF_FIRST}, {"last_over_time", VMB_RF_LAST}, {"range_over_time", VMB_RF_RANGE},
        {"sum2_over_time", VMB_RF_SUM2}, {"stddev_over_time", VMB_RF_STDDEV}, {"stdvar_over_time", VMB_RF_STDVAR},
        {"ideriv", VMB_RF_IDERIV}, {"irate", VMB_RF_IDERIV}, {"idelta", VMB_RF_IDELTA}, {"deriv", VMB_RF_DERIV},
        {"increase_pure", VMB_RF_INCREASE_PURE}, {"changes", VMB_RF_CHANGES}, {"changes_prometheus", VMB_RF_CHANGES_PROMETHEUS},
        {"resets", VMB_RF_RESETS}, {"decreases_over_time", VMB_RF_RESETS}, {"increases_over_time", VMB_RF_INCREASES},
        {"integrate", VMB_RF_INTEGRATE}, {"lag", VMB_RF_LAG}, {"lifetime", VMB_RF_LIFETIME},
        {"scrape_interval", VMB_RF_SCRAPE_INTERVAL}, {"tmin_over_time", VMB_RF_TMIN}, {"tmax_over_time", VMB_RF_TMAX},
        {"tfirst_over_time", VMB_RF_TFIRST}, {"tlast_over_time", VMB_RF_TLAST}, {"timestamp", VMB_RF_TLAST},
        {"timestamp_with_name", VMB_RF_TLAST}, {"tlast_change_over_time", VMB_RF_TLAST_CHANGE}, {"mode_over_time", VMB_RF_MODE},
        {"mad_over_time", VMB_RF_MAD}, {"outlier_iqr_over_time", VMB_RF_OUTLIER_IQR}, {"zscore_over_time", VMB_RF_ZSCORE},
        {"ascent_over_time", VMB_RF_ASCENT}, {"descent_over_time", VMB_RF_DESCENT}, {"distinct_over_time", VMB_RF_DISTINCT},
        {"geomean_over_time", VMB_RF_GEOMEAN}, {"predict_linear", VMB_RF_PREDICT_LINEAR}, {"holt_winters", VMB_RF_HOLT_WINTERS},
        {"hoeffding_bound_lower", VMB_RF_HOEFFDING_LOWER}, {"hoeffding_bound_upper", VMB_RF_HOEFFDING_UPPER},
        {"duration_over_time", VMB_RF_DURATION}, {"count_le_over_time", VMB_RF_COUNT_LE}, {"count_gt_over_time", VMB_RF_COUNT_GT},
        {"count_eq_over_time", VMB_RF_COUNT_EQ}, {"count_ne_over_time", VMB_RF_COUNT_NE}, {"share_le_over_time", VMB_RF_SHARE_LE},
        {"share_gt_over_time", VMB_RF_SHARE_GT}, {"share_eq_over_time", VMB_RF_SHARE_EQ}, {"sum_le_over_time", VMB_RF_SUM_LE},
        {"sum_gt_over_time", VMB_RF_SUM_GT}, {"sum_eq_over_time", VMB_RF_SUM_EQ}, {"present_over_time", VMB_RF_PRESENT},
        {"absent_over_time", VMB_RF_ABSENT}, {"stale_samples_over_time", VMB_RF_STALE_SAMPLES}, {"median_over_time", VMB_RF_MEDIAN},
        {"rate_over_sum", VMB_RF_RATE_OVER_SUM}, {"delta_prometheus", VMB_RF_DELTA_PROMETHEUS},
        {"increase_prometheus", VMB_RF_DELTA_PROMETHEUS}, {"rate_prometheus", VMB_RF_RATE_PROMETHEUS}};
    auto it = m.find(name);
    return it == m.end() ? -1 : it->second;
}

// getTimestamps eval.go:230
inline std::vector<int64_t> getTimestamps(int64_t start, int64_t end, int64_t step) {
    if (step <= 0 || start > end) throw Error(VMB_ERR_INVALID_ARG, "BUG: invalid start/end/step");
    std::vector<int64_t> ts((size_t)(1 + (end - start) / step));
    for (size_t i = 0; i < ts.size(); i++) ts[i] = start + (int64_t)i * step;
    return ts;
}

// rollupConfig rollup.go:574
struct rollupConfig {
    std::string Func;  // MetricsQL function name
    int64_t Start = 0, End = 0, Step = 0, Window = 0;
    bool MayAdjustWindow = false;
    int64_t LookbackDelta = 0;
    bool isDefaultRollup = false;
    int samplesScannedPerCall = 0;
    bool removeCounterResets = false;  // preFunc of eval.go:1855 (rollupFuncsRemoveCounterResets rollup.go:223)
    bool dropStaleNaNs = false;        // eval.go:1985
    int64_t minStalenessInterval = 0;  // -search.minStalenessInterval rollup.go:20
    std::vector<double> args, args2;   // per-point scalar args (phi / limit / secs / sf, tf) or empty
    std::vector<int64_t> Timestamps;

    vmb_rollup_cfg cfg() const {
        vmb_rollup_cfg c{};
        c.func_id = rollupFuncId(Func);
        if (c.func_id < 0) throw Error(VMB_ERR_INVALID_ARG, "unknown rollup func " + Func);
        c.flags = (MayAdjustWindow ? VMB_RC_MAY_ADJUST_WINDOW : 0u) | (isDefaultRollup ? VMB_RC_IS_DEFAULT_ROLLUP : 0u) |
                  (removeCounterResets ? VMB_RC_REMOVE_COUNTER_RESETS : 0u) | (dropStaleNaNs ? VMB_RC_DROP_STALE_NANS : 0u);
        c.start = Start;
        c.end = End;
        c.step = Step;
        c.window = Window;
        c.lookback_delta = LookbackDelta;
        c.min_staleness_ms = minStalenessInterval;
        c.samples_scanned_per_call = samplesScannedPerCall;
        c.args = args.empty() ? nullptr : args.data();
        c.args2 = args2.empty() ? nullptr : args2.data();
        return c;
    }

    // Do rollup.go:688: appends len(Timestamps) values to dstValues, returns samplesScanned.  One series per call (kept for
    // compatibility; use evalRollupFunc for whole queries).
    uint64_t Do(Ctx& ctx, std::vector<double>& dstValues, const std::vector<double>& values,
                const std::vector<int64_t>& timestamps) const {
        vmb_rollup_cfg c = cfg();
        int64_t points = vmb_rollup_points(&c);
        if (points < 0) throw Error(VMB_ERR_INVALID_ARG, "BUG: invalid rollupConfig");  // rollup.go:703-714
        uint64_t offs[2] = {0, values.size()};
        vmb_series* s = nullptr;
        int rc = vmb_series_from_host(ctx.get(), timestamps.data(), values.data(), offs, 1, &s);
        if (rc) throw Error(rc, "vmb_series_from_host");
        size_t off = dstValues.size();
        dstValues.resize(off + (size_t)points);
        uint64_t scanned = 0;
        rc = vmb_rollup(ctx.get(), s, &c, dstValues.data() + off, 0, &scanned);
        vmb_series_free(s);
        if (rc) throw Error(rc, "rollupConfig.Do");
        return scanned;
    }
};

// getRollupConfigs rollup.go:374 (single-config functions) + the preFunc / dropStaleNaNs decisions of eval.go:1855, :1985
inline rollupConfig getRollupConfigs(const std::string& funcName, int64_t start, int64_t end, int64_t step, int64_t window,
                                     int64_t lookbackDelta) {
    static const std::unordered_set<std::string> canAdjust = {"default_rollup", "deriv", "deriv_fast", "ideriv", "irate", "rate",
                                                              "rate_over_sum", "scrape_interval", "timestamp"};  // rollup.go:199
    static const std::unordered_set<std::string> removeResets = {"increase", "increase_prometheus", "increase_pure", "irate",
                                                                 "rate", "rate_prometheus"};  // rollup.go:223
    static const std::unordered_map<std::string, int> perCall = {  // rollup.go:238
        {"absent_over_time", 1}, {"count_over_time", 1}, {"default_rollup", 1}, {"delta", 2}, {"delta_prometheus", 2},
        {"deriv_fast", 2}, {"first_over_time", 1}, {"idelta", 2}, {"ideriv", 2}, {"increase", 2}, {"increase_prometheus", 2},
        {"increase_pure", 2}, {"irate", 2}, {"lag", 1}, {"last_over_time", 1}, {"lifetime", 2}, {"present_over_time", 1},
        {"rate", 2}, {"rate_prometheus", 2}, {"scrape_interval", 2}, {"tfirst_over_time", 1}, {"timestamp", 1},
        {"timestamp_with_name", 1}, {"tlast_over_time", 1}};
    rollupConfig rc;
    rc.Func = funcName;
    rc.Start = start;
    rc.End = end;
    rc.Step = step;
    rc.Window = window;
    rc.LookbackDelta = lookbackDelta;
    rc.MayAdjustWindow = canAdjust.count(funcName) != 0;
    rc.isDefaultRollup = funcName == "default_rollup";
    auto it = perCall.find(funcName);
    rc.samplesScannedPerCall = it == perCall.end() ? 0 : it->second;
    rc.removeCounterResets = removeResets.count(funcName) != 0;
    rc.dropStaleNaNs = !(funcName == "default_rollup" || funcName == "stale_samples_over_time");
    rc.Timestamps = getTimestamps(start, end, step);
    return rc;
}

// evalRollupFuncNoCache eval.go:1680 -> evalRollupNoIncrementalAggregate eval.go:1845: all blocks of the query at once.
// out: [nseries x len(rc.Timestamps)] row-major; returns samplesScanned.
inline uint64_t evalRollupFunc(Ctx& ctx, const rollupConfig& rc, const std::vector<vmb_block_desc>& descs,
                               const std::vector<uint8_t>& payload, int64_t trMin, int64_t trMax, std::vector<double>& out,
                               size_t nseries) {
    vmb_rollup_cfg c = rc.cfg();
    out.resize(nseries * rc.Timestamps.size());
    uint64_t scanned = 0;
    int r = vmb_eval_rollup_host(ctx.get(), descs.data(), descs.size(), payload.data(), payload.size(), trMin, trMax, &c,
                                 out.data(), nullptr, &scanned);
    if (r) throw Error(r, "evalRollupFunc");
    return scanned;
}

// evalRollupWithIncrementalAggregate eval.go:1804 (aggr_incremental.go): aggr(rollup(m[d])) by (...) for all blocks of the
// query at once; groupIDs = one dense id per series (marshalMetricNameSorted of the group-by labels, assigned by the caller).
// out: [ngroups x len(rc.Timestamps)] row-major; returns samplesScanned.
inline int aggrFuncId(const std::string& name) {
    static const std::unordered_map<std::string, int> ids = {{"sum", VMB_AGGR_SUM}, {"min", VMB_AGGR_MIN}, {"max", VMB_AGGR_MAX},
                                                             {"avg", VMB_AGGR_AVG}, {"count", VMB_AGGR_COUNT}, {"sum2", VMB_AGGR_SUM2},
                                                             {"geomean", VMB_AGGR_GEOMEAN}, {"any", VMB_AGGR_ANY},
                                                             {"group", VMB_AGGR_GROUP}};
    auto it = ids.find(name);
    if (it == ids.end()) throw Error(VMB_ERR_INVALID_ARG, "aggregate without incremental form: " + name);
    return it->second;
}
inline uint64_t evalRollupFuncWithIncrementalAggregate(Ctx& ctx, const std::string& aggrName, const rollupConfig& rc,
                                                       const std::vector<vmb_block_desc>& descs, const std::vector<uint8_t>& payload,
                                                       const std::vector<uint32_t>& groupIDs, uint32_t ngroups, int64_t trMin,
                                                       int64_t trMax, std::vector<double>& out) {
    vmb_rollup_cfg c = rc.cfg();
    out.resize((size_t)ngroups * rc.Timestamps.size());
    uint64_t scanned = 0;
    int r = vmb_eval_rollup_aggr_host(ctx.get(), descs.data(), descs.size(), payload.data(), payload.size(), trMin, trMax, &c,
                                      aggrFuncId(aggrName), groupIDs.data(), ngroups, out.data(), nullptr, &scanned);
    if (r) throw Error(r, "evalRollupFuncWithIncrementalAggregate");
    return scanned;
}

}  // namespace promql
}  // namespace vmb
#endif
