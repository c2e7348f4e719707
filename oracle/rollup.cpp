// ORACLE (test infrastructure, not product code): CPU restatement of the reference rollup executor.
// Follows /root/reference/app/vmselect/promql/rollup.go (rollupConfig.doInternal :701, removeCounterResets :921,
// the rollup funcs :1030-2445), aggr.go (quantile :870, modeNoNaNs :541), aggr_incremental.go (:189-458) and
// eval.go (dropStaleNaNs :1985).  Sequential, same evaluation order as the Go code so that results are bit-comparable
// with the reference's own test vectors (tests/test_oracle_rollup.py transcribes rollup_test.go).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "vm_oracle.h"

namespace {

const double kNaN = NAN;

struct Rfa {  // rollupFuncArg rollup.go:523
    double prevValue;
    int64_t prevTimestamp;
    const double* values;
    const int64_t* timestamps;
    size_t n;
    double realPrevValue;
    double realNextValue;
    int64_t currTimestamp;
    size_t idx;
    int64_t window;
    const double* args;
    const double* args2;
};

inline bool is_stale_nan(double f) {
    uint64_t b;
    memcpy(&b, &f, 8);
    return b == 0x7ff0000000000002ULL;
}

// quantileSorted aggr.go:922
double quantile_sorted(double phi, const double* v, size_t n) {
    if (n == 0 || isnan(phi)) return kNaN;
    if (phi < 0) return -INFINITY;
    if (phi > 1) return INFINITY;
    double nn = (double)n;
    double rank = phi * (nn - 1);
    double lower = fmax(0, floor(rank));
    double upper = fmin(nn - 1, lower + 1);
    double weight = rank - floor(rank);
    return v[(int)lower] * (1 - weight) + v[(int)upper] * weight;
}

// quantile aggr.go:870 (+ prepareForQuantileFloat64 :879)
double quantile(double phi, const double* values, size_t n) {
    std::vector<double> a;
    a.reserve(n);
    for (size_t i = 0; i < n; i++)
        if (!isnan(values[i])) a.push_back(values[i]);
    std::sort(a.begin(), a.end());
    return quantile_sorted(phi, a.data(), a.size());
}

// modeNoNaNs aggr.go:541
double mode_no_nans(double prevValue, double* a, size_t n) {
    if (n == 0) return prevValue;
    std::sort(a, a + n);
    ptrdiff_t j = -1;
    ptrdiff_t dMax = 0;
    double mode = prevValue;
    for (size_t i = 0; i < n; i++) {
        double v = a[i];
        if (prevValue == v) continue;
        ptrdiff_t d = (ptrdiff_t)i - j;
        if (d > dMax || isnan(mode)) {
            dMax = d;
            mode = prevValue;
        }
        j = (ptrdiff_t)i;
        prevValue = v;
    }
    ptrdiff_t d = (ptrdiff_t)n - j;
    if (d > dMax || isnan(mode)) mode = prevValue;
    return mode;
}

bool are_const_values(const double* v, size_t n) {  // rollup.go:1136
    if (n <= 1) return true;
    double p = v[0];
    for (size_t i = 1; i < n; i++) {
        if (v[i] != p) return false;
        p = v[i];
    }
    return true;
}

// linearRegression rollup.go:1099
void linear_regression(const double* values, const int64_t* ts, size_t n, int64_t intercept, double* vout, double* kout) {
    if (n == 0) {
        *vout = kNaN;
        *kout = kNaN;
        return;
    }
    if (are_const_values(values, n)) {
        *vout = values[0];
        *kout = 0;
        return;
    }
    double vSum = 0, tSum = 0, tvSum = 0, ttSum = 0;
    int cnt = 0;
    for (size_t i = 0; i < n; i++) {
        double v = values[i];
        if (isnan(v)) continue;
        double dt = (double)(ts[i] - intercept) / 1e3;
        vSum += v;
        tSum += dt;
        tvSum += dt * v;
        ttSum += dt * dt;
        cnt++;
    }
    if (cnt == 0) {
        *vout = kNaN;
        *kout = kNaN;
        return;
    }
    double k = 0;
    double tDiff = ttSum - tSum * tSum / (double)cnt;
    if (fabs(tDiff) >= 1e-6) k = (tvSum - tSum * vSum / (double)cnt) / tDiff;
    *vout = vSum / (double)cnt - k * tSum / (double)cnt;
    *kout = k;
}

double stdvar(const double* values, size_t n) {  // rollup.go:1808
    if (n == 0) return kNaN;
    if (n == 1) return 0;
    double avg = 0, count = 0, q = 0;
    for (size_t i = 0; i < n; i++) {
        double v = values[i];
        if (isnan(v)) continue;
        count++;
        double avgNew = avg + (v - avg) / count;
        q += (v - avg) * (v - avgNew);
        avg = avgNew;
    }
    if (count == 0) return kNaN;
    return q / count;
}

double r_avg(const Rfa& r) {  // rollup.go:1541
    if (r.n == 0) return kNaN;
    double sum = 0;
    for (size_t i = 0; i < r.n; i++) sum += r.values[i];
    return sum / (double)r.n;
}
double r_min(const Rfa& r) {  // :1561
    if (r.n == 0) return kNaN;
    double m = r.values[0];
    for (size_t i = 0; i < r.n; i++)
        if (r.values[i] < m) m = r.values[i];
    return m;
}
double r_max(const Rfa& r) {  // :1580
    if (r.n == 0) return kNaN;
    double m = r.values[0];
    for (size_t i = 0; i < r.n; i++)
        if (r.values[i] > m) m = r.values[i];
    return m;
}
double r_sum(const Rfa& r) {  // :1690
    if (r.n == 0) return kNaN;
    double sum = 0;
    for (size_t i = 0; i < r.n; i++) sum += r.values[i];
    return sum;
}
double r_last(const Rfa& r) { return r.n == 0 ? kNaN : r.values[r.n - 1]; }  // rollupDefault :2390

double r_lag(const Rfa& r) {  // :2055
    if (r.n == 0) {
        if (isnan(r.prevValue)) return kNaN;
        return (double)(r.currTimestamp - r.prevTimestamp) / 1e3;
    }
    return (double)(r.currTimestamp - r.timestamps[r.n - 1]) / 1e3;
}
double r_scrape_interval(const Rfa& r) {  // :2067
    if (isnan(r.prevValue)) {
        if (r.n < 2) return kNaN;
        return ((double)(r.timestamps[r.n - 1] - r.timestamps[0]) / 1e3) / (double)(r.n - 1);
    }
    if (r.n == 0) return kNaN;
    return ((double)(r.timestamps[r.n - 1] - r.prevTimestamp) / 1e3) / (double)r.n;
}

double r_delta(const Rfa& r) {  // rollupDelta :1859
    const double* values = r.values;
    size_t n = r.n;
    double prevValue = r.prevValue;
    if (isnan(prevValue)) {
        if (n == 0) return kNaN;
        if (!isnan(r.realPrevValue)) return values[n - 1] - r.realPrevValue;
        double d = 0;
        if (n > 1) d = values[1] - values[0];
        else if (!isnan(r.realNextValue)) d = r.realNextValue - values[0];
        if (fabs(values[0]) < 10 * (fabs(d) + 1)) {
            prevValue = 0;
        } else {
            prevValue = values[0];
            values++;
            n--;
        }
    }
    if (n == 0) return 0;
    return values[n - 1] - prevValue;
}

double r_deriv_fast(const Rfa& r) {  // rollupDerivFast :1954
    double prevValue = r.prevValue;
    int64_t prevTimestamp = r.prevTimestamp;
    if (isnan(prevValue)) {
        if (r.n == 0) return kNaN;
        if (r.n == 1) return kNaN;
        prevValue = r.values[0];
        prevTimestamp = r.timestamps[0];
    } else if (r.n == 0) {
        return 0;
    }
    double vEnd = r.values[r.n - 1];
    int64_t tEnd = r.timestamps[r.n - 1];
    double dv = vEnd - prevValue;
    double dt = (double)(tEnd - prevTimestamp) / 1e3;
    return dv / dt;
}

double r_ideriv(const Rfa& r) {  // rollupIderiv :1991
    const double* values = r.values;
    const int64_t* timestamps = r.timestamps;
    size_t n = r.n;
    if (n < 2) {
        if (n == 0) return kNaN;
        if (isnan(r.prevValue)) return kNaN;
        return (values[0] - r.prevValue) / ((double)(timestamps[0] - r.prevTimestamp) / 1e3);
    }
    double vEnd = values[n - 1];
    int64_t tEnd = timestamps[n - 1];
    size_t m = n - 1;  // len(values) == len(timestamps) == m
    size_t tn = m;
    while (tn > 0 && timestamps[tn - 1] >= tEnd) tn--;
    int64_t tStart;
    double vStart;
    if (tn == 0) {
        if (isnan(r.prevValue)) return 0;
        tStart = r.prevTimestamp;
        vStart = r.prevValue;
    } else {
        tStart = timestamps[tn - 1];
        vStart = values[tn - 1];
    }
    double dv = vEnd - vStart;
    int64_t dt = tEnd - tStart;
    return dv / ((double)dt / 1e3);
}

double r_idelta(const Rfa& r) {  // :1915
    if (r.n == 0) {
        if (isnan(r.prevValue)) return kNaN;
        return 0;
    }
    double last = r.values[r.n - 1];
    if (r.n == 1) {
        if (isnan(r.prevValue)) return last;
        return last - r.prevValue;
    }
    return last - r.values[r.n - 2];
}

double r_increase_pure(const Rfa& r) {  // :1835
    double prevValue = r.prevValue;
    if (isnan(prevValue)) {
        if (r.n == 0) return kNaN;
        prevValue = 0;
        if (!isnan(r.realPrevValue)) prevValue = r.realPrevValue;
    }
    if (r.n == 0) return 0;
    return r.values[r.n - 1] - prevValue;
}

double r_changes_prometheus(const Rfa& r) {  // :2080
    if (r.n < 1) return kNaN;
    double prev = r.values[0];
    int cnt = 0;
    for (size_t i = 1; i < r.n; i++) {
        double v = r.values[i];
        if (v != prev) {
            if (fabs(v - prev) < 1e-12 * fabs(v)) continue;
            cnt++;
            prev = v;
        }
    }
    return (double)cnt;
}

double r_changes(const Rfa& r) {  // :2106
    const double* values = r.values;
    size_t n = r.n;
    double prev = r.prevValue;
    int cnt = 0;
    if (isnan(prev)) {
        if (n == 0) return kNaN;
        if (!isnan(r.realPrevValue)) {
            prev = r.realPrevValue;
        } else {
            cnt++;
            prev = values[0];
            values++;
            n--;
        }
    }
    for (size_t i = 0; i < n; i++) {
        double v = values[i];
        if (v != prev) {
            if (fabs(v - prev) < 1e-12 * fabs(v)) continue;
            cnt++;
            prev = v;
        }
    }
    return (double)cnt;
}

double r_increases_or_resets(const Rfa& r, bool increases) {  // rollupIncreases :2139 / rollupResets :2174
    const double* values = r.values;
    size_t n = r.n;
    if (n == 0) {
        if (isnan(r.prevValue)) return kNaN;
        return 0;
    }
    double prev = r.prevValue;
    if (isnan(prev)) {
        prev = values[0];
        values++;
        n--;
    }
    if (n == 0) return 0;
    int cnt = 0;
    for (size_t i = 0; i < n; i++) {
        double v = values[i];
        bool hit = increases ? (v > prev) : (v < prev);
        if (hit) {
            if (fabs(v - prev) < 1e-12 * fabs(v)) continue;  // NB: Go `continue` skips prevValue = v
            cnt++;
        }
        prev = v;
    }
    return (double)cnt;
}

double r_integrate(const Rfa& r) {  // :2417
    const double* values = r.values;
    const int64_t* timestamps = r.timestamps;
    size_t n = r.n;
    double prevValue = r.prevValue;
    int64_t prevTimestamp = r.currTimestamp - r.window;
    if (isnan(prevValue)) {
        if (n == 0) return kNaN;
        prevValue = values[0];
        prevTimestamp = timestamps[0];
        values++;
        timestamps++;
        n--;
    }
    double sum = 0;
    for (size_t i = 0; i < n; i++) {
        int64_t t = timestamps[i];
        double dt = (double)(t - prevTimestamp) / 1e3;
        sum += prevValue * dt;
        prevTimestamp = t;
        prevValue = values[i];
    }
    double dt = (double)(r.currTimestamp - prevTimestamp) / 1e3;
    sum += prevValue * dt;
    return sum;
}

double r_lifetime(const Rfa& r) {  // :2040
    if (isnan(r.prevValue)) {
        if (r.n < 2) return kNaN;
        return (double)(r.timestamps[r.n - 1] - r.timestamps[0]) / 1e3;
    }
    if (r.n == 0) return kNaN;
    return (double)(r.timestamps[r.n - 1] - r.prevTimestamp) / 1e3;
}

double r_tmin(const Rfa& r) {  // :1603
    if (r.n == 0) return kNaN;
    double m = r.values[0];
    int64_t t = r.timestamps[0];
    for (size_t i = 0; i < r.n; i++)
        if (r.values[i] <= m) {
            m = r.values[i];
            t = r.timestamps[i];
        }
    return (double)t / 1e3;
}
double r_tmax(const Rfa& r) {  // :1623
    if (r.n == 0) return kNaN;
    double m = r.values[0];
    int64_t t = r.timestamps[0];
    for (size_t i = 0; i < r.n; i++)
        if (r.values[i] >= m) {
            m = r.values[i];
            t = r.timestamps[i];
        }
    return (double)t / 1e3;
}
double r_tlast_change(const Rfa& r) {  // :1669
    if (r.n == 0) return kNaN;
    double last = r.values[r.n - 1];
    for (ptrdiff_t i = (ptrdiff_t)r.n - 2; i >= 0; i--)
        if (r.values[i] != last) return (double)r.timestamps[i + 1] / 1e3;
    if (isnan(r.prevValue) || r.prevValue != last) return (double)r.timestamps[0] / 1e3;
    return kNaN;
}

double r_mad(const double* values, size_t n) {  // mad :1476
    double median = quantile(0.5, values, n);
    std::vector<double> ds(n);
    for (size_t i = 0; i < n; i++) ds[i] = fabs(values[i] - median);
    return quantile(0.5, ds.data(), n);
}

double r_outlier_iqr(const Rfa& r) {  // :1427
    if (r.n < 2) return kNaN;
    double q25 = quantile(0.25, r.values, r.n);
    double q75 = quantile(0.75, r.values, r.n);
    double iqr = 1.5 * (q75 - q25);
    double v = r.values[r.n - 1];
    if (v > q75 + iqr || v < q25 - iqr) return v;
    return kNaN;
}

double r_zscore(const Rfa& r) {  // :2361
    double si = r_scrape_interval(r);
    double lag = r_lag(r);
    if (isnan(si) || isnan(lag) || lag > si) return kNaN;
    double d = r_last(r) - r_avg(r);
    if (d == 0) return 0;
    return d / sqrt(stdvar(r.values, r.n));
}

double r_ascent_descent(const Rfa& r, bool ascent) {  // :2315 / :2338
    const double* values = r.values;
    size_t n = r.n;
    double prev = r.prevValue;
    if (isnan(prev)) {
        if (n == 0) return kNaN;
        prev = values[0];
        values++;
        n--;
    }
    double s = 0;
    for (size_t i = 0; i < n; i++) {
        double v = values[i];
        double d = ascent ? (v - prev) : (prev - v);
        if (d > 0) s += d;
        prev = v;
    }
    return s;
}

double r_distinct(const Rfa& r) {  // :2403 (Go map[float64]: NaN keys never collide, +0 == -0)
    if (r.n == 0) return kNaN;
    std::vector<double> a;
    size_t nans = 0;
    for (size_t i = 0; i < r.n; i++) {
        if (isnan(r.values[i])) nans++;
        else a.push_back(r.values[i]);
    }
    std::sort(a.begin(), a.end());
    size_t d = 0;
    for (size_t i = 0; i < a.size(); i++)
        if (i == 0 || a[i] != a[i - 1]) d++;
    return (double)(d + nans);
}

double r_geomean(const Rfa& r) {  // :1741
    if (r.n == 0) return kNaN;
    double p = 1.0;
    for (size_t i = 0; i < r.n; i++) p *= r.values[i];
    return pow(p, 1 / (double)r.n);
}

double r_holt_winters(const Rfa& r) {  // :1030
    const double* values = r.values;
    size_t n = r.n;
    if (n == 0) return kNaN;
    double sf = r.args[r.idx];
    if (sf < 0 || sf > 1) return kNaN;
    double tf = r.args2[r.idx];
    if (tf < 0 || tf > 1) return kNaN;
    double s0 = r.prevValue;
    if (isnan(s0)) {
        s0 = values[0];
        values++;
        n--;
        if (n == 0) return s0;
    }
    double b0 = values[0] - s0;
    for (size_t i = 0; i < n; i++) {
        double v = values[i];
        double s1 = sf * v + (1 - sf) * (s0 + b0);
        double b1 = tf * (s1 - s0) + (1 - tf) * b0;
        s0 = s1;
        b0 = b1;
    }
    return s0;
}

void hoeffding(const Rfa& r, double* bound, double* avg) {  // :1353
    if (r.n == 0) {
        *bound = kNaN;
        *avg = kNaN;
        return;
    }
    if (r.n == 1) {
        *bound = 0;
        *avg = r.values[0];
        return;
    }
    double vMax = r_max(r), vMin = r_min(r), vAvg = r_avg(r);
    double vRange = vMax - vMin;
    *avg = vAvg;
    if (vRange <= 0) {
        *bound = 0;
        return;
    }
    double phi = r.args[r.idx];
    if (phi >= 1) {
        *bound = INFINITY;
        return;
    }
    if (phi <= 0) {
        *bound = 0;
        return;
    }
    *bound = vRange * sqrt(log(1 / (1 - phi)) / (2 * (double)r.n));
}

double r_duration(const Rfa& r) {  // :1151
    if (r.n == 0) return kNaN;
    int64_t tPrev = r.timestamps[0];
    int64_t dSum = 0;
    int64_t dMax = (int64_t)(r.args[r.idx] * 1000);
    for (size_t i = 0; i < r.n; i++) {
        int64_t d = r.timestamps[i] - tPrev;
        if (d <= dMax) dSum += d;
        tPrev = r.timestamps[i];
    }
    return (double)dSum / 1000;
}

enum { F_LE, F_GT, F_EQ, F_NE };
double r_filter(const Rfa& r, int cmp, bool sum, bool share) {  // newRollupFilter :1321, newRollupAvgFilter :1275
    if (r.n == 0) return kNaN;
    double lim = r.args[r.idx];
    double acc = 0;
    int cnt = 0;
    for (size_t i = 0; i < r.n; i++) {
        double v = r.values[i];
        bool hit = cmp == F_LE ? v <= lim : cmp == F_GT ? v > lim : cmp == F_EQ ? v == lim : v != lim;
        if (hit) {
            acc += v;
            cnt++;
        }
    }
    if (sum) return acc;
    if (share) return (double)cnt / (double)r.n;
    return (double)cnt;
}

// rollup_candlestick helpers :2228-2282
size_t candlestick_len(const Rfa& r) {
    size_t n = r.n;
    while (n > 0 && r.timestamps[n - 1] >= r.currTimestamp) n--;
    return n;
}
double candlestick_first(const Rfa& r) {
    if (r.prevTimestamp + r.window >= r.currTimestamp) return r.prevValue;
    return kNaN;
}

double call_func(int f, const Rfa& r) {
    switch (f) {
        case VMO_RF_DEFAULT_ROLLUP:
        case VMO_RF_LAST: return r_last(r);
        case VMO_RF_RATE: return r_deriv_fast(r);
        case VMO_RF_DELTA: return r_delta(r);
        case VMO_RF_AVG: return r_avg(r);
        case VMO_RF_MIN: return r_min(r);
        case VMO_RF_MAX: return r_max(r);
        case VMO_RF_SUM: return r_sum(r);
        case VMO_RF_COUNT: return r.n == 0 ? kNaN : (double)r.n;
        case VMO_RF_QUANTILE: return quantile(r.args[r.idx], r.values, r.n);
        case VMO_RF_MEDIAN: return quantile(0.5, r.values, r.n);
        case VMO_RF_FIRST: return r.n == 0 ? kNaN : r.values[0];
        case VMO_RF_RANGE: return r_max(r) - r_min(r);
        case VMO_RF_SUM2: {
            if (r.n == 0) return kNaN;
            double s = 0;
            for (size_t i = 0; i < r.n; i++) s += r.values[i] * r.values[i];
            return s;
        }
        case VMO_RF_STDDEV: return sqrt(stdvar(r.values, r.n));
        case VMO_RF_STDVAR: return stdvar(r.values, r.n);
        case VMO_RF_IDERIV: return r_ideriv(r);
        case VMO_RF_IDELTA: return r_idelta(r);
        case VMO_RF_DERIV: {
            double v, k;
            linear_regression(r.values, r.timestamps, r.n, r.currTimestamp, &v, &k);
            return k;
        }
        case VMO_RF_INCREASE_PURE: return r_increase_pure(r);
        case VMO_RF_CHANGES: return r_changes(r);
        case VMO_RF_CHANGES_PROMETHEUS: return r_changes_prometheus(r);
        case VMO_RF_RESETS: return r_increases_or_resets(r, false);
        case VMO_RF_INCREASES: return r_increases_or_resets(r, true);
        case VMO_RF_INTEGRATE: return r_integrate(r);
        case VMO_RF_LAG: return r_lag(r);
        case VMO_RF_LIFETIME: return r_lifetime(r);
        case VMO_RF_SCRAPE_INTERVAL: return r_scrape_interval(r);
        case VMO_RF_TMIN: return r_tmin(r);
        case VMO_RF_TMAX: return r_tmax(r);
        case VMO_RF_TFIRST: return r.n == 0 ? kNaN : (double)r.timestamps[0] / 1e3;
        case VMO_RF_TLAST: return r.n == 0 ? kNaN : (double)r.timestamps[r.n - 1] / 1e3;
        case VMO_RF_TLAST_CHANGE: return r_tlast_change(r);
        case VMO_RF_MODE: {
            std::vector<double> a(r.values, r.values + r.n);
            return mode_no_nans(r.prevValue, a.data(), a.size());
        }
        case VMO_RF_MAD: return r_mad(r.values, r.n);
        case VMO_RF_OUTLIER_IQR: return r_outlier_iqr(r);
        case VMO_RF_ZSCORE: return r_zscore(r);
        case VMO_RF_ASCENT: return r_ascent_descent(r, true);
        case VMO_RF_DESCENT: return r_ascent_descent(r, false);
        case VMO_RF_DISTINCT: return r_distinct(r);
        case VMO_RF_GEOMEAN: return r_geomean(r);
        case VMO_RF_PREDICT_LINEAR: {
            double v, k;
            linear_regression(r.values, r.timestamps, r.n, r.currTimestamp, &v, &k);
            if (isnan(v)) return kNaN;
            return v + k * r.args[r.idx];
        }
        case VMO_RF_HOLT_WINTERS: return r_holt_winters(r);
        case VMO_RF_HOEFFDING_LOWER: {
            double b, a;
            hoeffding(r, &b, &a);
            return a - b;
        }
        case VMO_RF_HOEFFDING_UPPER: {
            double b, a;
            hoeffding(r, &b, &a);
            return a + b;
        }
        case VMO_RF_DURATION: return r_duration(r);
        case VMO_RF_COUNT_LE: return r_filter(r, F_LE, false, false);
        case VMO_RF_COUNT_GT: return r_filter(r, F_GT, false, false);
        case VMO_RF_COUNT_EQ: return r_filter(r, F_EQ, false, false);
        case VMO_RF_COUNT_NE: return r_filter(r, F_NE, false, false);
        case VMO_RF_SHARE_LE: return r_filter(r, F_LE, false, true);
        case VMO_RF_SHARE_GT: return r_filter(r, F_GT, false, true);
        case VMO_RF_SHARE_EQ: return r_filter(r, F_EQ, false, true);
        case VMO_RF_SUM_LE: return r_filter(r, F_LE, true, false);
        case VMO_RF_SUM_GT: return r_filter(r, F_GT, true, false);
        case VMO_RF_SUM_EQ: return r_filter(r, F_EQ, true, false);
        case VMO_RF_PRESENT: return r.n > 0 ? 1 : kNaN;
        case VMO_RF_ABSENT: return r.n == 0 ? 1 : kNaN;
        case VMO_RF_STALE_SAMPLES: {
            if (r.n == 0) return kNaN;
            int c = 0;
            for (size_t i = 0; i < r.n; i++)
                if (is_stale_nan(r.values[i])) c++;
            return (double)c;
        }
        case VMO_RF_RATE_OVER_SUM: {  // :1705
            if (r.n == 0) return kNaN;
            double sum = 0;
            for (size_t i = 0; i < r.n; i++) sum += r.values[i];
            return sum / ((double)r.window / 1e3);
        }
        case VMO_RF_DELTA_PROMETHEUS:  // :1903
            if (r.n < 2) return kNaN;
            return r.values[r.n - 1] - r.values[0];
        case VMO_RF_RATE_PROMETHEUS: {  // :1946
            if (r.n < 2) return kNaN;
            double delta = r.values[r.n - 1] - r.values[0];
            if (isnan(delta) || r.window == 0) return kNaN;
            return delta / ((double)r.window / 1e3);
        }
        case VMO_RF_OPEN: {
            double v = candlestick_first(r);
            if (!isnan(v)) return v;
            size_t n = candlestick_len(r);
            if (n == 0) return kNaN;
            return r.values[0];
        }
        case VMO_RF_CLOSE: {
            size_t n = candlestick_len(r);
            if (n == 0) return candlestick_first(r);
            return r.values[n - 1];
        }
        case VMO_RF_HIGH:
        case VMO_RF_LOW: {
            size_t n = candlestick_len(r);
            const double* values = r.values;
            double m = candlestick_first(r);
            if (isnan(m)) {
                if (n == 0) return kNaN;
                m = values[0];
                values++;
                n--;
            }
            for (size_t i = 0; i < n; i++) {
                if (f == VMO_RF_HIGH ? values[i] > m : values[i] < m) m = values[i];
            }
            return m;
        }
    }
    return kNaN;
}

// binarySearchInt64 rollup.go:857
size_t binary_search_int64(const int64_t* a, size_t n, int64_t v) {
    size_t i = 0, j = n;
    while (i < j) {
        size_t h = (i + j) >> 1;
        if (h < n && a[h] < v) i = h + 1;
        else j = h;
    }
    return i;
}

// seekFirstTimestampIdxAfter rollup.go:825
size_t seek_first_ts_idx_after(const int64_t* ts, size_t n, int64_t seek, size_t nHint) {
    if (n == 0 || ts[0] > seek) return 0;
    size_t startIdx = nHint >= 2 ? nHint - 2 : 0;
    if (startIdx >= n) startIdx = n - 1;
    size_t endIdx = std::min(nHint + 2, n);
    if (startIdx > 0 && ts[startIdx] <= seek) {
        ts += startIdx;
        n -= startIdx;
        endIdx -= startIdx;
    } else {
        startIdx = 0;
    }
    if (endIdx < n && ts[endIdx] > seek) n = endIdx;
    if (n < 16) {
        for (size_t i = 0; i < n; i++)
            if (ts[i] > seek) return startIdx + i;
        return startIdx + n;
    }
    return startIdx + binary_search_int64(ts, n, seek + 1);
}

}  // namespace

extern "C" {

double vmo_quantile(double phi, const double* values, size_t n) { return quantile(phi, values, n); }
double vmo_mode_no_nans(double prev, double* a, size_t n) { return mode_no_nans(prev, a, n); }
double vmo_linear_regression(const double* values, const int64_t* ts, size_t n, int64_t intercept, double* k) {
    double v;
    linear_regression(values, ts, n, intercept, &v, k);
    return v;
}

// direct call of one rollupFunc on a hand-built rollupFuncArg (what rollup_test.go:testRollupFunc does)
double vmo_rollup_func_call(int func_id, double prev_value, int64_t prev_ts, const double* values, const int64_t* ts,
                            size_t n, double real_prev, double real_next, int64_t curr_ts, size_t idx, int64_t window,
                            const double* args, const double* args2) {
    Rfa r;
    r.prevValue = prev_value;
    r.prevTimestamp = prev_ts;
    r.values = values;
    r.timestamps = ts;
    r.n = n;
    r.realPrevValue = real_prev;
    r.realNextValue = real_next;
    r.currTimestamp = curr_ts;
    r.idx = idx;
    r.window = window;
    r.args = args;
    r.args2 = args2;
    return call_func(func_id, r);
}

int64_t vmo_rollup_points(int64_t start, int64_t end, int64_t step) { return 1 + (end - start) / step; }  // eval.go:243

// getScrapeInterval rollup.go:871
int64_t vmo_get_scrape_interval(const int64_t* timestamps, size_t n, int64_t defaultInterval) {
    if (n < 2) return defaultInterval;
    int64_t tsPrev = timestamps[n - 1];
    size_t m = n - 1;
    const int64_t* t = timestamps;
    if (m > 20) {
        t = timestamps + (m - 20);
        m = 20;
    }
    double intervals[20];
    size_t k = 0;
    for (ptrdiff_t i = (ptrdiff_t)m - 1; i >= 0; i--) {
        intervals[k++] = (double)(tsPrev - t[i]);
        tsPrev = t[i];
    }
    double q = quantile(0.6, intervals, k);
    int64_t si = (int64_t)q;
    if (si <= 0) return defaultInterval;
    return si;
}

// getMaxPrevInterval rollup.go:899
int64_t vmo_get_max_prev_interval(int64_t si) {
    if (si <= 2 * 1000) return si + 4 * si;
    if (si <= 4 * 1000) return si + 2 * si;
    if (si <= 8 * 1000) return si + si;
    if (si <= 16 * 1000) return si + si / 2;
    if (si <= 32 * 1000) return si + si / 4;
    return si + si / 8;
}

// removeCounterResets rollup.go:921
void vmo_remove_counter_resets(double* values, const int64_t* timestamps, size_t n, int64_t maxStalenessInterval) {
    if (n == 0) return;
    double correction = 0;
    double prevValue = values[0];
    for (size_t i = 0; i < n; i++) {
        double v = values[i];
        double d = v - prevValue;
        if (d < 0) {
            if ((-d * 8) < prevValue) correction += prevValue - v;
            else correction += prevValue;
        }
        if (i > 0 && maxStalenessInterval > 0) {
            int64_t gap = timestamps[i] - timestamps[i - 1];
            if (gap > maxStalenessInterval) {
                correction = 0;
                prevValue = v;
                continue;
            }
        }
        prevValue = v;
        values[i] = v + correction;
        if (i > 0 && values[i] < values[i - 1]) values[i] = values[i - 1];
    }
}

// deltaValues rollup.go:960
void vmo_delta_values(double* values, size_t n) {
    if (n == 0) return;
    double prevDelta = 0;
    double prevValue = values[0];
    for (size_t i = 0; i + 1 < n; i++) {
        double v = values[i + 1];
        prevDelta = v - prevValue;
        values[i] = prevDelta;
        prevValue = v;
    }
    values[n - 1] = prevDelta;
}

// derivValues rollup.go:976
void vmo_deriv_values(double* values, const int64_t* timestamps, size_t n) {
    if (n == 0) return;
    double prevDeriv = 0;
    double prevValue = values[0];
    int64_t prevTs = timestamps[0];
    for (size_t i = 0; i + 1 < n; i++) {
        double v = values[i + 1];
        int64_t ts = timestamps[i + 1];
        if (ts == prevTs) {
            values[i] = prevDeriv;
            continue;
        }
        double dt = (double)(ts - prevTs) / 1e3;
        prevDeriv = (v - prevValue) / dt;
        values[i] = prevDeriv;
        prevValue = v;
        prevTs = ts;
    }
    values[n - 1] = prevDeriv;
}

// dropStaleNaNs eval.go:1985 (in-place compaction); returns the new length
size_t vmo_drop_stale_nans(double* values, int64_t* timestamps, size_t n) {
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        if (is_stale_nan(values[i])) continue;
        values[o] = values[i];
        timestamps[o] = timestamps[i];
        o++;
    }
    return o;
}

// rollupConfig.doInternal rollup.go:701
uint64_t vmo_rollup_do(const vmo_rollup_cfg* rc, double* out, const double* values, const int64_t* timestamps, size_t n) {
    int64_t maxPrevInterval = rc->step;
    if (rc->start < rc->end) {
        int64_t si = vmo_get_scrape_interval(timestamps, n, rc->step);
        maxPrevInterval = vmo_get_max_prev_interval(si);
    }
    if (rc->lookback_delta > 0 && maxPrevInterval > rc->lookback_delta) maxPrevInterval = rc->lookback_delta;
    if (rc->min_staleness_ms > 0 && maxPrevInterval < rc->min_staleness_ms) maxPrevInterval = rc->min_staleness_ms;
    int64_t window = rc->window;
    if (window <= 0) {
        window = rc->step;
        if (rc->may_adjust_window && window < maxPrevInterval) window = maxPrevInterval;
        if (rc->is_default_rollup && rc->lookback_delta > 0 && window > rc->lookback_delta) window = rc->lookback_delta;
    }
    Rfa rfa;
    memset(&rfa, 0, sizeof(rfa));
    rfa.window = window;
    rfa.args = rc->args;
    rfa.args2 = rc->args2;
    size_t i = 0, j = 0, ni = 0, nj = 0;
    uint64_t samplesScanned = n;
    uint64_t perCall = (uint64_t)rc->samples_scanned_per_call;
    int64_t points = vmo_rollup_points(rc->start, rc->end, rc->step);
    int64_t tEnd = rc->start;
    for (int64_t p = 0; p < points; p++, tEnd += rc->step) {
        int64_t tStart = tEnd - window;
        ni = seek_first_ts_idx_after(timestamps + i, n - i, tStart, ni);
        i += ni;
        if (j < i) j = i;
        nj = seek_first_ts_idx_after(timestamps + j, n - j, tEnd, nj);
        j += nj;

        rfa.prevValue = kNaN;
        rfa.prevTimestamp = tStart - maxPrevInterval;
        if (i < n && i > 0 && timestamps[i - 1] > rfa.prevTimestamp) {
            rfa.prevValue = values[i - 1];
            rfa.prevTimestamp = timestamps[i - 1];
        }
        rfa.values = values + i;
        rfa.timestamps = timestamps + i;
        rfa.n = j - i;
        rfa.realPrevValue = kNaN;
        if (i > 0) {
            double pv = values[i - 1];
            int64_t pt = timestamps[i - 1];
            int64_t curr = tStart;
            if (rfa.n > 0) curr = rfa.timestamps[0];
            if (rc->lookback_delta == 0 || (curr - pt) < rc->lookback_delta) rfa.realPrevValue = pv;
        }
        rfa.realNextValue = j < n ? values[j] : kNaN;
        rfa.currTimestamp = tEnd;
        rfa.idx = (size_t)p;
        out[p] = call_func(rc->func_id, rfa);
        if (perCall > 0) samplesScanned += perCall;
        else samplesScanned += rfa.n;
    }
    return samplesScanned;
}

// ---- aggr_incremental.go
void vmo_aggr_update(int aggr, double* dv, double* dc, const double* values, size_t p) {
    for (size_t i = 0; i < p; i++) {
        double v = values[i];
        switch (aggr) {
            case VMO_AGGR_SUM:  // :200
                if (isnan(v)) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                dv[i] += v;
                break;
            case VMO_AGGR_MIN:  // :242
                if (isnan(v)) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                if (v < dv[i]) dv[i] = v;
                break;
            case VMO_AGGR_MAX:  // :284
                if (isnan(v)) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                if (v > dv[i]) dv[i] = v;
                break;
            case VMO_AGGR_AVG:  // :325
                if (isnan(v)) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                dv[i] += v;
                dc[i]++;
                break;
            case VMO_AGGR_COUNT:
            case VMO_AGGR_GROUP:  // :381
                if (isnan(v)) break;
                dv[i]++;
                break;
            case VMO_AGGR_SUM2:  // :420
                if (isnan(v)) break;
                if (dc[i] == 0) { dv[i] = v * v; dc[i] = 1; break; }
                dv[i] += v * v;
                break;
            case VMO_AGGR_GEOMEAN:  // :459
                if (isnan(v)) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                dv[i] *= v;
                dc[i]++;
                break;
            case VMO_AGGR_ANY:  // :517
                if (dc[0] > 0 && i == 0) return;
                dc[i] = 1;
                dv[i] = v;
                break;
        }
    }
}

void vmo_aggr_merge(int aggr, double* dv, double* dc, const double* sv, const double* sc, size_t p) {
    if (aggr == VMO_AGGR_ANY) {  // :528
        if (dc[0] > 0) return;
        dc[0] = sc[0];
        for (size_t i = 0; i < p; i++) dv[i] = sv[i];
        return;
    }
    for (size_t i = 0; i < p; i++) {
        double v = sv[i];
        switch (aggr) {
            case VMO_AGGR_COUNT:
            case VMO_AGGR_GROUP:  // :392
                dv[i] += v;
                break;
            case VMO_AGGR_SUM:
            case VMO_AGGR_SUM2:  // :218, :438
                if (sc[i] == 0) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                dv[i] += v;
                break;
            case VMO_AGGR_MIN:  // :261
                if (sc[i] == 0) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                if (v < dv[i]) dv[i] = v;
                break;
            case VMO_AGGR_MAX:  // :303
                if (sc[i] == 0) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = 1; break; }
                if (v > dv[i]) dv[i] = v;
                break;
            case VMO_AGGR_AVG:  // :346
                if (sc[i] == 0) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = sc[i]; break; }
                dv[i] += v;
                dc[i] += sc[i];
                break;
            case VMO_AGGR_GEOMEAN:  // :479
                if (sc[i] == 0) break;
                if (dc[i] == 0) { dv[i] = v; dc[i] = sc[i]; break; }
                dv[i] *= v;
                dc[i] += sc[i];
                break;
        }
    }
}

void vmo_aggr_finalize(int aggr, double* dv, const double* dc, size_t p) {
    for (size_t i = 0; i < p; i++) {
        switch (aggr) {
            case VMO_AGGR_AVG:  // :368
                if (dc[i] == 0) dv[i] = kNaN;
                else dv[i] /= dc[i];
                break;
            case VMO_AGGR_COUNT:  // :400
                if (dv[i] == 0) dv[i] = kNaN;
                break;
            case VMO_AGGR_GROUP:  // :409
                if (dv[i] == 0) dv[i] = kNaN;
                else dv[i] = 1;
                break;
            case VMO_AGGR_GEOMEAN:  // :502
                if (dc[i] == 0) dv[i] = kNaN;
                else dv[i] = pow(dv[i], 1 / dc[i]);
                break;
            default:  // finalizeAggrCommon :189
                if (dc[i] == 0) dv[i] = kNaN;
                break;
        }
    }
}

}  // extern "C"
