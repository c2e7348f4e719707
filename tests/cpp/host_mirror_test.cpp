// C++ host mirror (include/vmb200.hpp) exercised with the reference's own literal vectors, written like the Go tests:
//   app/vmselect/promql/rollup_test.go:806 TestRollupNoWindowPartialPoints, :977 TestRollupFuncsNoWindow (first/count/delta),
//   lib/encoding/encoding_test.go:197 TestMarshalUnmarshalInt64ArrayGeneric, lib/decimal/decimal_test.go:136.
// Build: g++ -std=c++17 -Iinclude tests/cpp/host_mirror_test.cpp -Lvictoriametrics_b200 -lvmb200   (needs a B200 to run;
// `--compile-only` style checks run on CPU in tests/test_cpp_host_mirror.py)
#include <cmath>
#include <cstdio>
#include <cstring>

#include "vmb200.hpp"

static int failures = 0;
#define CHECK(cond)                                                            \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                        \
        }                                                                      \
    } while (0)

static const std::vector<double> testValues = {123, 34, 44, 21, 54, 34, 99, 12, 44, 32, 34, 34};       // rollup_test.go:13
static const std::vector<int64_t> testTimestamps = {5, 15, 24, 36, 49, 60, 78, 80, 97, 115, 120, 130};  // rollup_test.go:14

static bool rowsEqual(const std::vector<double>& got, const std::vector<double>& exp) {  // testRowsEqual rollup_test.go:1509
    if (got.size() != exp.size()) return false;
    for (size_t i = 0; i < got.size(); i++) {
        if (std::isnan(exp[i]) != std::isnan(got[i])) return false;
        if (std::isnan(exp[i])) continue;
        if (std::fabs(got[i] - exp[i]) / std::fabs(exp[i] == 0 ? 1 : exp[i]) > 1e-13) return false;
    }
    return true;
}

int main() {
    const double nan = NAN;
    vmb::Ctx ctx(0);
    using namespace vmb;
    {  // TestRollupNoWindowPartialPoints/beforeStart
        promql::rollupConfig rc;
        rc.Func = "first_over_time";
        rc.Start = 0; rc.End = 25; rc.Step = 5; rc.Window = 0;
        rc.Timestamps = promql::getTimestamps(rc.Start, rc.End, rc.Step);
        std::vector<double> values;
        uint64_t scanned = rc.Do(ctx, values, testValues, testTimestamps);
        CHECK(scanned == 15);
        CHECK(rowsEqual(values, {nan, 123, nan, 34, nan, 44}));
    }
    {  // TestRollupNoWindowPartialPoints/middle
        promql::rollupConfig rc;
        rc.Func = "first_over_time";
        rc.Start = -50; rc.End = 150; rc.Step = 50;
        std::vector<double> values;
        uint64_t scanned = rc.Do(ctx, values, testValues, testTimestamps);
        CHECK(scanned == 24);
        CHECK(rowsEqual(values, {nan, nan, 123, 34, 32}));
    }
    {  // TestRollupFuncsNoWindow/count + delta
        promql::rollupConfig rc;
        rc.Func = "count_over_time";
        rc.Start = 0; rc.End = 160; rc.Step = 40;
        std::vector<double> values;
        CHECK(rc.Do(ctx, values, testValues, testTimestamps) == 24);
        CHECK(rowsEqual(values, {nan, 4, 4, 3, 1}));
        rc.Func = "delta";
        values.clear();
        CHECK(rc.Do(ctx, values, testValues, testTimestamps) == 24);
        CHECK(rowsEqual(values, {nan, 21, -9, 22, 0}));
    }
    {  // rate through getRollupConfigs: removeCounterResets + MayAdjustWindow (rollup.go:374)
        promql::rollupConfig rc = promql::getRollupConfigs("rate", 0, 160, 40, 0, 0);
        CHECK(rc.removeCounterResets && rc.MayAdjustWindow && rc.samplesScannedPerCall == 2);
    }
    {  // TestMarshalUnmarshalInt64ArrayGeneric
        struct { std::vector<int64_t> va; encoding::MarshalType mt; } cases[] = {
            {{1, 20, 234}, encoding::MarshalTypeNearestDelta2}, {{1, 20, -2345, 678934, 342}, encoding::MarshalTypeNearestDelta},
            {{1}, encoding::MarshalTypeConst},                  {{1, 2}, encoding::MarshalTypeDeltaConst},
            {{-10, -1, 8, 17, 26}, encoding::MarshalTypeDeltaConst}, {{100, 100, 100, 100}, encoding::MarshalTypeConst}};
        for (auto& c : cases) {
            std::vector<uint8_t> b = {'f', 'o', 'o'};
            encoding::Marshaled m = encoding::MarshalValues(b, c.va, 64);
            CHECK(m.mt == c.mt);
            CHECK(m.firstValue == c.va[0]);
            CHECK(std::memcmp(b.data(), "foo", 3) == 0);  // append semantics
            std::vector<int64_t> out = {7};
            int rc = encoding::UnmarshalValues(ctx, out, b.data() + 3, b.size() - 3, m.mt, m.firstValue, (int)c.va.size());
            CHECK(rc == 0);
            CHECK(out.size() == c.va.size() + 1 && out[0] == 7);
            CHECK(std::equal(c.va.begin(), c.va.end(), out.begin() + 1));
        }
        std::vector<int64_t> out;
        uint8_t junk[2] = {1, 2};
        CHECK(encoding::UnmarshalValues(ctx, out, junk, 2, encoding::MarshalTypeNearestDelta, 0, 4) == VMB_ERR_SHORT_SRC);
        CHECK(out.empty());
    }
    {  // TestAppendDecimalToFloat (bit-exact)
        std::vector<double> f = {1, 2};
        decimal::AppendDecimalToFloat(ctx, f, {874957, 1130435}, -5);
        const double exp[4] = {1, 2, 8.74957, 1.130435e1};
        CHECK(f.size() == 4 && std::memcmp(f.data(), exp, sizeof(exp)) == 0);
        std::vector<int64_t> d;
        CHECK(decimal::AppendFloatToDecimal(d, {-24, 0, 4.123, 0.3}) == -3);
        CHECK((d == std::vector<int64_t>{-24000, 0, 4123, 300}));
    }
    {  // evalRollupFunc vs evalRollupFuncWithIncrementalAggregate (eval.go:1845 / :1804): sum(delta(m)) over two series made
       // of the reference's test vectors (values scaled / shifted), marshaled into blocks by MarshalTimestamps / MarshalValues
        std::vector<vmb_block_desc> descs;
        std::vector<uint8_t> payload;
        for (int s = 0; s < 2; s++) {
            std::vector<int64_t> va;
            for (double v : testValues) va.push_back((int64_t)v * (s + 1) + 7 * s);
            vmb_block_desc d;
            std::memset(&d, 0, sizeof(d));
            d.ts_off = payload.size();
            encoding::Marshaled mt = encoding::MarshalTimestamps(payload, testTimestamps, 64);
            d.ts_size = (uint32_t)(payload.size() - d.ts_off);
            d.val_off = payload.size();
            encoding::Marshaled mv = encoding::MarshalValues(payload, va, 64);
            d.val_size = (uint32_t)(payload.size() - d.val_off);
            d.first_value = mv.firstValue;
            d.min_ts = mt.firstValue;
            d.max_ts = testTimestamps.back();
            d.rows = (uint32_t)va.size();
            d.series_idx = (uint32_t)s;
            d.scale = 0;
            d.ts_mt = (uint8_t)mt.mt;
            d.val_mt = (uint8_t)mv.mt;
            d.precision_bits = 64;
            descs.push_back(d);
        }
        promql::rollupConfig rc = promql::getRollupConfigs("delta", 0, 160, 40, 0, 0);
        std::vector<double> rolled, summed;
        uint64_t sc1 = promql::evalRollupFunc(ctx, rc, descs, payload, INT64_MIN, INT64_MAX, rolled, 2);
        uint64_t sc2 = promql::evalRollupFuncWithIncrementalAggregate(ctx, "sum", rc, descs, payload, {0, 0}, 1, INT64_MIN, INT64_MAX,
                                                                      summed);
        CHECK(sc1 == sc2 && sc1 == 2 * (12 + 2 * 5));  // len(values) + samplesScannedPerCall("delta") = 2 per point (rollup.go:238)
        const size_t P = rc.Timestamps.size();
        CHECK(rolled.size() == 2 * P && summed.size() == P);
        CHECK(rowsEqual(std::vector<double>(rolled.begin(), rolled.begin() + P), {nan, 21, -9, 22, 0}));  // TestRollupFuncsNoWindow/delta
        for (size_t p = 0; p < P; p++) {
            double a = rolled[p], b = rolled[P + p];
            double exp = std::isnan(a) ? b : (std::isnan(b) ? a : a + b);  // updateAggrSum skips NaN
            CHECK(std::isnan(exp) ? std::isnan(summed[p]) : summed[p] == exp);
        }
    }
    std::printf(failures ? "host_mirror_test: %d FAILURES\n" : "host_mirror_test: OK\n", failures);
    return failures ? 1 : 0;
}
