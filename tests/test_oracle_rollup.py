"""Pins the CPU oracle's rollup executor against the reference's own vectors (app/vmselect/promql/rollup_test.go,
aggr_incremental_test.go), extracted into tests/golden/go_kats.json."""
import math

import numpy as np
import pytest
from conftest import gofloat
from rollup_names import AGGR, GO_FUNC, REMOVE_COUNTER_RESETS, RF

NAN = float("nan")


def eq_rel(got, exp, rel=1e-13):
    """rollup_test.go:1509 testRowsEqual semantics"""
    if math.isnan(exp):
        return math.isnan(got)
    if math.isnan(got):
        return False
    if exp == got:
        return True
    if math.isinf(exp) or math.isinf(got):
        return False
    return abs(got - exp) / abs(exp) <= rel


def call_func(oracle, name, values, timestamps, arg=None, arg2=None, prev=NAN, prev_ts=0, real_prev=NAN, real_next=0.0,
              curr_ts=0, window=None):
    """rollup_test.go:225 testRollupFunc: zero-valued rfa except prevValue/realPrevValue = nan"""
    v = np.array(values, dtype=np.float64)
    t = np.array(timestamps, dtype=np.int64)
    if name in REMOVE_COUNTER_RESETS and len(v):
        oracle.lib().vmo_remove_counter_resets(v.ctypes.data_as(oracle.f64p), t.ctypes.data_as(oracle.i64p), len(v), 0)
    if window is None:
        window = int(t[-1] - t[0]) if len(t) else 0
    a = np.array([arg if arg is not None else 0.0], dtype=np.float64)
    a2 = np.array([arg2 if arg2 is not None else 0.0], dtype=np.float64)
    return oracle.lib().vmo_rollup_func_call(RF[name], prev, prev_ts, v.ctypes.data_as(oracle.f64p),
                                             t.ctypes.data_as(oracle.i64p), len(v), real_prev, real_next, curr_ts, 0,
                                             window, a.ctypes.data_as(oracle.f64p), a2.ctypes.data_as(oracle.f64p))


def test_rollup_func_success_kats(kats, oracle):
    tv = [gofloat(x) for x in kats["test_values"]]
    tt = kats["test_timestamps"]
    for name, exp in kats["rollup_func_success"]:
        got = call_func(oracle, name, tv, tt)
        e = gofloat(exp)
        if math.isnan(e):
            assert math.isnan(got), name
        else:
            assert abs(got - e) <= 1e-14, (name, got, e)  # rollup_test.go:253 eps


def test_rollup_func_one_arg_kats(kats, oracle):
    tv = [gofloat(x) for x in kats["test_values"]]
    tt = kats["test_timestamps"]
    for name, arg, exp in kats["rollup_func_one_arg"]:
        got = call_func(oracle, name, tv, tt, arg=gofloat(arg))
        e = gofloat(exp)
        if math.isnan(e):
            assert math.isnan(got), (name, arg)
        elif math.isinf(e):
            assert got == e, (name, arg)
        else:
            assert abs(got - e) <= 1e-14, (name, arg, got, e)


def test_holt_winters_kats(kats, oracle):
    tv = [gofloat(x) for x in kats["test_values"]]
    tt = kats["test_timestamps"]
    for sf, tf, exp in kats["rollup_holt_winters"]:
        got = call_func(oracle, "holt_winters", tv, tt, arg=gofloat(sf), arg2=gofloat(tf))
        e = gofloat(exp)
        assert (math.isnan(got) if math.isnan(e) else abs(got - e) <= 1e-14), (sf, tf, got, e)


def test_outlier_iqr_kats(kats, oracle):
    for values, exp in kats["rollup_outlier_iqr"]:
        v = [gofloat(x) for x in values]
        got = call_func(oracle, "outlier_iqr_over_time", v, list(range(len(v))))
        e = gofloat(exp)
        assert (math.isnan(got) if math.isnan(e) else got == e)


def test_rollup_delta_kats(kats, oracle):
    for prev, real_prev, real_next, values, exp in kats["rollup_delta"]:
        v = [gofloat(x) for x in values]
        got = call_func(oracle, "delta", v, [0] * len(v), prev=gofloat(prev), real_prev=gofloat(real_prev),
                        real_next=gofloat(real_next), window=0)
        e = gofloat(exp)
        assert (math.isnan(got) if math.isnan(e) else got == e), (prev, real_prev, real_next, values)


def test_deriv_fast_prometheus_kats(kats, oracle):
    for values, window, exp in kats["rollup_deriv_fast_prometheus"]:
        v = [gofloat(x) for x in values]
        got = call_func(oracle, "rate_prometheus", v, [0] * len(v), window=window, prev=0.0, real_prev=0.0)
        # NB: the reference test does NOT run removeCounterResets here (it calls rollupDerivFastPrometheus directly)
        e = gofloat(exp)
        assert (math.isnan(got) if math.isnan(e) else got == e), (values, window)


def test_linear_regression_kats(kats, oracle):
    for values, ts, ev, ek in kats["linear_regression"]:
        v = np.array([gofloat(x) for x in values])
        t = np.array(ts, dtype=np.int64)
        k = np.zeros(1)
        got_v = oracle.lib().vmo_linear_regression(v.ctypes.data_as(oracle.f64p), t.ctypes.data_as(oracle.i64p), len(v),
                                                   int(t[0]) + 100, k.ctypes.data_as(oracle.f64p))
        # aggr_incremental_test.go:171 compareValues semantics (a NaN expectation is not enforced there)
        for got, exp in ((got_v, gofloat(ev)), (k[0], gofloat(ek))):
            if math.isnan(got):
                assert math.isnan(exp)
            else:
                assert not (abs(got - exp) > 1e-14)
    # single constant value => (values[0], 0) per rollup.go:1103 areConstValues
    v = np.array([1.0])
    t = np.array([1], dtype=np.int64)
    k = np.ones(1)
    assert oracle.lib().vmo_linear_regression(v.ctypes.data_as(oracle.f64p), t.ctypes.data_as(oracle.i64p), 1, 101,
                                              k.ctypes.data_as(oracle.f64p)) == 1.0 and k[0] == 0


def test_ideriv_duplicate_timestamps(oracle):
    # rollup_test.go:46 TestRollupIderivDuplicateTimestamps
    f = lambda **kw: call_func(oracle, "ideriv", **kw)
    assert f(values=[1, 2, 3, 4, 5], timestamps=[100, 100, 200, 300, 300], prev=0.0) == 20
    assert f(values=[1, 2, 3, 4, 5], timestamps=[100, 100, 300, 300, 300], prev=0.0) == 15
    assert math.isnan(f(values=[], timestamps=[]))
    assert math.isnan(f(values=[15], timestamps=[100]))
    assert f(values=[15], timestamps=[100], prev=10.0, prev_ts=90) == 500
    assert f(values=[15], timestamps=[100], prev=10.0, prev_ts=100) == math.inf
    assert f(values=[15, 20], timestamps=[100, 100], prev=10.0, prev_ts=100) == math.inf


def test_remove_counter_resets_kats(kats, oracle):
    # rollup_test.go:119 TestRemoveCounterResets
    tv = [gofloat(x) for x in kats["test_values"]]
    tt = kats["test_timestamps"]

    def rcr(values, ts, stale):
        v = np.array(values, dtype=np.float64)
        t = np.array(ts, dtype=np.int64)
        oracle.lib().vmo_remove_counter_resets(v.ctypes.data_as(oracle.f64p), t.ctypes.data_as(oracle.i64p), len(v), stale)
        return v.tolist()

    assert rcr(tv, tt, 0) == [123, 157, 167, 188, 221, 255, 320, 332, 364, 396, 398, 398]
    assert rcr([-100, -200, -300, -400], [0, 1, 2, 3], 0) == [-100, -100, -100, -100]
    assert rcr([100, 95, 120, 119, 139, 50], [0, 1, 2, 3, 4, 5], 0) == [100, 100, 125, 125, 145, 195]
    assert rcr([10, 12, 14, 4, 6, 8, 6, 8, 4, 6], [10, 20, 30, 60, 70, 80, 90, 100, 120, 130], 10) == \
        [10, 12, 14, 4, 6, 8, 14, 16, 4, 6]
    assert rcr([10, 12, 2, 4], [10, 20, 30, 60], 10) == [10, 12, 14, 4]
    out = rcr([34.094223, 2.7518, 2.140669, 0.044878, 1.887095, 2.546569, 2.490149, 0.045, 0.035684, 0.062454, 0.058296],
              list(range(11)), 0)
    assert all(b >= a for a, b in zip(out, out[1:]))


def test_delta_deriv_values_kats(kats, oracle):
    # rollup_test.go:172 TestDeltaValues, :196 TestDerivValues
    tv = np.array([gofloat(x) for x in kats["test_values"]])
    tt = np.array(kats["test_timestamps"], dtype=np.int64)
    L = oracle.lib()
    v = tv.copy()
    L.vmo_delta_values(v.ctypes.data_as(oracle.f64p), len(v))
    assert v.tolist() == [-89, 10, -23, 33, -20, 65, -87, 32, -12, 2, 0, 0]
    v = tv.copy()
    L.vmo_remove_counter_resets(v.ctypes.data_as(oracle.f64p), tt.ctypes.data_as(oracle.i64p), len(v), 0)
    L.vmo_delta_values(v.ctypes.data_as(oracle.f64p), len(v))
    assert v.tolist() == [34, 10, 21, 33, 34, 65, 12, 32, 32, 2, 0, 0]
    v = tv.copy()
    L.vmo_deriv_values(v.ctypes.data_as(oracle.f64p), tt.ctypes.data_as(oracle.i64p), len(v))
    exp = [-8900, 1111.111111111111, -1916.6666666666665, 2538.4615384615386, -1818.1818181818182, 3611.1111111111113,
           -43500, 1882.3529411764705, -666.6666666666667, 400, 0, 0]
    assert all(eq_rel(a, b) for a, b in zip(v.tolist(), exp))
    v = np.array([1, 2, 3, 4, 5, 6, 7], dtype=np.float64)
    t = np.array([100, 100, 200, 200, 300, 400, 400], dtype=np.int64)
    L.vmo_deriv_values(v.ctypes.data_as(oracle.f64p), t.ctypes.data_as(oracle.i64p), len(v))
    assert v.tolist() == [0, 20, 20, 20, 10, 10, 10]


def test_rollup_do_kats(kats, oracle):
    """every rollupConfig.Do sub-test of rollup_test.go (expected vectors AND samplesScanned)"""
    assert len(kats["rollup_do"]) >= 50
    for t in kats["rollup_do"]:
        name = GO_FUNC[t["func"]]
        values = [gofloat(x) for x in t["values"]]
        out, scanned = oracle.rollup_do(RF[name], values, t["timestamps"], t["start"], t["end"], t["step"], t["window"],
                                        lookback_delta=t["lookback_delta"], may_adjust_window=t["may_adjust_window"])
        exp = [gofloat(x) for x in t["expected"]]
        assert len(out) == len(exp), t["test"]
        for g, e in zip(out.tolist(), exp):
            assert eq_rel(g, e), (t["test"], out.tolist(), exp)
        if t["samples_scanned"] is not None:
            assert scanned == t["samples_scanned"], (t["test"], scanned)


def test_rollup_big_number_of_values(oracle):
    # rollup_test.go:1484
    n = 10000
    values = np.arange(n, dtype=np.float64)
    ts = (np.arange(n) // 2).astype(np.int64)
    out, scanned = oracle.rollup_do(RF["default_rollup"], values, ts, 0, n, n // 5, n // 4)
    assert scanned == 22002
    exp = [1, 4001, 8001, 9999, NAN, NAN]
    assert all(eq_rel(g, e) for g, e in zip(out.tolist(), exp))


def test_incremental_aggr(oracle):
    # aggr_incremental_test.go:15 TestIncrementalAggr
    nan = NAN
    tss = [[1, nan, 2, nan], [3, nan, nan, 4], [nan, nan, 5, 6], [7, nan, 8, 9], [4, nan, nan, nan], [2, nan, 3, 2],
           [0, nan, 1, 1]]
    expected = {"sum": [17, nan, 19, 22], "min": [0, nan, 1, 1], "max": [7, nan, 8, 9],
                "avg": [2.8333333333333335, nan, 3.8, 4.4], "count": [6, nan, 5, 5], "sum2": [79, nan, 103, 138],
                "geomean": [0, nan, 2.9925557394776896, 3.365865436338599]}
    L = oracle.lib()
    for name, exp in expected.items():
        for nworkers in (1, 2, 3, 7):
            partial = [(np.zeros(4), np.zeros(4)) for _ in range(nworkers)]
            used = [False] * nworkers
            for i, ts in enumerate(tss):
                w = i % nworkers
                used[w] = True
                v = np.array(ts, dtype=np.float64)
                L.vmo_aggr_update(AGGR[name], partial[w][0].ctypes.data_as(oracle.f64p),
                                  partial[w][1].ctypes.data_as(oracle.f64p), v.ctypes.data_as(oracle.f64p), 4)
            gv, gc = None, None
            for w in range(nworkers):
                if not used[w]:
                    continue
                if gv is None:
                    gv, gc = partial[w]
                    continue
                L.vmo_aggr_merge(AGGR[name], gv.ctypes.data_as(oracle.f64p), gc.ctypes.data_as(oracle.f64p),
                                 partial[w][0].ctypes.data_as(oracle.f64p), partial[w][1].ctypes.data_as(oracle.f64p), 4)
            L.vmo_aggr_finalize(AGGR[name], gv.ctypes.data_as(oracle.f64p), gc.ctypes.data_as(oracle.f64p), 4)
            for g, e in zip(gv.tolist(), exp):
                assert eq_rel(g, e, 1e-12), (name, nworkers, gv.tolist())


def test_drop_stale_nans_and_scrape_interval(oracle):
    from conftest import STALE_NAN
    v = np.array([1, STALE_NAN, 3, STALE_NAN, 5], dtype=np.float64)
    t = np.array([10, 20, 30, 40, 50], dtype=np.int64)
    n = oracle.lib().vmo_drop_stale_nans(v.ctypes.data_as(oracle.f64p), t.ctypes.data_as(oracle.i64p), 5)
    assert n == 3 and v[:3].tolist() == [1, 3, 5] and t[:3].tolist() == [10, 30, 50]
    ts = np.arange(0, 15000 * 100, 15000, dtype=np.int64)
    assert oracle.lib().vmo_get_scrape_interval(ts.ctypes.data_as(oracle.i64p), len(ts), 1) == 15000
    assert oracle.lib().vmo_get_max_prev_interval(15000) == 22500
    assert oracle.lib().vmo_get_max_prev_interval(1000) == 5000
    assert oracle.lib().vmo_get_max_prev_interval(60000) == 67500
