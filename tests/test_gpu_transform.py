"""Transform functions on device matrices (csrc/transform.inc) against direct restatements of the Go loops of
app/vmselect/promql/transform.go (cited per function)."""
import math

import numpy as np
import pytest

from conftest import SEED0

pytestmark = pytest.mark.gpu
NAN = float("nan")


def _matrix(rng, rows, points, nan_frac=0.25, lead=True):
    m = rng.normal(scale=50.0, size=(rows, points))
    m[rng.random(m.shape) < nan_frac] = NAN
    if lead:
        for r in range(0, rows, 3):
            m[r, : int(rng.integers(0, min(points, 9)))] = NAN       # leading NaNs (skipLeadingNaNs)
        for r in range(1, rows, 4):
            m[r, points - int(rng.integers(0, min(points, 9))):] = NAN  # trailing NaNs
    m[rows // 2] = NAN                                                # a series without any value
    return m


def _run(name, m, *args):
    import torch
    import victoriametrics_b200 as vm
    d = torch.from_numpy(m.copy()).cuda()
    vm.promql.transform(name, d.data_ptr(), m.shape[0], m.shape[1], *args)
    torch.cuda.synchronize()
    return d.cpu().numpy()


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


# ---- row functions: sequential restatements
def _running(values, rf):  # newTransformFuncRunning transform.go:1308
    v = values.copy()
    i0 = 0
    while i0 < len(v) and math.isnan(v[i0]):
        i0 += 1
    if i0 == len(v):
        return v
    prev = v[i0]
    for i in range(i0 + 1, len(v)):
        if not math.isnan(v[i]):
            prev = rf(prev, v[i], i - i0)
        v[i] = prev
    return v


RF = {"sum": lambda a, b, idx: a + b, "max": lambda a, b, idx: a if a > b else b, "min": lambda a, b, idx: a if a < b else b,
      "avg": lambda a, b, idx: a + (b - a) / float(idx + 1)}  # :1175-1196


def _set_last(v):  # setLastValues :1650
    k = len(v) - 1
    while k >= 0 and math.isnan(v[k]):
        k -= 1
    if k < 0:
        return v
    return np.full_like(v, v[k])


def _row_ref(name, values):
    v = values.copy()
    if name.startswith("running_"):
        return _running(v, RF[name[8:]])
    if name in ("range_sum", "range_min", "range_max", "range_avg"):
        return _set_last(_running(v, RF[name[6:]]))
    if name == "range_last":
        return _set_last(v)
    if name == "range_first":  # :1620
        i0 = 0
        while i0 < len(v) and math.isnan(v[i0]):
            i0 += 1
        return v if i0 == len(v) else np.full_like(v, v[i0])
    if name == "keep_last_value":  # :1214
        last = v[0]
        for i in range(len(v)):
            if not math.isnan(v[i]):
                last = v[i]
            else:
                v[i] = last
        return v
    if name == "keep_next_value":  # :1237
        nxt = v[-1]
        for i in range(len(v) - 1, -1, -1):
            if not math.isnan(v[i]):
                nxt = v[i]
            else:
                v[i] = nxt
        return v
    if name == "remove_resets":  # removeCounterResetsMaybeNaNs :2906
        i0 = 0
        while i0 < len(v) and math.isnan(v[i0]):
            i0 += 1
        if i0 == len(v):
            return v
        corr, prev = 0.0, v[i0]
        for i in range(i0, len(v)):
            if math.isnan(v[i]):
                continue
            d = v[i] - prev
            if d < 0:
                corr += (prev - v[i]) if (-d * 8) < prev else prev
            prev = v[i]
            v[i] = v[i] + corr
        return v
    if name == "interpolate":  # :1261
        lo, hi = 0, len(v)
        while lo < hi and math.isnan(v[lo]):
            lo += 1
        while hi > lo and math.isnan(v[hi - 1]):
            hi -= 1
        w = v[lo:hi]
        i = 0
        while i < len(w):
            if not math.isnan(w[i]):
                i += 1
                continue
            prev = w[i - 1]
            j = i + 1
            while j < len(w) and math.isnan(w[j]):
                j += 1
            nxt = w[j]
            delta = (nxt - prev) / float(j - i + 1)
            while i < j:
                prev += delta
                w[i] = prev
                i += 1
        return v
    raise KeyError(name)


@pytest.mark.parametrize("name", ["running_sum", "running_min", "running_max", "running_avg", "range_sum", "range_min", "range_max",
                                  "range_avg", "range_first", "range_last", "keep_last_value", "keep_next_value", "remove_resets",
                                  "interpolate"])
@pytest.mark.parametrize("shape", [(70, 97), (33, 32), (5, 1), (1, 400)])
def test_row_functions_bit_exact(name, shape):
    rng = np.random.default_rng(SEED0 + 900 + len(name) * 7 + shape[1])
    m = _matrix(rng, *shape)
    if name == "remove_resets":  # counters with resets and gaps
        m = np.abs(np.cumsum(np.abs(m.copy() * 0 + rng.normal(scale=5.0, size=m.shape)), axis=1))
        for r in range(m.shape[0]):
            for _ in range(3):
                k = int(rng.integers(0, m.shape[1]))
                m[r, k:] -= m[r, k] * float(rng.choice([1.0, 0.05]))
        m[rng.random(m.shape) < 0.2] = NAN
    got = _run(name, m)
    want = np.stack([_row_ref(name, m[r]) for r in range(m.shape[0])])
    assert _same(got, want), name


# ---- element functions
def test_exact_element_functions():
    rng = np.random.default_rng(SEED0 + 950)
    m = _matrix(rng, 40, 61, lead=False)
    m[0, :5] = [0.0, -0.0, np.inf, -np.inf, 2.5]
    P = m.shape[1]
    assert _same(_run("abs", m), np.abs(m))
    assert _same(_run("ceil", m), np.ceil(m))
    assert _same(_run("floor", m), np.floor(m))
    with np.errstate(invalid="ignore"):
        assert _same(_run("sqrt", m), np.sqrt(m))
    assert _same(_run("deg", m), m * 180 / math.pi)
    assert _same(_run("rad", m), m * math.pi / 180)
    sg = np.where(m < 0, -1.0, np.where(m > 0, 1.0, 0.0))  # sgn(NaN) = 0, transform.go:2362
    assert np.array_equal(_run("sgn", m), sg)
    lo, hi = rng.normal(scale=10, size=P) - 20, rng.normal(scale=10, size=P) + 20
    assert _same(_run("clamp", m, lo, hi), np.where(m > hi, hi, np.where(m < lo, lo, m)))
    assert _same(_run("clamp_min", m, -3.0), np.where(m < -3.0, -3.0, m))
    assert _same(_run("clamp_max", m, hi), np.where(m > hi, hi, m))


@pytest.mark.parametrize("nearest", [1.0, 0.1, 0.25, 5.0, 100.0, 0.003])
def test_round_matches_go_formula(nearest):
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 960)
    m = rng.normal(scale=300.0, size=(20, 50))
    m[rng.random(m.shape) < 0.1] = NAN
    _, e = vm.decimal.append_float_to_decimal(np.array([nearest]))
    p10 = float("1e%d" % (-int(e)))
    want = np.empty_like(m)
    for idx, v in np.ndenumerate(m):
        if math.isnan(v):
            want[idx] = NAN
            continue
        x = v + 0.5 * math.copysign(nearest, v)
        x = x - math.fmod(x, nearest)
        x = float(math.trunc(x * p10))
        want[idx] = x / p10
    assert _same(_run("round", m, nearest), want)


@pytest.mark.parametrize("name,fn", [("exp", np.exp), ("ln", np.log), ("log2", np.log2), ("log10", np.log10), ("sin", np.sin), ("cos", np.cos),
                                     ("tan", np.tan), ("asin", np.arcsin), ("acos", np.arccos), ("atan", np.arctan), ("sinh", np.sinh),
                                     ("cosh", np.cosh), ("tanh", np.tanh), ("asinh", np.arcsinh), ("acosh", np.arccosh), ("atanh", np.arctanh)])
def test_library_math_functions_within_tolerance(name, fn):
    """these go through the CUDA math library (Go: pure-Go / assembly math); tolerance = north_star's 1e-9 relative, here 1e-12 + 1e-15 abs"""
    rng = np.random.default_rng(SEED0 + 970)
    m = rng.normal(scale=2.0, size=(16, 40))
    if name in ("asin", "acos", "atanh"):
        m = np.tanh(m)
    if name == "acosh":
        m = 1.0 + np.abs(m)
    if name in ("ln", "log2", "log10"):
        m = np.abs(m) + 1e-3
    m[rng.random(m.shape) < 0.1] = NAN
    with np.errstate(all="ignore"):
        want = fn(m)
    got = _run(name, m)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-15, equal_nan=True)
