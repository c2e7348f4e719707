"""Post-rollup operations on device matrices (csrc/matrix_ops.inc): binary operators (binary_op.go:155-203 with the element functions
of metricsql/binaryop), quantile / median across series (aggr.go:1217-1240) and mergeSeries (rollup_result_cache.go:618), against
direct numpy restatements of the Go element functions."""
import math

import numpy as np
import pytest

from conftest import SEED0

pytestmark = pytest.mark.gpu
NAN = float("nan")


def _ref_binop(op, is_bool, a, b):
    isn = math.isnan
    cmp_ops = {"==": lambda: (isn(b) if isn(a) else a == b), "!=": lambda: ((not isn(b)) if isn(a) else (True if isn(b) else a != b)),
               ">": lambda: a > b, "<": lambda: a < b, ">=": lambda: a >= b, "<=": lambda: a <= b}
    if op in cmp_ops:
        c = cmp_ops[op]()
        if not is_bool:
            return a if c else NAN
        if isn(a):
            return NAN
        return 1.0 if c else 0.0
    if op == "+":
        return a + b
    if op == "-":
        return a - b
    if op == "*":
        return a * b
    if op == "/":
        return np.float64(a) / np.float64(b)
    if op == "%":
        return math.fmod(a, b) if not (isn(a) or isn(b) or math.isinf(a) or b == 0) else (NAN if (isn(a) or isn(b) or math.isinf(a) or b == 0) else 0)
    if op == "^":
        if isn(a):
            return NAN
        try:
            return math.pow(a, b)
        except (OverflowError, ValueError):
            return float(np.float64(a) ** np.float64(b))
    if op == "atan2":
        return math.atan2(a, b)
    if op == "default":
        return b if isn(a) else a
    if op == "if":
        return NAN if isn(b) else a
    if op == "ifnot":
        return a if isn(b) else NAN
    raise KeyError(op)


@pytest.mark.parametrize("op", ["+", "-", "*", "/", "%", "^", "atan2", "==", "!=", ">", "<", ">=", "<=", "default", "if", "ifnot"])
def test_binary_ops_elementwise(op):
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 61 + len(op))
    S, P = 37, 211
    special = np.array([NAN, 0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 2.5, 1e300, -3.0])
    left = rng.normal(size=(S, P)) * 10
    right = rng.normal(size=(S // 2 + 1, P)) * 3
    left[rng.random((S, P)) < 0.2] = rng.choice(special, int((rng.random((S, P)) < 0.2).sum()) or 1)[0]
    m = rng.random(left.shape) < 0.15
    left[m] = rng.choice(special, m.sum())
    m = rng.random(right.shape) < 0.15
    right[m] = rng.choice(special, m.sum())
    lrows = rng.integers(0, S, 50).astype(np.uint32)
    rrows = rng.integers(0, right.shape[0], 50).astype(np.uint32)
    L, R = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    for is_bool in (False, True):
        dst = torch.empty((50, P), dtype=torch.float64, device="cuda")
        vm.promql.binary_op(op, L.data_ptr(), R.data_ptr(), 50, P, dst.data_ptr(), lrows, rrows, is_bool=is_bool)
        got = dst.cpu().numpy()
        with np.errstate(all="ignore"):
            exp = np.array([[_ref_binop(op, is_bool, float(left[lrows[i], j]), float(right[rrows[i], j])) for j in range(P)] for i in range(50)])
        assert np.array_equal(np.isnan(got), np.isnan(exp)), (op, is_bool)
        assert np.allclose(got, exp, rtol=1e-12, atol=0, equal_nan=True), (op, is_bool)
    # vector op scalar (binary_op.go:218-227): the scalar is a one-row matrix
    sc = torch.full((1, P), 2.0, dtype=torch.float64, device="cuda")
    dst = torch.empty((S, P), dtype=torch.float64, device="cuda")
    vm.promql.binary_op(op, L.data_ptr(), sc.data_ptr(), S, P, dst.data_ptr(), None, np.zeros(S, dtype=np.uint32))
    with np.errstate(all="ignore"):
        exp = np.array([[_ref_binop(op, False, float(left[i, j]), 2.0) for j in range(P)] for i in range(S)])
    assert np.allclose(dst.cpu().numpy(), exp, rtol=1e-12, atol=0, equal_nan=True)


def test_quantile_and_median_across_series(oracle):
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 62)
    S, P, G = 300, 90, 4
    vals = rng.normal(size=(S, P))
    vals[rng.random((S, P)) < 0.2] = NAN
    vals[:, 5] = NAN
    groups = rng.integers(0, G, S).astype(np.uint32)
    phis = np.concatenate([np.linspace(0, 1, P - 3), [-0.5, 1.5, NAN]])
    V = torch.from_numpy(vals).cuda()
    out = torch.empty((G, P), dtype=torch.float64, device="cuda")
    vm.promql.aggr_quantile(phis, V.data_ptr(), S, P, out.data_ptr(), groups, G)
    got = out.cpu().numpy()
    for g in range(G):
        rows = np.nonzero(groups == g)[0]
        for p in range(P):
            col = np.ascontiguousarray(vals[rows, p])
            exp = oracle.lib().vmo_quantile(float(phis[p]), col.ctypes.data_as(oracle.f64p), len(col))
            assert (np.isnan(got[g, p]) and np.isnan(exp)) or got[g, p] == exp or abs(got[g, p] - exp) <= 1e-12 * abs(exp), (g, p, got[g, p], exp)
    med = torch.empty((1, P), dtype=torch.float64, device="cuda")
    vm.promql.aggr_quantile(0.5, V.data_ptr(), S, P, med.data_ptr())
    assert np.allclose(med.cpu().numpy()[0], np.nanmedian(np.where(np.isnan(vals).all(0), 0, vals), axis=0) * np.where(np.isnan(vals).all(0), NAN, 1),
                       rtol=1e-12, equal_nan=True)


def test_merge_series_concat():
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 63)
    pa, pb = 17, 40
    a, b = rng.normal(size=(6, pa)), rng.normal(size=(5, pb))
    A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    a_rows = np.array([2, -1, 0, 5, -1, 3, 1], dtype=np.int64)   # -1: the series is new in b
    b_rows = np.array([0, 1, 2, 3, 4, -1, -1], dtype=np.int64)   # -1: the series is only in the cached part (:693-709)
    dst = torch.empty((7, pa + pb), dtype=torch.float64, device="cuda")
    vm.promql.merge_series(A.data_ptr(), a_rows, pa, B.data_ptr(), b_rows, pb, dst.data_ptr())
    got = dst.cpu().numpy()
    for i in range(7):
        ea = a[a_rows[i]] if a_rows[i] >= 0 else np.full(pa, NAN)
        eb = b[b_rows[i]] if b_rows[i] >= 0 else np.full(pb, NAN)
        assert np.array_equal(np.concatenate([ea, eb]), got[i], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["and", "unless", "default"])
def test_set_operators_with_right_groups(op):
    """binaryOpAnd / binaryOpUnless / binaryOpDefault (binary_op.go:430,610,463): several right series under one tag-set key"""
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(4242 + len(op))
    P, G = 97, 23
    nl, nr = 60, 75
    left = rng.normal(size=(nl, P))
    right = rng.normal(size=(nr, P))
    left[rng.random(left.shape) < 0.3] = np.nan
    right[rng.random(right.shape) < 0.6] = np.nan
    lg = rng.integers(0, G, nl).astype(np.uint32)
    rg = np.concatenate([np.arange(G), rng.integers(0, G, nr - G)]).astype(np.uint32)  # every key present on the right
    rng.shuffle(rg)
    # the reference loops, restated on numpy
    want = left.copy()
    for i in range(nl):
        rows = [r for r in range(nr) if rg[r] == lg[i]]
        for j in range(P):
            has = [right[r, j] for r in rows if not np.isnan(right[r, j])]
            if op == "and" and not has:
                want[i, j] = np.nan
            if op == "unless" and has:
                want[i, j] = np.nan
            if op == "default" and np.isnan(want[i, j]) and has:
                want[i, j] = has[0]
    dl = torch.from_numpy(left).cuda()
    dr = torch.from_numpy(right).cuda()
    tmp = torch.empty((G, P), dtype=torch.float64, device="cuda")
    dst = torch.empty((nl, P), dtype=torch.float64, device="cuda")
    vm.promql.set_op(op, dl.data_ptr(), lg, nl, dr.data_ptr(), rg, nr, G, P, dst.data_ptr(), tmp.data_ptr())
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
