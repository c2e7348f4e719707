"""topk / bottomk over the rolled matrix (aggr.go:646 newAggrFuncTopK): the reference's own query vectors
(exec_test.go:6592-6900) and randomized differentials against a direct restatement of the per-point sort; the multi-process
protocol (candidate lists per shard -> merge -> apply per shard) is run with two shards on one GPU."""
import numpy as np

from conftest import SEED0
import pytest

pytestmark = pytest.mark.gpu
NAN = float("nan")


class Buf:
    def __init__(self, nbytes):
        import torch
        self.t = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device="cuda")
        self.ptr = self.t.data_ptr()


def _topk_ref(vals, ks, groups, reverse):
    """newAggrFuncTopK per point: sort the group's series (NaN first), blank all but the last k; ties do not occur here"""
    out = vals.copy()
    S, P = vals.shape
    ks = np.broadcast_to(np.asarray(ks, dtype=np.float64), (P,))
    for g in np.unique(groups):
        rows = np.nonzero(groups == g)[0]
        for p in range(P):
            k = ks[p]
            kn = 0 if (np.isnan(k) or k < 0) else int(min(k, len(rows)))
            col = vals[rows, p]
            key = np.where(np.isnan(col), -np.inf if not reverse else np.inf, col)
            order = np.argsort(-key if reverse else key, kind="stable")  # ascending "less" order
            out[rows[order[:len(rows) - kn]], p] = np.nan
    keep = ~np.all(np.isnan(out), axis=1)
    return out, keep


def _run(vm, vals, ks, groups=None, ngroups=1, reverse=False):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(vals)).cuda()
    keep = vm.promql.topk(ks, t.data_ptr(), vals.shape[0], vals.shape[1], Buf, group_ids=groups, ngroups=ngroups, reverse=reverse)
    return t.cpu().numpy(), keep


def _same(a, b):
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_topk_reference_query_vectors():
    """exec_test.go: topk(k, label_set(10,"foo","bar") or label_set(time()/150,"baz","sss")) on 1000..2000 s step 200 s"""
    import victoriametrics_b200 as vm
    t = np.arange(1000, 2001, 200, dtype=np.float64)
    base = np.stack([np.full(6, 10.0), t / 150])
    got, keep = _run(vm, base, 1)                                   # exec_test.go:6598 topk(1)
    assert _same(got, np.array([[10, 10, 10, NAN, NAN, NAN], [NAN, NAN, NAN, 10.666666666666666, 12, 13.333333333333334]]))
    assert keep.tolist() == [True, True]
    got, keep = _run(vm, base, -1)                                  # :6592 topk(-1) -> nothing
    assert not keep.any() and np.isnan(got).all()
    got, keep = _run(vm, base, NAN)                                 # :6889 topk(NaN) -> nothing
    assert not keep.any()
    for k in (2, 100500):                                           # :6865, :6895
        got, keep = _run(vm, base, k)
        assert _same(got, base) and keep.all()
    nan_series = np.stack([np.full(6, NAN), t / 150])               # :6850 topk(1, nan_timeseries)
    got, keep = _run(vm, nan_series, 1)
    assert keep.tolist() == [False, True] and _same(got[1], t / 150)


@pytest.mark.parametrize("reverse", [False, True])
def test_topk_random_differential(reverse):
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 17 + reverse)
    S, P, G = 300, 257, 7
    vals = rng.normal(size=(S, P)) * 100
    vals[rng.random((S, P)) < 0.15] = NAN
    vals[5] = NAN
    vals[:, 13] = NAN
    groups = rng.integers(0, G, S).astype(np.uint32)
    ks = rng.choice([0, 1, 2, 3, 5, 8, 64, 1000, -1, NAN, 2.9], P).astype(np.float64)
    ks = np.minimum(np.where(np.isnan(ks), ks, ks), 64 + 0 * ks)  # k <= 64 on the GPU
    exp, ekeep = _topk_ref(vals, ks, groups, reverse)
    got, keep = _run(vm, vals, ks, groups, G, reverse)
    assert _same(got, exp)
    assert keep.tolist() == ekeep.tolist()
    # scalar k, one group
    exp, ekeep = _topk_ref(vals, 4, np.zeros(S, dtype=np.uint32), reverse)
    got, keep = _run(vm, vals, 4, None, 1, reverse)
    assert _same(got, exp) and keep.tolist() == ekeep.tolist()


def test_topk_two_shards_protocol():
    """series split over two 'ranks': candidates per shard, merge of the gathered lists, apply per shard == single shot"""
    import torch
    import ctypes as C
    import victoriametrics_b200 as vm
    from victoriametrics_b200 import _lib
    rng = np.random.default_rng(SEED0 + 99)
    S, P, G, K = 200, 100, 3, 5
    vals = rng.normal(size=(S, P))
    vals[rng.random((S, P)) < 0.1] = NAN
    groups = rng.integers(0, G, S).astype(np.uint32)
    exp, ekeep = _topk_ref(vals, K, groups, False)
    ctx = vm.default_context()
    shards = [np.arange(0, S, 2), np.arange(1, S, 2)]
    gsz = np.bincount(groups, minlength=G).astype(np.uint32)
    dev, cands = [], []
    for rows in shards:
        t = torch.from_numpy(np.ascontiguousarray(vals[rows])).cuda()
        c = torch.empty(G * P * K * 2, dtype=torch.float64, device="cuda")  # {value, global series id} per entry
        g = np.ascontiguousarray(groups[rows])
        _lib.check(_lib.lib().vmb_topk_candidates(ctx.h, C.c_void_p(t.data_ptr()), len(rows), P, g.ctypes.data_as(_lib.u32p), G, K, 0,
                                                  len(dev) * S, C.c_void_p(c.data_ptr())))
        dev.append((t, g))
        cands.append(c)
    gathered = torch.cat(cands)  # what an all-gather delivers on every rank
    merged = torch.empty(G * P * K * 2, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().vmb_topk_merge(ctx.h, C.c_void_p(gathered.data_ptr()), 2, G * P, K, 0, C.c_void_p(merged.data_ptr())))
    ks = np.full(P, float(K))
    got = np.empty_like(vals)
    keep = np.zeros(S, dtype=bool)
    for r, (rows, (t, g)) in enumerate(zip(shards, dev)):
        flags = np.zeros(len(rows), dtype=np.uint8)
        _lib.check(_lib.lib().vmb_topk_apply(ctx.h, C.c_void_p(t.data_ptr()), len(rows), P, g.ctypes.data_as(_lib.u32p), G,
                                             gsz.ctypes.data_as(_lib.u32p), C.c_void_p(merged.data_ptr()), K,
                                             ks.ctypes.data_as(_lib.f64p), 0, r * S, flags.ctypes.data_as(_lib.u8p)))
        got[rows] = t.cpu().numpy()
        keep[rows] = flags.astype(bool)
    assert _same(got, exp) and keep.tolist() == ekeep.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("reverse", [False, True])
def test_topk_ties_keep_exactly_k(reverse):
    """topk(5, up)-like input: many equal values.  The reference keeps exactly k series per (group, point) (aggr.go:646 +
    fillNaNsAtIdx :786), an arbitrary subset of the tied ones (unstable sort.Slice); here the k lowest series ids of the tie."""
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 5150)
    S, P, G, K = 120, 40, 3, 5
    vals = rng.integers(0, 3, (S, P)).astype(np.float64)  # heavy ties
    vals[:, 0] = 1.0                                       # one point where every series ties
    vals[rng.random((S, P)) < 0.05] = NAN
    groups = (np.arange(S) % G).astype(np.uint32)
    got, keep = _run(vm, vals, K, groups, G, reverse)
    for g in range(G):
        rows = np.nonzero(groups == g)[0]
        for p in range(P):
            col = vals[rows, p]
            alive = ~np.isnan(got[rows, p])
            nvalid = int((~np.isnan(col)).sum())
            assert alive.sum() == min(K, nvalid), (g, p)
            # the survivors are the k best under (value, ascending series id)
            order = sorted((i for i in range(len(rows)) if not np.isnan(col[i])), key=lambda i: ((col[i] if reverse else -col[i]), rows[i]))
            assert sorted(np.nonzero(alive)[0].tolist()) == sorted(order[:K]), (g, p)
            assert np.array_equal(got[rows, p][alive], col[alive])
    assert keep.tolist() == (~np.all(np.isnan(got), axis=1)).tolist()
