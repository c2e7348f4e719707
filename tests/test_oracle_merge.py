"""oracle pinning for the series-assembly step: the reference's own vectors for mergeSortBlocks
(app/vmselect/netstorage/netstorage_test.go:11) and DeduplicateSamples / needsDedup (lib/storage/dedup_test.go)"""
import numpy as np

from conftest import gofloat


def _same(a, b):  # equalWithNans dedup_test.go:43 (+ plain NaN == NaN)
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_needs_dedup_kats(kats, oracle):
    for k in kats["needs_dedup"]:
        ts = np.array(k["timestamps"], dtype=np.int64)
        got = oracle.lib().vmo_needs_dedup(ts.ctypes.data_as(oracle.i64p), len(ts), k["interval"])
        assert bool(got) == k["expected"], k


def test_deduplicate_samples_kats(kats, oracle):
    assert len(kats["dedup_samples"]) >= 24
    for k in kats["dedup_samples"]:
        vals = [gofloat(x) for x in k["values"]]
        ts, v = oracle.deduplicate_samples(k["timestamps"], vals, k["interval"])
        assert ts.tolist() == k["timestamps_expected"], k
        assert _same(v, [gofloat(x) for x in k["values_expected"]]), k
        ts2, v2 = oracle.deduplicate_samples(ts, v, k["interval"])  # idempotent (dedup_test.go:159)
        assert ts2.tolist() == ts.tolist() and _same(v2, v)


def test_merge_sort_blocks_kats(kats, oracle):
    assert len(kats["merge_sort_blocks"]) == 14
    for k in kats["merge_sort_blocks"]:
        ts, v = oracle.merge_sort_blocks([b["timestamps"] for b in k["blocks"]],
                                         [[gofloat(x) for x in b["values"]] for b in k["blocks"]], k["dedup_interval"])
        assert ts.tolist() == k["timestamps_expected"], k
        assert _same(v, [gofloat(x) for x in k["values_expected"]]), k


def test_merge_properties(oracle):
    """random overlapping blocks: the merge is a permutation of the input sorted by timestamp; with non-overlapping
    blocks in any order it is their concatenation in time order"""
    rng = np.random.default_rng(5)
    for _ in range(50):
        nb = int(rng.integers(1, 7))
        tss = [np.sort(rng.integers(0, 200, int(rng.integers(0, 40)))).astype(np.int64) for _ in range(nb)]
        vs = [rng.normal(size=len(t)) for t in tss]
        ts, v = oracle.merge_sort_blocks(tss, vs, 0)
        assert np.all(np.diff(ts) >= 0) and len(ts) == sum(len(t) for t in tss)
        allp = sorted(zip(np.concatenate(tss).tolist(), np.concatenate(vs).tolist())) if len(ts) else []
        assert sorted(zip(ts.tolist(), v.tolist())) == allp
    parts = [np.arange(100, 150), np.arange(0, 50), np.arange(50, 100)]
    ts, v = oracle.merge_sort_blocks(parts, [p * 1.5 for p in parts], 0)
    assert ts.tolist() == list(range(150)) and np.array_equal(v, np.arange(150) * 1.5)
