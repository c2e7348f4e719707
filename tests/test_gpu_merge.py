"""series assembly on the GPU: mergeSortBlocks netstorage.go:566 + DeduplicateSamples dedup.go:30.
Reference vectors (netstorage_test.go:11, dedup_test.go) through the CUDA path, then randomized differentials against the
oracle: overlapping / touching / replicated blocks in any arrival order, several scales, time-range trimming, dedup on/off."""
import numpy as np

from conftest import SEED0
import pytest

import blockgen
from conftest import gofloat

pytestmark = pytest.mark.gpu


@pytest.fixture()
def vmctx():
    import victoriametrics_b200 as vm
    ctx = vm.default_context()
    yield vm, ctx
    ctx.set_dedup_interval(0)


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


def _block_from_floats(vm, ts, vals, series_idx=0):
    """a marshaled block holding exactly these floats (decimal.AppendFloatToDecimal picks mantissas + scale)"""
    m, e = vm.decimal.append_float_to_decimal(np.asarray(vals, dtype=np.float64))
    b = blockgen.OBlock(np.asarray(ts, dtype=np.int64), np.asarray(m, dtype=np.int64), int(e), 64, series_idx)
    rc, _, fv, _ = b.oracle_unmarshal()
    assert rc == 0 and _same(fv, vals), (vals, fv)  # the KAT values survive the decimal round trip
    return b


def _decode(vm, blocks, tr_min=None, tr_max=None):
    descs, payload = blockgen.to_blockset(blocks)
    B = vm.storage.Blocks(descs, payload)
    kw = {}
    if tr_min is not None:
        kw = dict(tr_min=tr_min, tr_max=tr_max)
    series, status = vm.storage.decode_blocks(B, **kw)
    out = series.to_lists()
    series.close()
    assert not status.any()
    return out


def test_merge_sort_blocks_kats_on_gpu(vmctx, kats):
    vm, ctx = vmctx
    ran = 0
    for k in kats["merge_sort_blocks"]:
        if not k["blocks"] or any(len(b["timestamps"]) == 0 for b in k["blocks"]):
            continue  # a stored block always has rows (block_header.go:236)
        ctx.set_dedup_interval(k["dedup_interval"])
        blocks = [_block_from_floats(vm, b["timestamps"], [gofloat(x) for x in b["values"]]) for b in k["blocks"]]
        (ts, v), = _decode(vm, blocks)
        assert ts.tolist() == k["timestamps_expected"], k
        assert _same(v, [gofloat(x) for x in k["values_expected"]]), k
        ran += 1
    assert ran == 12


def test_deduplicate_samples_kats_on_gpu(vmctx, kats):
    vm, ctx = vmctx
    ran = 0
    for k in kats["dedup_samples"]:
        if not k["timestamps"]:
            continue
        ctx.set_dedup_interval(k["interval"])
        blocks = [_block_from_floats(vm, k["timestamps"], [gofloat(x) for x in k["values"]])]
        (ts, v), = _decode(vm, blocks)
        assert ts.tolist() == k["timestamps_expected"], k
        assert _same(v, [gofloat(x) for x in k["values_expected"]]), k
        ran += 1
    assert ran >= 22


def _random_series_blocks(rng, sidx, t0=1_700_000_000_000):
    """1..6 blocks of one series with a random mix of disjoint, touching, overlapping and replicated time ranges"""
    nb = int(rng.integers(1, 7))
    blocks = []
    cursor = t0
    for _ in range(nb):
        rows = int(rng.choice([1, 2, 5, 33, 64, 200, 1000]))
        mode = rng.integers(0, 5)
        if mode == 0 and blocks:      # exact replica of the previous block
            p = blocks[-1]
            blocks.append(blockgen.OBlock(p.ts.copy(), p.vals.copy(), p.scale, 64, sidx))
            continue
        if mode == 1 and blocks:      # same timestamps, different values
            p = blocks[-1]
            blocks.append(blockgen.OBlock(p.ts.copy(), rng.integers(0, 10**6, len(p.ts)).astype(np.int64), int(rng.choice([-2, 0, 1])), 64, sidx))
            continue
        if mode == 2:                 # overlaps what came before
            start = cursor - int(rng.integers(0, 40)) * 1000
        elif mode == 3:               # touches: starts exactly at the previous end
            start = cursor
        else:                         # disjoint
            start = cursor + int(rng.integers(1, 10)) * 1000
        step = int(rng.choice([250, 1000, 1000, 15000]))
        ts = start + np.cumsum(rng.integers(0, 2 * step + 1, rows)).astype(np.int64)
        vals = np.cumsum(rng.integers(0, 1000, rows)).astype(np.int64)
        blocks.append(blockgen.OBlock(ts, vals, int(rng.choice([-2, 0, 3])), 64, sidx))
        cursor = max(cursor, int(ts[-1]))
    order = rng.permutation(len(blocks))  # arrival order is arbitrary (netstorage.go:444)
    return [blocks[i] for i in order]


def _oracle_series(oracle, blocks, dedup, tr_min=-(1 << 63), tr_max=(1 << 63) - 1):
    tss, vs = [], []
    for b in blocks:
        rc, ts, fv, _ = b.oracle_unmarshal(tr_min, tr_max)
        assert rc == 0
        tss.append(ts)
        vs.append(fv)
    return oracle.merge_sort_blocks(tss, vs, dedup)


@pytest.mark.parametrize("dedup", [0, 1, 1000, 30000])
def test_merge_random_differential(vmctx, oracle, dedup):
    vm, ctx = vmctx
    rng = np.random.default_rng(SEED0 + 1000 + dedup)
    ctx.set_dedup_interval(dedup)
    per_series = [_random_series_blocks(rng, s) for s in range(120)]
    flat = [b for blocks in per_series for b in blocks]
    got = _decode(vm, flat)
    assert len(got) == len(per_series)
    for s, blocks in enumerate(per_series):
        ets, ev = _oracle_series(oracle, blocks, dedup)
        assert np.array_equal(got[s][0], ets), (s, dedup)
        assert _same(got[s][1], ev), (s, dedup)
    # and with a time range that cuts through the blocks
    lo, hi = 1_700_000_000_000 + 20_000, 1_700_000_000_000 + 600_000
    got = _decode(vm, flat, lo, hi)
    for s, blocks in enumerate(per_series):
        ets, ev = _oracle_series(oracle, blocks, dedup, lo, hi)
        assert np.array_equal(got[s][0], ets), (s, dedup, "trim")
        assert _same(got[s][1], ev)


def test_dedup_negative_timestamps_and_long_runs(vmctx, oracle):
    """negative timestamps take the sequential replay (Go's % truncates toward zero); long runs of equal timestamps
    cross the 32-row chunks of the parallel path"""
    vm, ctx = vmctx
    rng = np.random.default_rng(SEED0 + 77)
    cases = []
    for n, lo in ((500, -100_000), (3000, -5_000), (8192, 0), (8192, 10**12)):
        ts = lo + np.sort(rng.integers(0, n * 40, n)).astype(np.int64)
        ts[100:180] = ts[100]  # 80 equal timestamps
        ts = np.sort(ts)
        cases.append(blockgen.OBlock(ts, rng.integers(-1000, 1000, n).astype(np.int64), -1, 64, len(cases)))
    for dedup in (7, 100, 1000):
        ctx.set_dedup_interval(dedup)
        got = _decode(vm, cases)
        for s, b in enumerate(cases):
            ets, ev = _oracle_series(oracle, [b], dedup)
            assert np.array_equal(got[s][0], ets), (s, dedup)
            assert _same(got[s][1], ev), (s, dedup)


def test_rate_over_merged_series_all_entry_points(vmctx, oracle):
    """overlapping blocks through the one-call paths (device-resident and host pipeline): rate() of the merged series"""
    import torch
    from rollup_names import RF
    vm, ctx = vmctx
    rng = np.random.default_rng(SEED0 + 4242)
    per_series = [_random_series_blocks(rng, s) for s in range(40)]
    flat = [b for blocks in per_series for b in blocks]
    descs, payload = blockgen.to_blockset(flat)
    t0 = 1_700_000_000_000
    start, end, step, window = t0 + 60_000, t0 + 900_000, 15_000, 120_000
    for dedup in (0, 15000):
        ctx.set_dedup_interval(dedup)
        exp = []
        for blocks in per_series:
            ts, fv = _oracle_series(oracle, blocks, dedup)
            ts, fv = ts.copy(), fv.copy()
            if len(fv):
                oracle.lib().vmo_remove_counter_resets(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), len(fv), 0)
            o, _ = oracle.rollup_do(RF["rate"], fv, ts, start, end, step, window)
            exp.append(o)
        exp = np.stack(exp)
        got_h, _ = vm.promql.eval_rollup_func_host("rate", descs, payload, start, end, step, window)
        assert np.allclose(got_h, exp, rtol=1e-12, atol=0, equal_nan=True), dedup
        B = vm.storage.Blocks(descs, payload)
        out = torch.empty(exp.shape, dtype=torch.float64, device="cuda")
        vm.promql.eval_rollup_func("rate", B, start, end, step, window, out_dev_ptr=out.data_ptr())
        assert np.allclose(out.cpu().numpy(), exp, rtol=1e-12, atol=0, equal_nan=True), dedup
