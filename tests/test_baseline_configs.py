"""BASELINE.json `configs` as parity cases.
  configs[0]  (CPU only)  encode+decode round trip, 10k series x 1k int64 samples, nearest-delta2, bit-exact plumbing
  configs[1..4] (-m gpu)  the query shapes at reduced size against the oracle pipeline, plus size-independent properties
                          at a larger size (host path == device path bit for bit, sampled series == oracle)."""
import numpy as np

from conftest import SEED0
import pytest

import blockgen
from rollup_names import AGGR, RF

T0 = 1_700_000_000_000


def f64bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


# ------------------------------------------------------------------------------------------------ configs[0] (CPU)
def test_config0_encode_decode_roundtrip_10k_series(oracle):
    """v[i] = v[i-1] + 30000 + round(N(0,1000)); precisionBits 64 => MarshalTypeNearestDelta2 (plain,
    zstd does not reach the 0.9 ratio: SURVEY.md 7 table) ; product encoder bytes == oracle encoder bytes; round trip exact"""
    from victoriametrics_b200 import encoding
    rng = np.random.default_rng(SEED0 + 100)
    nser, n = 10_000, 1024
    inc = 30000 + np.round(rng.normal(0, 1000, (nser, n))).astype(np.int64)
    inc[:, 0] = rng.integers(0, 10 ** 9, nser)
    vals = np.cumsum(inc, axis=1)
    payload, offs, mts, firsts = encoding.marshal_columns(vals)
    assert set(mts.tolist()) <= {1, 5}
    assert (mts == 5).mean() > 0.9  # plain nearest-delta2, as the reference chooses for this data
    have_ref = bool(oracle.lib().vmo_zstd_ref_available())
    for s in list(range(0, nser, 97)) + [nser - 1]:
        b = payload[int(offs[s]):int(offs[s + 1])]
        assert int(firsts[s]) == int(vals[s, 0])
        rc, out = oracle.unmarshal_int64_array(b, int(mts[s]), int(firsts[s]), n)
        assert rc == 0 and np.array_equal(out, vals[s]), s
        if have_ref:
            ob, omt, ofirst = oracle.marshal_int64_array(vals[s])
            if omt == 5 and mts[s] == 5:
                assert np.array_equal(ob, b), s  # byte-identical marshaled form


# ------------------------------------------------------------------------------------------------ GPU configs
def _oracle_rollup_matrix(oracle, blocks, func, start, end, step, window, lookback=0, arg=None):
    import victoriametrics_b200 as vm
    rc = vm.promql.get_rollup_configs(func, start, end, step, window, lookback, args=arg)
    out = []
    for blk in blocks:
        r, ts, fv, _ = blk.oracle_unmarshal()
        assert r == 0
        ts, fv = ts.copy(), fv.copy()
        n = len(ts)
        if rc.dropStaleNaNs and n:
            n = oracle.lib().vmo_drop_stale_nans(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), n)
        ts, fv = ts[:n].copy(), fv[:n].copy()
        if rc.removeCounterResets and n:
            oracle.lib().vmo_remove_counter_resets(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), n,
                                                   lookback + window if lookback else 0)
        o, _ = oracle.rollup_do(RF[func], fv, ts, start, end, step, window, lookback_delta=lookback,
                                may_adjust_window=rc.MayAdjustWindow, is_default_rollup=rc.isDefaultRollup,
                                samples_scanned_per_call=rc.samplesScannedPerCall, args=arg)
        out.append(o)
    return np.stack(out)


def _mk(rng, n, kind, rows=2048, tkind="regular"):
    return [blockgen.OBlock(blockgen.gen_timestamps(rng, tkind, rows, T0), blockgen.gen_values(rng, kind, rows), -2, 64, i)
            for i in range(n)]


@pytest.mark.gpu
def test_config1_decode_plus_rate_5m_step15(oracle):
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 101)
    blocks = _mk(rng, 48, "counter_resets", rows=8192)
    descs, payload = blockgen.to_blockset(blocks)
    start, end, step, window = T0 + 300000, T0 + 15000 * 8191, 15000, 300000
    got, scanned = vm.promql.eval_rollup_func_host("rate", descs, payload, start, end, step, window)
    exp = _oracle_rollup_matrix(oracle, blocks, "rate", start, end, step, window)
    assert got.shape == (48, 8172)
    assert np.allclose(got, exp, rtol=1e-12, atol=0, equal_nan=True)
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    assert scanned == 48 * (8192 + 2 * 8172)  # len(values) + samplesScannedPerCall(rate)=2 per point (rollup.go:238)


@pytest.mark.gpu
@pytest.mark.parametrize("func,arg", [("avg_over_time", None), ("max_over_time", None), ("quantile_over_time", 0.99)])
def test_config2_gauge_over_time_functions(oracle, func, arg):
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 102)
    blocks = _mk(rng, 40, "gauge", rows=4096, tkind="jitter")
    assert all(b.vmt in (4, 6) for b in blocks)
    descs, payload = blockgen.to_blockset(blocks)
    B = vm.storage.Blocks(descs, payload)
    for step in (15000, 60000):
        start, end, window = T0 + 300000, T0 + 15000 * 4000, 300000
        got, _ = vm.promql.eval_rollup_func(func, B, start, end, step, window, args=arg)
        exp = _oracle_rollup_matrix(oracle, blocks, func, start, end, step, window, arg=arg)
        assert np.allclose(got, exp, rtol=1e-12, atol=0, equal_nan=True), (func, step)


@pytest.mark.gpu
def test_config3_sum_rate_by_label_single_rank(oracle):
    """sum(rate(m[5m])) by (mode): the per-GPU partial of config 4 (the cross-rank all-reduce is covered by
    tests/test_dist_aggr.py over gloo and by bench.py --gpus N over NCCL)"""
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 103)
    S, G = 64, 8
    blocks = _mk(rng, S, "counter", rows=2048)
    descs, payload = blockgen.to_blockset(blocks)
    B = vm.storage.Blocks(descs, payload)
    start, end, step, window = T0 + 300000, T0 + 15000 * 2000, 15000, 300000
    rc = vm.promql.get_rollup_configs("rate", start, end, step, window)
    groups = (np.arange(S) % G).astype(np.uint32)

    class Buf:
        def __init__(self, nbytes):
            self.t = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
            self.ptr = self.t.data_ptr()

    series, _ = vm.storage.decode_blocks(B)
    ia = vm.promql.IncrementalAggr("sum", G, rc.points, Buf)
    ia.update(series, rc, groups)
    got = ia.finalize(vm.default_context())
    rolled = _oracle_rollup_matrix(oracle, blocks, "rate", start, end, step, window)
    exp_v, exp_c = np.zeros((G, rc.points)), np.zeros((G, rc.points))
    for s in range(S):
        row = np.ascontiguousarray(rolled[s])
        g = int(groups[s])
        oracle.lib().vmo_aggr_update(AGGR["sum"], exp_v[g].ctypes.data_as(oracle.f64p), exp_c[g].ctypes.data_as(oracle.f64p),
                                     row.ctypes.data_as(oracle.f64p), rc.points)
    for g in range(G):
        oracle.lib().vmo_aggr_finalize(AGGR["sum"], exp_v[g].ctypes.data_as(oracle.f64p), exp_c[g].ctypes.data_as(oracle.f64p), rc.points)
    assert np.allclose(got, exp_v, rtol=1e-12, atol=0, equal_nan=True)
    # the one-call variant on compressed device blocks (vmb_eval_rollup_aggr_device) yields the same bits
    ctx = vm.default_context()
    ctx.set_fused(False)
    try:
        ia2 = vm.promql.IncrementalAggr("sum", G, rc.points, Buf)
        scanned = ia2.update_blocks(B, rc, groups)
        got2 = ia2.finalize(ctx)
    finally:
        ctx.set_fused(True)
    assert np.array_equal(got, got2, equal_nan=True)
    assert scanned > 0
    # ... and through the fused kernel, which folds every finished series into the group cells with atomic adds: the order of the
    # additions is not fixed (as in the reference, whose workers race for the series), the bits may differ in the last place
    ia4 = vm.promql.IncrementalAggr("sum", G, rc.points, Buf)
    scanned4 = ia4.update_blocks(B, rc, groups)
    got4 = ia4.finalize(ctx)
    assert scanned4 == scanned
    assert np.allclose(got4, got, rtol=1e-13, atol=0, equal_nan=True)
    # and the host path (chunked pipeline folding every chunk on the GPU): several chunks, every aggregate
    import os
    os.environ["VMB_PIPE_CHUNK_BLOCKS"] = "10"
    try:
        for aggr in ("sum", "min", "max", "avg", "count", "sum2", "geomean", "any", "group"):
            got3, scanned3 = vm.promql.eval_rollup_aggr_host(aggr, "rate", descs, payload, groups, G, start, end, step, window)
            e_v, e_c = np.zeros((G, rc.points)), np.zeros((G, rc.points))
            for s_ in range(S):
                row = np.ascontiguousarray(rolled[s_])
                g = int(groups[s_])
                oracle.lib().vmo_aggr_update(AGGR[aggr], e_v[g].ctypes.data_as(oracle.f64p), e_c[g].ctypes.data_as(oracle.f64p),
                                             row.ctypes.data_as(oracle.f64p), rc.points)
            for g in range(G):
                oracle.lib().vmo_aggr_finalize(AGGR[aggr], e_v[g].ctypes.data_as(oracle.f64p), e_c[g].ctypes.data_as(oracle.f64p), rc.points)
            assert np.allclose(got3, e_v, rtol=1e-12, atol=0, equal_nan=True), aggr
            assert scanned3 == scanned
    finally:
        del os.environ["VMB_PIPE_CHUNK_BLOCKS"]


@pytest.mark.gpu
@pytest.mark.parametrize("lookback,window_rows", [(0, 20), (45000, 20), (0, 3000)])
def test_streamed_counter_resets_bit_exact(oracle, lookback, window_rows):
    """the one-call paths run removeCounterResets inside the rollup kernel (rows corrected in shared memory as they stream
    through, state carried from fill to fill; window_rows=3000 forces the correct-in-global-memory fallback):
    last_over_time at step = scrape interval exposes every corrected value -> compare bit for bit"""
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 77 + window_rows)
    blocks = []
    for i in range(24):
        kind = ("counter_resets", "counter", "gauge", "counter_big")[i % 4]
        tkind = ("regular", "jitter", "irregular")[i % 3]
        blocks.append(blockgen.OBlock(blockgen.gen_timestamps(rng, tkind, 8192, T0), np.abs(blockgen.gen_values(rng, kind, 8192)),
                                      -2, 64, i))
    descs, payload = blockgen.to_blockset(blocks)
    start, end, step, window = T0 + 15000, T0 + 15000 * 8300, 15000, 15000 * window_rows
    rc = vm.promql.RollupConfig("last_over_time", start, end, step, window, LookbackDelta=lookback, removeCounterResets=True)
    exp = []
    for blk in blocks:
        r, ts, fv, _ = blk.oracle_unmarshal()
        assert r == 0
        ts, fv = ts.copy(), fv.copy()
        oracle.lib().vmo_remove_counter_resets(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), len(fv),
                                               lookback + window if lookback else 0)
        o, _ = oracle.rollup_do(RF["last_over_time"], fv, ts, start, end, step, window, lookback_delta=lookback)
        exp.append(o)
    exp = np.stack(exp)

    def same_bits(a, b):  # NaN (an empty window) matches any NaN; everything else bit for bit
        both_nan = np.isnan(a) & np.isnan(b)
        return bool(np.all(both_nan | (a.view(np.uint64) == b.view(np.uint64))))
    got_h, _ = vm.promql.eval_rollup_func_host("last_over_time", descs, payload, start, end, step, window, rc=rc)
    assert same_bits(got_h, exp)
    B = vm.storage.Blocks(descs, payload)
    out = torch.empty(exp.shape, dtype=torch.float64, device="cuda")
    vm.promql.eval_rollup_func("last_over_time", B, start, end, step, window, out_dev_ptr=out.data_ptr(), rc=rc)
    assert same_bits(out.cpu().numpy(), exp)
    # and the two-step API (stand-alone preamble kernel mutating the batch) agrees
    series, _ = vm.storage.decode_blocks(B)
    got2 = rc.do_series(series)[0]
    series.close()
    assert same_bits(got2, exp)


@pytest.mark.gpu
def test_regular_timestamp_mode_edges(oracle):
    """series with MarshalTypeDeltaConst timestamps take the rollup path that derives timestamps from the row index:
    tiny and huge scrape intervals, 2-row series, staleness markers (rows removed -> mode off), time-range trimming,
    query ranges that start before / end after the data, steps that do not divide the scrape interval"""
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 311)
    blocks = []
    for i, (n, dt) in enumerate([(2, 15000), (3, 1), (100, 7), (8192, 15000), (8192, 1000), (5000, 60000), (64, 999),
                                 (8192, 15000), (777, 15000), (16, 10**6), (8192, 1)]):
        ts = (T0 + dt * np.arange(n)).astype(np.int64)
        kind = "special" if i in (7, 8) else ("counter_resets" if i % 2 else "counter")
        vals = np.abs(blockgen.gen_values(rng, kind, n)) if kind != "special" else blockgen.gen_values(rng, kind, n)
        blocks.append(blockgen.OBlock(ts, vals, -2, 24 if i == 4 else 64, i))  # one block with lossy precisionBits
    assert all(b.tmt == 2 for b in blocks)
    descs, payload = blockgen.to_blockset(blocks)
    B = vm.storage.Blocks(descs, payload)
    for (start, end, step, window, tr) in [
            (T0 + 300000, T0 + 15000 * 8191, 15000, 300000, None),
            (T0 - 600000, T0 + 15000 * 9000, 15000, 60000, None),          # grid wider than the data on both sides
            (T0 + 1000, T0 + 200000, 7000, 21000, None),                    # step unrelated to any scrape interval
            (T0 + 300000, T0 + 15000 * 4000, 30000, 300000, (T0 + 15000 * 100 + 1, T0 + 15000 * 3000)),  # trimmed blocks
            (T0, T0 + 5000, 1, 5, None),                                    # millisecond grid over the 1 ms / 7 ms series
            (T0 + 6000, T0 + 8100, 1000, 5000, None)]:                      # 5000-row windows on the 1 ms series: more rows
                                                                            # than the shared-memory window holds
        kw = {} if tr is None else dict(tr_min=tr[0], tr_max=tr[1])
        exp = []
        for b in blocks:
            r, ts, fv, _ = b.oracle_unmarshal(*(tr or (-(1 << 63), (1 << 63) - 1)))
            assert r == 0
            ts, fv = ts.copy(), fv.copy()
            n = oracle.lib().vmo_drop_stale_nans(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), len(ts)) if len(ts) else 0
            ts, fv = ts[:n].copy(), fv[:n].copy()
            if n:
                oracle.lib().vmo_remove_counter_resets(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), n, 0)
            o, _ = oracle.rollup_do(RF["rate"], fv, ts, start, end, step, window, samples_scanned_per_call=2)
            exp.append(o)
        exp = np.stack(exp)
        out = torch.empty(exp.shape, dtype=torch.float64, device="cuda")
        vm.promql.eval_rollup_func("rate", B, start, end, step, window, out_dev_ptr=out.data_ptr(), **kw)
        got = out.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(exp)), (start, step)
        assert np.allclose(got, exp, rtol=1e-12, atol=0, equal_nan=True), (start, step)
        got_h, _ = vm.promql.eval_rollup_func_host("rate", descs, payload, start, end, step, window, **kw)
        assert np.allclose(got_h, exp, rtol=1e-12, atol=0, equal_nan=True), (start, step)


@pytest.mark.gpu
def test_random_query_grids_through_the_fast_paths(oracle):
    """random (start, end, step, window, lookbackDelta) grids over regular / jittered / irregular series: whichever rollup path
    the kernel picks per series (arithmetic timestamps, 32-bit relative timestamps, generic 64-bit, global-memory windows)
    must agree with the oracle"""
    import os
    import torch
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + int(os.environ.get("VMB_FUZZ_SEED", "5150")))
    blocks = []
    for i in range(30):
        n = int(rng.choice([2, 17, 300, 2049, 5000, 8192]))
        tkind = ("regular", "jitter", "irregular", "regular")[i % 4]
        ts = blockgen.gen_timestamps(rng, tkind, n, T0)
        if tkind == "regular" and i % 8 == 3:
            ts = (T0 + int(rng.choice([1, 250, 1000, 60000])) * np.arange(n)).astype(np.int64)
        blocks.append(blockgen.OBlock(ts, np.abs(blockgen.gen_values(rng, ("counter_resets", "counter", "gauge")[i % 3], n)), -2, 64, i))
    descs, payload = blockgen.to_blockset(blocks)
    B = vm.storage.Blocks(descs, payload)
    span = 15000 * 8192
    for trial in range(36):
        step = int(rng.choice([1000, 7000, 15000, 30000, 60000, 300000, 3600000]))
        window = int(rng.choice([0, step, 2 * step, 5 * step, 20 * step, 300000, 90001, 3600000, 15000 * 3000]))
        start = T0 + int(rng.integers(-2 * step, span // 2))
        npts = int(rng.integers(1, 1200))
        end = start + step * npts
        lookback = int(rng.choice([0, 0, 0, 300000, 45000]))
        func = ("rate", "rate", "increase", "avg_over_time", "max_over_time", "irate")[trial % 6]
        rc = vm.promql.get_rollup_configs(func, start, end, step, window, lookback)
        exp = []
        for b in blocks:
            r, ts, fv, _ = b.oracle_unmarshal()
            ts, fv = ts.copy(), fv.copy()
            if rc.removeCounterResets:
                oracle.lib().vmo_remove_counter_resets(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), len(fv),
                                                       lookback + window if lookback else 0)
            o, _ = oracle.rollup_do(RF[func], fv, ts, start, end, step, window, lookback_delta=lookback,
                                    may_adjust_window=rc.MayAdjustWindow, samples_scanned_per_call=rc.samplesScannedPerCall)
            exp.append(o)
        exp = np.stack(exp)
        out = torch.empty(exp.shape, dtype=torch.float64, device="cuda")
        vm.promql.eval_rollup_func(func, B, start, end, step, window, lookback, out_dev_ptr=out.data_ptr())
        got = out.cpu().numpy()
        key = (trial, func, start - T0, step, window, lookback, npts)
        assert np.array_equal(np.isnan(got), np.isnan(exp)), key
        assert np.allclose(got, exp, rtol=1e-12, atol=0, equal_nan=True), key


@pytest.mark.gpu
def test_config4_mixed_codec_increase_1h_step60(oracle):
    """40 % delta2 counters, 30 % gauges, 20 % const, 10 % delta-const -> increase(m[1h]) step 60 s"""
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 104)
    kinds = ["counter"] * 4 + ["counter_smooth"] * 2 + ["counter_big"] * 2 + ["gauge"] * 6 + ["const"] * 4 + ["delta_const"] * 2
    blocks = [blockgen.OBlock(blockgen.gen_timestamps(rng, "regular", 4096, T0), blockgen.gen_values(rng, k, 4096), -2, 64, i)
              for i, k in enumerate(kinds * 3)]
    assert {b.vmt for b in blocks} >= {1, 2, 3, 4, 5}
    descs, payload = blockgen.to_blockset(blocks)
    start, end, step, window = T0 + 3600000, T0 + 15000 * 4095, 60000, 3600000
    got, _ = vm.promql.eval_rollup_func_host("increase", descs, payload, start, end, step, window)
    exp = _oracle_rollup_matrix(oracle, blocks, "increase", start, end, step, window)
    assert np.allclose(got, exp, rtol=1e-12, atol=0, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("encoder", ["reference", "library"])
def test_full_block_size_properties_20k_blocks(oracle, encoder):
    """size-independent properties at a bench-like size (20 000 blocks x 8192 rows; blocks marshaled by the reference encoder
    -- oracle marshalInt64Array + the reference's libzstd 1.5.7 -- and by the library's own):
    (1) host pipeline == device path bit for bit; (2) rate >= 0 wherever defined (counter resets removed);
    (3) 1000 sampled series == oracle (CPU pipeline of bench.py's reference arm) within 1e-12; (4) samplesScanned closed form"""
    import torch
    import victoriametrics_b200 as vm
    import bench
    nb, rows = 20_000, 8192
    if encoder == "reference" and not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref/libzstd_ref.so not built")
    descs, payload, _ = bench.gen_blocks(nb, rows, seed=77, encoder=encoder)
    start, end, step = bench.query_range(rows, 300000, 15000)
    P = 1 + (end - start) // step
    B = vm.storage.Blocks(descs, payload)
    dev = torch.empty((nb, P), dtype=torch.float64, device="cuda")
    _, sc1 = vm.promql.eval_rollup_func("rate", B, start, end, step, 300000, out_dev_ptr=dev.data_ptr())
    torch.cuda.synchronize()
    host, sc2 = vm.promql.eval_rollup_func_host("rate", descs, payload, start, end, step, 300000)
    assert sc1 == sc2 == nb * (rows + 2 * P)
    d = dev.cpu().numpy()
    assert np.array_equal(f64bits(d), f64bits(host))
    assert not np.isnan(d).any() and (d >= 0).all() and np.isfinite(d).all()
    arm = bench.CpuArm(descs, payload, "rate", start, end, step, 300000)
    res = bench.parity_check(arm, lambda idx: d[idx], 1000)
    assert res["ok"] and res["series"] == 1000, res
