"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI (victoriametrics_b200 ctypes mirror), against
the CPU oracle on the same seeded inputs.  Integer / byte work is bit-exact; float rollups use the tolerance stated in
BASELINE.json's north_star (1e-9 relative) or tighter where written."""
import base64
import ctypes as C
import json
import math
import os
import struct

import numpy as np

from conftest import SEED0
import pytest

import blockgen
from conftest import STALE_NAN, gofloat
from rollup_names import AGGR, GO_FUNC, RF, RF_IDS

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
I64_MIN, I64_MAX = -(1 << 63), (1 << 63) - 1


@pytest.fixture(scope="module")
def vm():
    import victoriametrics_b200 as v
    v.default_context()
    return v


def f64bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def close(got, exp, rel):
    if math.isnan(exp):
        return math.isnan(got)
    if math.isnan(got):
        return False
    if got == exp:
        return True
    if math.isinf(exp) or math.isinf(got):
        return False
    return abs(got - exp) <= rel * max(abs(exp), abs(got))


def assert_allclose_nan(got, exp, rel, what=""):
    got = np.asarray(got, dtype=np.float64)
    exp = np.asarray(exp, dtype=np.float64)
    assert got.shape == exp.shape, (what, got.shape, exp.shape)
    nan_g, nan_e = np.isnan(got), np.isnan(exp)
    assert np.array_equal(nan_g, nan_e), (what, "NaN pattern differs", np.argwhere(nan_g != nan_e)[:5])
    g, e = got[~nan_g], exp[~nan_e]
    inf = np.isinf(e) | np.isinf(g)
    assert np.array_equal(g[inf], e[inf]), (what, "inf mismatch")
    g, e = g[~inf], e[~inf]
    err = np.abs(g - e)
    tol = rel * np.maximum(np.abs(e), np.abs(g))
    bad = err > tol
    assert not bad.any(), (what, "max rel err", float((err[bad] / np.maximum(np.abs(e[bad]), 1e-300)).max()), g[bad][:3], e[bad][:3])


# ------------------------------------------------------------------------------------------------ codec: per-call
def test_unmarshal_values_per_call_all_types(vm, oracle):
    rng = np.random.default_rng(SEED0 + 1)
    seen = set()
    for kind in blockgen.VALUE_KINDS:
        for n in (1, 2, 3, 33, 512, 513, 8192):
            vals = blockgen.gen_values(rng, kind, n)
            b, mt, first = oracle.marshal_int64_array(vals)
            seen.add(mt)
            rc, exp = oracle.unmarshal_int64_array(b, mt, first, n)
            assert rc == 0
            got = vm.encoding.unmarshal_values(b, mt, first, n)
            assert np.array_equal(got, exp), (kind, n, mt)
    assert seen == {1, 2, 3, 4, 5, 6}


def test_unmarshal_nearest_delta_kats_on_gpu(vm, kats, oracle):
    for key, mt in (("marshal_nearest_delta", 6), ("marshal_nearest_delta2", 5)):
        for va, pb, first, hx in kats[key]:
            if len(va) < (2 if mt == 5 else 1):
                continue
            src = np.frombuffer(bytes.fromhex(hx), dtype=np.uint8)
            rc, exp = oracle.unmarshal_int64_array(src, mt, first, len(va))
            assert rc == 0
            got = vm.encoding.unmarshal_values(src, mt, first, len(va))
            assert np.array_equal(got, exp), (key, va, pb)


def test_unmarshal_errors_match_reference_error_sites(vm, oracle):
    def rc_of(src, mt, first, n):
        try:
            vm.encoding.unmarshal_values(np.array(src, dtype=np.uint8), mt, first, n)
            return 0
        except vm.VmbError as e:
            return e.code
    cases = [
        ([1, 2], 6, 0, 4),                       # too small len(src) int.go:183
        ([0x80], 6, 0, 2),                       # truncated varint
        ([0xFF] * 9 + [0x02], 6, 0, 2),          # 10th byte > 1
        ([0xFF] * 10 + [0x01], 6, 0, 2),         # 11-byte varint
        ([1, 2, 3, 0], 6, 0, 4),                 # trailing byte nearest_delta.go:65
        ([1, 2, 3, 0x80], 6, 0, 4),              # trailing partial varint
        ([1], 3, 7, 4),                          # const with data encoding.go:217
        ([], 2, 7, 4),                           # delta const without delta encoding.go:235
        ([2, 0], 2, 7, 4),                       # delta const with tail encoding.go:238
        ([1, 2, 3], 1, 0, 4),                    # not a zstd frame encoding.go:181
    ]
    for src, mt, first, n in cases:
        orc, _ = oracle.unmarshal_int64_array(np.array(src, dtype=np.uint8), mt, first, n)
        grc = rc_of(src, mt, first, n)
        assert (orc != 0) == (grc != 0), (src, mt, orc, grc)
        assert grc == orc, (src, mt, orc, grc)
    with pytest.raises(vm.VmbError):
        vm.encoding.unmarshal_values(np.zeros(0, dtype=np.uint8), 9, 0, 4)  # unknown MarshalType encoding.go:248


def test_decimal_to_float_kats_bit_exact(vm, kats, oracle):
    for va, e, exp in kats["append_decimal_to_float"]:
        if not va:
            continue
        got = vm.decimal.append_decimal_to_float(va, e)
        assert f64bits(got).tolist() == [struct.unpack("<Q", struct.pack("<d", gofloat(s)))[0] for s in exp], (va, e)
    rng = np.random.default_rng(SEED0 + 2)
    for e in list(range(-30, 31)) + [-300, -323, -324, 308, 309, -40, 35]:
        va = np.concatenate([rng.integers(-(1 << 62), 1 << 62, 500), rng.integers(-100000, 100000, 500),
                             np.array([0, 1, -1, I64_MAX, I64_MIN, I64_MAX - 1, I64_MAX - 2, I64_MIN + 1])]).astype(np.int64)
        got = vm.decimal.append_decimal_to_float(va, e)
        exp = oracle.decimal_to_float(va, e)
        assert np.array_equal(f64bits(got), f64bits(exp)), e


# ------------------------------------------------------------------------------------------------ codec: batched blocks
def _check_blocks(vm, blocks, tr_min=I64_MIN, tr_max=I64_MAX, as_int=False):
    descs, payload = blockgen.to_blockset(blocks)
    B = vm.storage.Blocks(descs, payload)
    series, status = vm.storage.decode_blocks(B, tr_min, tr_max, values_as_int64=as_int, raise_on_block_error=False)
    per = series.to_lists(np.int64 if as_int else np.float64)
    assert len(per) == len(blocks)
    for i, blk in enumerate(blocks):
        rc, ts, fv, iv = blk.oracle_unmarshal(tr_min, tr_max)
        if rc == -102:  # the Go code would logger.Panicf("BUG: ..."): any error will do
            assert int(status[i]) != 0, (i, blk.tmt, blk.vmt, blk.rows)
        else:
            assert int(status[i]) == rc, (i, blk.tmt, blk.vmt, blk.rows, int(status[i]), rc)
        if rc:
            continue
        gts, gv = per[i]
        assert np.array_equal(gts, ts), (i, "timestamps", blk.tmt, blk.rows)
        if as_int:
            assert np.array_equal(gv, iv), (i, "int values", blk.vmt, blk.rows)
        else:
            assert np.array_equal(f64bits(gv), f64bits(fv)), (i, "float values", blk.vmt, blk.rows, blk.scale)
    series.close()
    B.close()


def test_decode_blocks_mixed_types_bit_exact(vm):
    rng = np.random.default_rng(SEED0 + 3)
    blocks = blockgen.random_blocks(rng, 400)
    types = {(b.tmt, b.vmt) for b in blocks}
    assert {t for t, _ in types} >= {2, 5} and {v for _, v in types} >= {1, 2, 3, 4, 5, 6}
    _check_blocks(vm, blocks)
    _check_blocks(vm, blocks[:100], as_int=True)


def test_decode_blocks_lossy_precision_bits(vm):
    rng = np.random.default_rng(SEED0 + 4)
    blocks = blockgen.random_blocks(rng, 120, pbs=(1, 4, 12, 24, 40, 63), ts_kinds=["jitter", "irregular", "regular"])
    _check_blocks(vm, blocks)  # exercises EnsureNonDecreasingSequence encoding.go:258


def test_decode_blocks_time_range_filter(vm):
    rng = np.random.default_rng(SEED0 + 5)
    blocks = blockgen.random_blocks(rng, 120, ts_kinds=["regular", "jitter", "dups"])
    t0 = 1_700_000_000_000
    for tr in ((t0 + 15000 * 100, t0 + 15000 * 3000), (t0 - 5, t0 + 3), (t0 + 10 ** 12, t0 + 2 * 10 ** 12), (I64_MIN, t0 + 1000)):
        _check_blocks(vm, blocks, *tr)


def test_decode_blocks_corrupt_blocks_are_reported_per_block(vm):
    rng = np.random.default_rng(SEED0 + 6)
    blocks = blockgen.random_blocks(rng, 60, value_kinds=["counter", "gauge", "counter_big", "counter_smooth"], ts_kinds=["jitter"])
    for i in range(0, 60, 3):
        b = blocks[i]
        if i % 2 == 0 and b.vdata.size > 8:
            b.vdata = b.vdata.copy()
            b.vdata[b.vdata.size // 2] ^= 0x5A      # payload corruption
        elif b.tdata.size > 4:
            b.tdata = b.tdata[:-1].copy()           # truncated timestamps
    _check_blocks(vm, blocks)
    # timestamps out of [MinTimestamp, MaxTimestamp] => checkTimestampsBounds error block.go:298
    b = blocks[1]
    if b.tmt in (5, 6):
        b.max_ts -= 1
        _check_blocks(vm, [b])


def test_decode_blocks_fuzzed_payloads(vm):
    """one random mutation per block -- flipped / inserted / dropped bytes anywhere in either column (varints, zstd frame
    headers, Huffman trees, FSE tables, bitstreams), wrong row counts, swapped marshal types: the GPU path must neither crash
    nor disagree with the oracle about which blocks are bad, and the good ones must still decode bit for bit"""
    rng = np.random.default_rng(SEED0 + int(os.environ.get("VMB_FUZZ_SEED", "20240922")))  # other seeds for ad-hoc campaigns
    blocks = blockgen.random_blocks(rng, 700, rows_choices=(2, 3, 33, 100, 512, 1000, 4096, 8192))
    for i, b in enumerate(blocks):
        kind = int(rng.integers(0, 8))
        col = "vdata" if rng.random() < 0.7 else "tdata"
        d = getattr(b, col).copy()
        if kind <= 2 and d.size:            # flip 1..3 bytes
            for _ in range(int(rng.integers(1, 4))):
                d[int(rng.integers(0, d.size))] ^= int(rng.integers(1, 256))
        elif kind == 3 and d.size > 1:      # truncate
            d = d[: int(rng.integers(0, d.size))].copy()
        elif kind == 4:                     # append garbage
            d = np.concatenate([d, rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8)])
        elif kind == 5 and d.size > 2:      # drop a byte in the middle
            k = int(rng.integers(0, d.size))
            d = np.concatenate([d[:k], d[k + 1:]])
        elif kind == 6:                     # claim another row count
            b.rows = max(1, b.rows + int(rng.choice([-1, 1, 7, -5])))
        else:                               # claim another marshal type
            if col == "vdata":
                b.vmt = int(rng.integers(1, 7))
            else:
                b.tmt = int(rng.integers(1, 7))
        setattr(b, col, d)
    bad = 0
    for lo in range(0, len(blocks), 100):   # several batches: the mutations meet different warp / group neighbours
        part = blocks[lo:lo + 100]
        _check_blocks(vm, part)
        bad += sum(1 for b in part if b.oracle_unmarshal()[0] != 0)
    assert 150 < bad < 690  # the fuzz produces both broken and still-valid blocks


def test_zstd_golden_frames_through_gpu(vm, oracle):
    """frames written by the reference's libzstd (levels -5..5; Huffman 1/4 streams, FSE sequences, RLE, raw)"""
    with open(os.path.join(HERE, "golden", "zstd_frames.json")) as f:
        frames = json.load(f)
    ok = 0
    for fr in frames:
        c = np.frombuffer(base64.b64decode(fr["frame"]), dtype=np.uint8)
        rc, raw = oracle.zstd_decompress(c)
        assert rc == 0
        if raw[-1] >= 0x80:
            continue
        n = 1 + int(np.count_nonzero(raw < 0x80))
        if n > 16384:
            continue
        for mt in (1, 4):
            orc, exp = oracle.unmarshal_int64_array(c, mt, 12345, n)
            try:
                got = vm.encoding.unmarshal_values(c, mt, 12345, n)
                grc = 0
            except vm.VmbError as e:
                grc = e.code
            assert grc == orc, (fr["name"], fr["level"], grc, orc)
            if orc == 0:
                assert np.array_equal(got, exp), (fr["name"], fr["level"], mt)
                ok += 1
    assert ok > 100


def test_zstd_sequences_batched_stress(vm, oracle):
    """many libzstd frames with very different sequence sections in ONE batch (frames share warps in k_zstd_seq_decode /
    k_zstd_seq_exec): long literal runs, long matches, matches overlapping themselves (period 1..7), RLE / predefined /
    FSE-described tables, repeat offsets, frames with 1 sequence next to frames with thousands"""
    rng = np.random.default_rng(SEED0 + 20240921)
    blocks = []
    for i in range(700):
        n = int(rng.choice([2, 3, 17, 64, 129, 500, 1024, 3000, 8192]))
        shape = i % 7
        if shape == 0:      # straight line with rare kinks: delta2 is a run of zeros -> offset-1 matches
            d = np.full(n, int(rng.integers(1, 1000)), dtype=np.int64)
            d[rng.random(n) < 0.01] += rng.integers(1, 50)
            v = np.cumsum(d)
        elif shape == 1:    # periodic increments (period 2..7): self-overlapping matches
            per = int(rng.integers(2, 8))
            pat = rng.integers(0, 300, per)
            v = np.cumsum(np.tile(pat, n // per + 1)[:n]).astype(np.int64)
        elif shape == 2:    # noise with a long repeated segment: long matches at a large offset
            seg = rng.integers(0, 5000, max(n // 4, 1))
            d = np.concatenate([seg, rng.integers(0, 5000, max(n // 4, 1)), seg, seg])[:n]
            d = np.resize(d, n)
            v = np.cumsum(d).astype(np.int64)
        elif shape == 3:    # mostly constant gauge with spikes
            v = np.full(n, int(rng.integers(0, 10**6)), dtype=np.int64)
            k = rng.random(n) < 0.02
            v[k] += rng.integers(-1000, 1000, int(k.sum()))
        elif shape == 4:    # smooth counter
            v = blockgen.gen_values(rng, "counter_smooth", n)
        elif shape == 5:    # small-alphabet gauge
            v = blockgen.gen_values(rng, "gauge_small", n)
        else:               # steps: long runs of one delta, then another
            d = np.repeat(rng.integers(0, 100, n // 50 + 1), 50)[:n]
            v = np.cumsum(d).astype(np.int64)
        ts = blockgen.gen_timestamps(rng, str(rng.choice(["regular", "jitter", "irregular"])), n)
        blocks.append(blockgen.OBlock(ts, np.asarray(v, dtype=np.int64), int(rng.choice([-2, 0, 2])), 64, i))
    kinds = {b.vmt for b in blocks} | {b.tmt for b in blocks}
    assert {1, 4} & kinds  # zstd columns are present
    descs, payload = blockgen.to_blockset(blocks)
    B = vm.storage.Blocks(descs, payload)
    series, status = vm.storage.decode_blocks(B, values_as_int64=True)
    assert not status.any()
    got = series.to_lists(np.int64)
    series.close()
    for b, (gts, gv) in zip(blocks, got):
        assert np.array_equal(gts, b.ts), (b.series_idx, b.tmt)
        rc, ets, _, eiv = b.oracle_unmarshal()
        assert rc == 0 and np.array_equal(gv, eiv), (b.series_idx, b.vmt)


def test_blocks_written_by_the_product_encoder_decode_identically(vm, oracle):
    rng = np.random.default_rng(SEED0 + 7)
    bs = vm.storage.BlockSet()
    raw = []
    for i in range(64):
        n = int(rng.choice([1, 2, 100, 1024, 8192]))
        ts = blockgen.gen_timestamps(rng, str(rng.choice(blockgen.TS_KINDS)), n)
        vals = blockgen.gen_values(rng, str(rng.choice(blockgen.VALUE_KINDS)), n)
        bs.add(vm.storage.Block(ts, vals, scale=-2, series_idx=i))
        raw.append((ts, vals))
    descs, payload = bs.finish()
    B = vm.storage.Blocks(descs, payload)
    series, status = vm.storage.decode_blocks(B, values_as_int64=True)
    assert not status.any()
    for (ts, vals), (gts, gv) in zip(raw, series.to_lists(np.int64)):
        assert np.array_equal(gts, ts) and np.array_equal(gv, vals)


# ------------------------------------------------------------------------------------------------ rollup
def test_rollup_do_reference_kats_on_gpu(vm, kats):
    """every rollupConfig.Do sub-test of rollup_test.go through the CUDA path: values (1e-13 rel, rollup_test.go:1547)
    and samplesScanned"""
    n = 0
    for t in kats["rollup_do"]:
        name = GO_FUNC[t["func"]]
        values = [gofloat(x) for x in t["values"]]
        rc = vm.promql.RollupConfig(name, t["start"], t["end"], t["step"], t["window"], LookbackDelta=t["lookback_delta"],
                                    MayAdjustWindow=t["may_adjust_window"])
        out, scanned = rc.do(values, t["timestamps"])
        exp = [gofloat(x) for x in t["expected"]]
        assert len(out) == len(exp), t["test"]
        for g, e in zip(out.tolist(), exp):
            assert close(g, e, 1e-13), (t["test"], out.tolist(), exp)
        if t["samples_scanned"] is not None:
            assert scanned == t["samples_scanned"], (t["test"], scanned)
        n += 1
    assert n >= 50


def _random_series(rng, n, kind):
    t0 = 1_700_000_000_000
    if kind == "counter":
        ts = t0 + 15000 * np.arange(n) + rng.integers(-50, 51, n)
        v = np.cumsum(rng.integers(0, 1500, n)).astype(np.float64) / 100
        if n > 1:
            for r in rng.integers(1, n, max(n // 300, 1)):
                v[r:] -= v[r]
        return ts.astype(np.int64), v
    if kind == "gauge":
        ts = t0 + np.cumsum(rng.integers(5000, 25000, n))
        return ts.astype(np.int64), np.round(rng.normal(50, 3, n), 2)
    if kind == "gaps":
        ts = t0 + np.cumsum(np.where(rng.random(n) < 0.03, rng.integers(100000, 900000, n), 15000))
        return ts.astype(np.int64), np.round(rng.normal(5, 3, n), 1)
    if kind == "stale":
        ts = t0 + 15000 * np.arange(n)
        v = np.round(rng.normal(50, 3, n), 2)
        v[rng.integers(0, n, max(n // 20, 1))] = STALE_NAN
        return ts.astype(np.int64), v
    if kind == "dups":
        ts = t0 + 1000 * np.cumsum(rng.integers(0, 3, n))
        return ts.astype(np.int64), rng.integers(0, 5, n).astype(np.float64)
    raise KeyError(kind)


ARG_FUNCS = {"quantile_over_time": 0.9, "predict_linear": 60.0, "holt_winters": 0.5, "hoeffding_bound_lower": 0.9,
             "hoeffding_bound_upper": 0.9, "duration_over_time": 20.0, "count_le_over_time": 50.0, "count_gt_over_time": 50.0,
             "count_eq_over_time": 3.0, "count_ne_over_time": 3.0, "share_le_over_time": 50.0, "share_gt_over_time": 50.0,
             "share_eq_over_time": 3.0, "sum_le_over_time": 50.0, "sum_gt_over_time": 50.0, "sum_eq_over_time": 3.0}


@pytest.mark.parametrize("name", RF_IDS)
def test_rollup_function_differential_vs_oracle(vm, oracle, name):
    import zlib
    rng = np.random.default_rng(SEED0 + zlib.crc32(name.encode()))
    fid = RF[name]
    configs = [(300000, 15000, 0), (60000, 60000, 0), (0, 15000, 0), (3600000, 300000, 300000), (45000, 7000, 120000)]
    for kind in ("counter", "gauge", "gaps", "stale", "dups"):
        ts_list, v_list = [], []
        for n in (1, 2, 40, 700):
            ts, v = _random_series(rng, n, kind)
            ts_list.append(ts)
            v_list.append(v)
        for window, step, lookback in configs:
            start = int(ts_list[-1][0]) - 30000
            end = int(ts_list[-1][min(len(ts_list[-1]) - 1, 300)]) + 60000
            arg = ARG_FUNCS.get(name)
            rc = vm.promql.get_rollup_configs(name, start, end, step, window, lookback, args=arg,
                                              args2=0.3 if name == "holt_winters" else None)
            got, gscanned = rc.do_many(ts_list, v_list)
            escanned = 0
            for s, (ts, v) in enumerate(zip(ts_list, v_list)):
                v2, t2 = v.copy(), ts.copy()
                n2 = len(v2)
                if rc.dropStaleNaNs:
                    n2 = oracle.lib().vmo_drop_stale_nans(v2.ctypes.data_as(oracle.f64p), t2.ctypes.data_as(oracle.i64p), len(v2))
                v2, t2 = v2[:n2].copy(), t2[:n2].copy()
                if rc.removeCounterResets and n2:
                    oracle.lib().vmo_remove_counter_resets(v2.ctypes.data_as(oracle.f64p), t2.ctypes.data_as(oracle.i64p), n2,
                                                           lookback + window if lookback else 0)
                exp, sc = oracle.rollup_do(fid, v2, t2, start, end, step, window, lookback_delta=lookback,
                                           may_adjust_window=rc.MayAdjustWindow, is_default_rollup=rc.isDefaultRollup,
                                           samples_scanned_per_call=rc.samplesScannedPerCall, args=arg,
                                           args2=0.3 if name == "holt_winters" else None)
                escanned += sc
                rel = 1e-9 if name in ("geomean_over_time",) else 1e-12
                assert_allclose_nan(got[s], exp, rel, (name, kind, window, step, lookback, s))
            assert gscanned == escanned, (name, kind, window, step)


def test_remove_counter_resets_bit_exact(vm, oracle):
    """the preFunc of rate()/increase(): sequential float semantics must survive the warp-parallel formulation"""
    rng = np.random.default_rng(SEED0 + 8)
    ts_list, v_list = [], []
    for n in (1, 2, 31, 32, 33, 64, 1000, 8192):
        for kind in ("counter", "gauge", "gaps"):
            ts, v = _random_series(rng, n, kind)
            ts_list.append(ts)
            v_list.append(np.abs(v))
    for lookback, window in ((0, 300000), (60000, 30000), (1, 1)):
        series = vm.storage.Series.from_host(ts_list, v_list)
        rc = vm.promql.RollupConfig("last_over_time", int(ts_list[0][0]), int(ts_list[0][0]) + 1000, 1000, window,
                                    LookbackDelta=lookback, removeCounterResets=True)
        rc.do_series(series)
        got = series.to_lists()
        for (ts, v), (gts, gv) in zip(zip(ts_list, v_list), got):
            e = v.copy()
            oracle.lib().vmo_remove_counter_resets(e.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), len(e),
                                                   lookback + window if lookback else 0)
            assert np.array_equal(f64bits(gv), f64bits(e)), (len(v), lookback)
        series.close()


def test_incremental_aggregates(vm, oracle):
    import torch
    rng = np.random.default_rng(SEED0 + 9)
    S, P, G = 97, 50, 5
    ts_list, v_list = [], []
    for s in range(S):
        ts, v = _random_series(rng, int(rng.integers(1, 400)), "gaps" if s % 3 else "counter")
        ts_list.append(ts)
        v_list.append(v)
    start = int(min(t[0] for t in ts_list))
    rc = vm.promql.get_rollup_configs("avg_over_time", start, start + 60000 * (P - 1), 60000, 120000)
    rolled, _ = rc.do_many(ts_list, v_list)
    groups = rng.integers(0, G, S).astype(np.uint32)

    class Buf:
        def __init__(self, nbytes):
            self.t = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
            self.ptr = self.t.data_ptr()

    ctx = vm.default_context()
    for name in ("sum", "min", "max", "avg", "count", "sum2", "geomean", "any", "group"):
        ia = vm.promql.IncrementalAggr(name, G, P, Buf)
        series = vm.storage.Series.from_host(ts_list, v_list)
        ia.update(series, rc, groups)
        torch.cuda.synchronize()
        got = ia.finalize(ctx)
        exp_v = np.zeros((G, P))
        exp_c = np.zeros((G, P))
        for s in range(S):
            g = int(groups[s])
            row = np.ascontiguousarray(rolled[s])
            oracle.lib().vmo_aggr_update(AGGR[name], exp_v[g].ctypes.data_as(oracle.f64p), exp_c[g].ctypes.data_as(oracle.f64p),
                                         row.ctypes.data_as(oracle.f64p), P)
        for g in range(G):
            oracle.lib().vmo_aggr_finalize(AGGR[name], exp_v[g].ctypes.data_as(oracle.f64p), exp_c[g].ctypes.data_as(oracle.f64p), P)
        assert_allclose_nan(got, exp_v, 1e-9 if name == "geomean" else 1e-15, name)
        series.close()


# ------------------------------------------------------------------------------------------------ whole path
def _oracle_pipeline(oracle, blocks, func, start, end, step, window, lookback, tr_min=I64_MIN, tr_max=I64_MAX):
    import victoriametrics_b200 as vm
    rc = vm.promql.get_rollup_configs(func, start, end, step, window, lookback)
    out = []
    for blk in blocks:
        r, ts, fv, _ = blk.oracle_unmarshal(tr_min, tr_max)
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
                                samples_scanned_per_call=rc.samplesScannedPerCall)
        out.append(o)
    return np.stack(out)


@pytest.mark.parametrize("func,window,step", [("rate", 300000, 15000), ("increase", 3600000, 60000),
                                              ("avg_over_time", 300000, 15000), ("max_over_time", 300000, 60000),
                                              ("default_rollup", 0, 30000)])
def test_eval_rollup_whole_path_host_and_device(vm, oracle, func, window, step):
    import torch
    rng = np.random.default_rng(SEED0 + 10)
    blocks = blockgen.random_blocks(rng, 150, rows_choices=(2, 100, 1000, 8192),
                                    value_kinds=["counter", "counter_resets", "gauge", "const", "delta_const", "counter_smooth"],
                                    ts_kinds=["regular", "jitter"], scales=(-2,))
    t0 = 1_700_000_000_000
    start, end = t0 + 300000, t0 + 15000 * 2000
    exp = _oracle_pipeline(oracle, blocks, func, start, end, step, window, 0)
    descs, payload = blockgen.to_blockset(blocks)
    got_host, scanned = vm.promql.eval_rollup_func_host(func, descs, payload, start, end, step, window)
    assert_allclose_nan(got_host, exp, 1e-12, func + " host path")
    B = vm.storage.Blocks(descs, payload)
    out = torch.empty(exp.shape, dtype=torch.float64, device="cuda")
    _, scanned2 = vm.promql.eval_rollup_func(func, B, start, end, step, window, out_dev_ptr=out.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(f64bits(out.cpu().numpy()), f64bits(got_host))
    assert scanned == scanned2


def test_delta_const_value_columns_that_wrap_are_counter_reset_candidates(vm, oracle):
    """A values column stored as MarshalTypeDeltaConst can decrease although its delta is positive: first + i*d wraps in
    int64 exactly like the Go loop (encoding.go:240), e.g. the two-row column {5216, MinInt64+1}.  removeCounterResets must
    still run on such a series (found by the seed-shifted campaign, VMB_SEED_OFFSET=1000)."""
    t0 = 1_700_000_000_000
    cols = [np.array([5216, -(1 << 63) + 1], dtype=np.int64),                       # positive wrapped delta, value drops
            np.array([(1 << 62) + 5, -(1 << 63) + 1005, -(1 << 62) + 2005], dtype=np.int64),  # delta 2^62+1000: wraps after row 0
            np.array([10, 7, 4, 1], dtype=np.int64),                                  # plain negative delta
            np.array([-5, 0, 5, 10], dtype=np.int64),                                 # increasing: no candidate
            np.array([(1 << 63) - 3, -(1 << 63) + 1], dtype=np.int64)]               # increasing delta that wraps to a drop
    blocks = []
    for i, v in enumerate(cols):
        ts = t0 + 15_000 * np.arange(len(v), dtype=np.int64)
        b = blockgen.OBlock(ts, v, -2, 64, i)
        assert b.vmt == 2, (i, b.vmt)  # MarshalTypeDeltaConst
        blocks.append(b)
    start, end, step, window = t0 + 15_000, t0 + 120_000, 15_000, 60_000
    descs, payload = blockgen.to_blockset(blocks)
    for func in ("rate", "increase", "irate"):
        exp = _oracle_pipeline(oracle, blocks, func, start, end, step, window, 0)
        got, _ = vm.promql.eval_rollup_func_host(func, descs, payload, start, end, step, window)
        assert_allclose_nan(got, exp, 1e-12, func)
