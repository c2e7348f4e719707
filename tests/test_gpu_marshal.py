"""The write path on the GPU (csrc/encode.cu, vmb_marshal_columns_gpu): encoding.MarshalValues for many columns at once.
For every kind of column the reference's type selection (encoding.go:119-171) must come out -- MarshalType equal to the oracle's --
the bytes must equal the host encoder's (same zstd writer), and the oracle (reference decoder restated) must decode them to the
input, bit for bit when precisionBits = 64, to exactly what the oracle's own lossy encoder produces when it is lower."""
import numpy as np
import pytest

import blockgen
from conftest import SEED0

pytestmark = pytest.mark.gpu


def _columns(rng, kind, ncols, rows):
    return np.stack([blockgen.gen_values(rng, kind, rows) for _ in range(ncols)])


@pytest.mark.parametrize("kind", blockgen.VALUE_KINDS + ["timestamps_regular", "timestamps_jitter"])
@pytest.mark.parametrize("rows", [8192, 1000, 130, 2, 1])
def test_gpu_marshal_equals_host_and_reference_type_selection(oracle, kind, rows):
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 300 + rows)
    ncols = 24
    if kind.startswith("timestamps_"):
        a = np.stack([blockgen.gen_timestamps(rng, kind.split("_")[1], rows) for _ in range(ncols)])
    else:
        a = _columns(rng, kind, ncols, rows)
    ctx = vm.default_context()
    p_gpu, o_gpu, mt_gpu, f_gpu = vm.encoding.marshal_columns(a, 64, ctx=ctx)
    p_cpu, o_cpu, mt_cpu, f_cpu = vm.encoding.marshal_columns(a, 64)
    assert np.array_equal(mt_gpu, mt_cpu) and np.array_equal(f_gpu, f_cpu) and np.array_equal(o_gpu, o_cpu)
    assert np.array_equal(p_gpu, p_cpu)
    have_ref = bool(oracle.lib().vmo_zstd_ref_available())
    for c in range(ncols):
        b = p_gpu[int(o_gpu[c]):int(o_gpu[c + 1])]
        rc, out = oracle.unmarshal_int64_array(b, int(mt_gpu[c]), int(f_gpu[c]), rows)
        assert rc == 0 and np.array_equal(out, a[c]), (kind, c)
        if have_ref:
            _, omt, ofirst = oracle.marshal_int64_array(a[c])
            # the reference compresses with libzstd: the type before the zstd stage (1/5 -> nearest delta2, 4/6 -> nearest delta,
            # 2, 3) must agree; whether zstd wins the 0.9 cut may differ for nearly incompressible streams (other compressor)
            fam = {1: "d2", 5: "d2", 4: "d", 6: "d", 2: "dc", 3: "c"}
            assert fam[int(mt_gpu[c])] == fam[omt] and ofirst == int(f_gpu[c]), (kind, c, int(mt_gpu[c]), omt)


@pytest.mark.parametrize("pb", [1, 4, 5, 6, 12, 20, 33, 63])
def test_gpu_marshal_lossy_precision_bits(oracle, pb):
    """precisionBits < 64: nearestDelta's trailing-zeros state machine (nearest_delta.go:83-125) on the GPU == the oracle's"""
    import victoriametrics_b200 as vm
    rng = np.random.default_rng(SEED0 + 400 + pb)
    ctx = vm.default_context()
    for kind in ("counter", "counter_resets", "gauge", "gauge_wide", "counter_big", "counter_smooth"):
        a = _columns(rng, kind, 12, 777)
        p_gpu, o_gpu, mt_gpu, f_gpu = vm.encoding.marshal_columns(a, pb, ctx=ctx)
        p_cpu, o_cpu, mt_cpu, f_cpu = vm.encoding.marshal_columns(a, pb)
        assert np.array_equal(mt_gpu, mt_cpu) and np.array_equal(p_gpu, p_cpu) and np.array_equal(o_gpu, o_cpu), (kind, pb)
        for c in range(a.shape[0]):
            b = p_gpu[int(o_gpu[c]):int(o_gpu[c + 1])]
            rc, out = oracle.unmarshal_int64_array(b, int(mt_gpu[c]), int(f_gpu[c]), a.shape[1])
            assert rc == 0
            ob, omt, ofirst = oracle.marshal_int64_array(a[c], pb)
            rc2, oout = oracle.unmarshal_int64_array(ob, omt, ofirst, a.shape[1])
            assert rc2 == 0 and np.array_equal(out, oout), (kind, pb, c)  # the same lossy sequence as the reference's encoder


def test_gpu_marshal_blocks_decode_and_roll_up(oracle):
    """round trip through both GPU halves: columns marshaled on the GPU, uploaded as blocks, decoded + rolled up by the read path"""
    import victoriametrics_b200 as vm
    from test_baseline_configs import T0
    rng = np.random.default_rng(SEED0 + 500)
    rows, n = 4096, 40
    vals = _columns(rng, "counter", n, rows)
    ctx = vm.default_context()
    payload, offs, mts, firsts = vm.encoding.marshal_columns(vals, 64, ctx=ctx)
    ts = T0 + 15000 * np.arange(rows, dtype=np.int64)
    tdata, tmt, tfirst = vm.encoding.marshal_timestamps(ts)
    arena = np.concatenate([tdata, payload])
    descs = vm.storage.descs_from_arrays(first_value=firsts, val_off=offs[:-1] + tdata.size, val_size=np.diff(offs).astype(np.uint32),
                                         rows=np.full(n, rows, dtype=np.uint32), series_idx=np.arange(n, dtype=np.uint32), scale=-2,
                                         val_mt=mts, precision_bits=64, min_ts=tfirst, max_ts=int(ts[-1]), ts_off=0, ts_size=tdata.size,
                                         ts_mt=tmt)
    start, end, step, window = T0 + 300000, T0 + 15000 * (rows - 1), 15000, 300000
    got, _ = vm.promql.eval_rollup_func_host("increase", descs, arena, start, end, step, window)
    from rollup_names import RF
    for s in (0, 7, n - 1):
        fv = oracle.decimal_to_float(vals[s], -2)
        oracle.lib().vmo_remove_counter_resets(fv.ctypes.data_as(oracle.f64p), ts.ctypes.data_as(oracle.i64p), rows, 0)
        exp, _ = oracle.rollup_do(RF["increase"], fv, ts, start, end, step, window, samples_scanned_per_call=2)
        assert np.allclose(got[s], exp, rtol=1e-12, atol=0, equal_nan=True)


def test_gpu_float_to_decimal_columns(oracle, kats):
    """decimal.AppendFloatToDecimal on the GPU == the host encoder == the oracle, per column: mantissas and common exponent"""
    import victoriametrics_b200 as vm
    from conftest import STALE_NAN
    rng = np.random.default_rng(SEED0 + 600)
    cols = []
    rows = 257
    for k in range(40):
        kind = k % 8
        if kind == 0:
            c = np.round(rng.normal(50, 20, rows), 2)
        elif kind == 1:
            c = rng.integers(0, 10 ** 9, rows).astype(np.float64)
        elif kind == 2:
            c = rng.normal(0, 1, rows) * 10.0 ** rng.integers(-12, 12, rows)
        elif kind == 3:
            c = np.zeros(rows)
        elif kind == 4:
            c = np.ones(rows)
        elif kind == 5:
            c = np.round(rng.normal(0, 1e6, rows), 3)
            c[rng.integers(0, rows, 5)] = [np.inf, -np.inf, STALE_NAN, 0.0, 1e300]
        elif kind == 6:
            c = rng.integers(-5, 5, rows) * 0.001
        else:
            c = np.cumsum(rng.integers(0, 1500, rows)) / 100.0
        cols.append(c)
    src = np.stack(cols)
    ctx = vm.default_context()
    got, scales = vm.decimal.append_float_to_decimal_columns(src, ctx)
    for k in range(src.shape[0]):
        hv, he = vm.decimal.append_float_to_decimal(src[k])
        ov, oe = oracle.float_to_decimal(src[k])
        assert he == oe and np.array_equal(hv, ov)
        assert int(scales[k]) == oe and np.array_equal(got[k], ov), k
