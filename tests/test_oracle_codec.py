"""Pins the CPU oracle's codec against the reference's own known-answer vectors (tests/golden/go_kats.json,
extracted from lib/encoding/*_test.go and lib/decimal/decimal_test.go) + round-trip properties mirroring
nearest_delta2_test.go:262 checkPrecisionBits and int_test.go:186."""
import ctypes as C
import struct

import numpy as np
import pytest
from conftest import gofloat


def bits(f):
    return struct.unpack("<Q", struct.pack("<d", f))[0]


def test_nearest_delta_kats(kats, oracle):
    for nxt, prev, pb, d_exp, tz_exp in kats["nearest_delta"]:
        d, tz = oracle.nearest_delta(nxt, prev, pb)
        assert (d, tz) == (d_exp, tz_exp), (nxt, prev, pb)


@pytest.mark.parametrize("key,delta2", [("marshal_nearest_delta", False), ("marshal_nearest_delta2", True)])
def test_marshal_nearest_delta_kats(kats, oracle, key, delta2):
    for va, pb, first_exp, hex_exp in kats[key]:
        b, first = oracle.marshal_nearest_delta(va, pb, delta2=delta2)
        assert first == first_exp
        assert bytes(b).hex() == hex_exp, (va, pb)


def test_type_detection_kats(kats, oracle):
    L = oracle.lib()
    for key, fn in (("is_const", L.vmo_is_const), ("is_delta_const", L.vmo_is_delta_const), ("is_gauge", L.vmo_is_gauge)):
        for a, exp in kats[key]:
            arr = np.array(a, dtype=np.int64)
            assert bool(fn(arr.ctypes.data_as(oracle.i64p), len(a))) == exp, (key, a)


def test_ensure_non_decreasing_kats(kats, oracle):
    for a, vmin, vmax, exp in kats["ensure_non_decreasing"]:
        arr = np.array(a, dtype=np.int64)
        oracle.lib().vmo_ensure_non_decreasing(arr.ctypes.data_as(oracle.i64p), len(a), vmin, vmax)
        assert arr.tolist() == exp


def check_precision_bits(a, b, pb):
    """nearest_delta2_test.go:262 checkPrecisionBits"""
    assert len(a) == len(b)
    for va, vb in zip(a.tolist(), b.tolist()):
        if va < vb:
            va, vb = vb, va
        eps = va - vb
        if eps == 0:
            continue
        if va < 0:
            va = -vb
        pbe = 1
        while eps < va:
            va >>= 1
            pbe += 1
        assert pbe >= pb, (va, vb, pb, pbe)


def test_marshal_array_generic_kats(kats, oracle):
    for va, pb, mt_exp in kats["marshal_array_generic"]:
        b, mt, first = oracle.marshal_int64_array(va, pb)
        assert mt == mt_exp, va
        rc, out = oracle.unmarshal_int64_array(b, mt, first, len(va))
        assert rc == 0
        check_precision_bits(np.array(va, dtype=np.int64), out, pb)


def test_varint_roundtrip_edges(oracle):
    # int_test.go:186 TestMarshalUnmarshalVarInt64 + edges
    vals = [0, 1, -1, 63, -63, 64, -64, 1 << 6, -(1 << 6), (1 << 13) - 1, 1 << 13, -(1 << 13), (1 << 63) - 1, -(1 << 63),
            (1 << 62), -(1 << 62), 1 << 20, 1 << 27, 1 << 34, 1 << 41, 1 << 48, 1 << 55, -(1 << 55) - 1]
    vals += [int(x) for x in np.random.default_rng(1).integers(-(1 << 63), (1 << 63) - 1, 500)]
    for shift in range(0, 63):
        vals += [1 << shift, -(1 << shift), (1 << shift) - 1]
    arr = np.array(vals, dtype=np.int64)
    b = oracle.marshal_varint64s(arr)
    rc, out, consumed = oracle.unmarshal_varint64s(b, len(arr))
    assert rc == 0 and consumed == len(b)
    assert np.array_equal(out, arr)
    # sizes: zig-zag LEB128
    for v in (0, 63, -64):
        assert len(oracle.marshal_varint64s([v])) == 1
    for v in (64, -65, 8191, -8192):
        assert len(oracle.marshal_varint64s([v])) == 2
    assert len(oracle.marshal_varint64s([-(1 << 63)])) == 10


def test_varint_errors(oracle):
    # too small src (int.go:183)
    rc, _, _ = oracle.unmarshal_varint64s(np.array([1, 2], dtype=np.uint8), 3)
    assert rc == -1
    # truncated multi-byte varint
    rc, _, _ = oracle.unmarshal_varint64s(np.array([0x80], dtype=np.uint8), 1)
    assert rc == -1
    # 10th byte > 1 => too big (int.go:271)
    rc, _, _ = oracle.unmarshal_varint64s(np.array([0xFF] * 9 + [0x02], dtype=np.uint8), 1)
    assert rc == -2
    # 11 bytes => too long (int.go:277)
    rc, _, _ = oracle.unmarshal_varint64s(np.array([0xFF] * 10 + [0x01], dtype=np.uint8), 1)
    assert rc == -3
    # trailing bytes after nearest-delta payload => error (nearest_delta.go:65)
    b, first = oracle.marshal_nearest_delta([1, 5, 9, 200], 64)
    rc, _ = oracle.unmarshal_nearest_delta(np.concatenate([b, np.array([0], dtype=np.uint8)]), first, 4)
    assert rc == -4


@pytest.mark.parametrize("pb", [1, 4, 8, 16, 23, 24, 32, 48, 63, 64])
def test_marshal_unmarshal_roundtrip_property(oracle, pb):
    # mirrors encoding_cgo_test.go:10 TestMarshalUnmarshalInt64Array with our own RNG
    rng = np.random.default_rng(pb)
    n = 8 * 1024
    v = 0
    va = np.empty(n, dtype=np.int64)
    noise = rng.normal(0, 1e2, n)
    for i in range(n):
        v += 30e3 + int(noise[i])
        va[i] = int(v)
    b, mt, first = oracle.marshal_int64_array(va, pb)
    assert mt in (1, 5)
    rc, out = oracle.unmarshal_int64_array(b, mt, first, n)
    assert rc == 0
    check_precision_bits(va, out, pb)
    if pb == 64:
        assert np.array_equal(out, va)
    # gauge
    ga = (1000 * rng.normal(0, 2e5, n)).astype(np.int64)
    b, mt, first = oracle.marshal_int64_array(ga, pb)
    assert mt in (4, 6)
    rc, out = oracle.unmarshal_int64_array(b, mt, first, n)
    assert rc == 0
    check_precision_bits(ga, out, min(pb + 2, 64) if pb < 6 else pb)


def test_const_and_delta_const(oracle):
    b, mt, first = oracle.marshal_int64_array([7] * 100)
    assert (len(b), mt, first) == (0, 3, 7)
    rc, out = oracle.unmarshal_int64_array(b, mt, first, 100)
    assert rc == 0 and out.tolist() == [7] * 100
    va = [1000 + 15000 * i for i in range(8192)]
    b, mt, first = oracle.marshal_int64_array(va)
    assert mt == 2 and first == 1000
    rc, out = oracle.unmarshal_int64_array(b, mt, first, 8192)
    assert rc == 0 and out.tolist() == va
    # const with trailing bytes => error (encoding.go:217)
    rc, _ = oracle.unmarshal_int64_array(np.array([1], dtype=np.uint8), 3, 7, 4)
    assert rc == -7
    rc, _ = oracle.unmarshal_int64_array(np.array([], dtype=np.uint8), 9, 7, 4)
    assert rc == -5


# ---------------------------------------------------------------- decimal

def test_append_decimal_to_float_kats_bit_exact(kats, oracle):
    for va, e, exp in kats["append_decimal_to_float"]:
        out = oracle.decimal_to_float(va, e)
        assert [bits(x) for x in out] == [bits(gofloat(s)) for s in exp], (va, e)


def test_positive_float_to_decimal_kats(kats, oracle):
    for f, v_exp, e_exp in kats["positive_float_to_decimal"]:
        v = C.c_int64(0)
        e = C.c_int16(0)
        oracle.lib().vmo_positive_float_to_decimal(gofloat(f), C.byref(v), C.byref(e))
        assert (v.value, e.value) == (v_exp, e_exp), f


def test_from_float_kats(kats, oracle):
    for f, v_exp, e_exp in kats["from_float"]:
        v = C.c_int64(0)
        e = C.c_int16(0)
        oracle.lib().vmo_from_float(gofloat(f), C.byref(v), C.byref(e))
        assert (v.value, e.value) == (v_exp, e_exp), f


def test_append_float_to_decimal_kats(kats, oracle):
    for fa, da_exp, e_exp in kats["append_float_to_decimal"]:
        da, e = oracle.float_to_decimal([gofloat(s) for s in fa])
        assert e == e_exp and da.tolist() == da_exp, fa


def test_calibrate_scale_kats(kats, oracle):
    L = oracle.lib()
    for a, b, ae, be, a_exp, b_exp, e_exp in kats["calibrate_scale"]:
        for rev in (False, True):
            aa = np.array(a, dtype=np.int64)
            bb = np.array(b, dtype=np.int64)
            if not rev:
                e = L.vmo_calibrate_scale(aa.ctypes.data_as(oracle.i64p), len(a), ae, bb.ctypes.data_as(oracle.i64p), len(b), be)
            else:
                e = L.vmo_calibrate_scale(bb.ctypes.data_as(oracle.i64p), len(b), be, aa.ctypes.data_as(oracle.i64p), len(a), ae)
            assert e == e_exp and aa.tolist() == a_exp and bb.tolist() == b_exp, (a, b, ae, be, rev)


def test_float_decimal_roundtrip(oracle):
    # decimal_test.go:468 TestFloatToDecimalRoundtrip style
    for f in [0, 1, 0.123, 1.2345, 12000, 1e-30, 1e30, 1234567890123, 12.34567890125, 15e18, 0.000874957]:
        for sign in (1, -1):
            v = C.c_int64(0)
            e = C.c_int16(0)
            oracle.lib().vmo_from_float(sign * f, C.byref(v), C.byref(e))
            back = oracle.lib().vmo_to_float(v.value, e.value)
            assert back == pytest.approx(sign * f, rel=1e-12)


def test_block_header_roundtrip(oracle):
    bh = oracle.BlockHeader()
    for i in range(24):
        bh.tsid[i] = i + 1
    bh.min_ts, bh.max_ts, bh.first_value = -5, 1 << 40, -(1 << 62)
    bh.ts_off, bh.val_off, bh.ts_size, bh.val_size, bh.rows = 123456789012, 99, 17, 4321, 8192
    bh.scale, bh.ts_mt, bh.val_mt, bh.precision_bits = -13, 2, 1, 64
    buf = (C.c_uint8 * 81)()
    oracle.lib().vmo_block_header_marshal(buf, C.byref(bh))
    # FirstValue is zig-zag big-endian (int.go:69): -(1<<62) -> 0x7fff...ff
    assert bytes(buf[40:48]).hex() == "7fffffffffffffff"
    assert bytes(buf[76:78]).hex() == "0019"  # scale -13 zig-zag = 25
    bh2 = oracle.BlockHeader()
    oracle.lib().vmo_block_header_unmarshal(C.byref(bh2), buf)
    for f, _ in oracle.BlockHeader._fields_:
        a, b = getattr(bh, f), getattr(bh2, f)
        assert (bytes(a) == bytes(b)) if f == "tsid" else (a == b), f


def test_pow10_is_the_published_go_table_product(oracle):
    """Go's math.Pow10 (src/math/pow10.go, unchanged since Go 1.11):
         0 <= n <= 308:   pow10postab32[n/32] * pow10tab[n%32]      (tables of the decimal literals 1e0..1e31 and 1e0,1e32,..,1e288)
         -323 <= n <= 0:  pow10negtab32[-n/32] / pow10tab[-n%32]    (1e-0, 1e-32, .., 1e-320)
         else +Inf / 0.
    Every entry rebuilt here from correctly rounded literals (Python's float() is IEEE round-to-nearest like the Go compiler's
    constant conversion) -- for n >= 32 the product can differ from the literal 1eN by one ulp, which is what decimal.go:100 sees."""
    L = oracle.lib()
    differs_from_literal = 0
    for n in range(-330, 316):
        got = L.vmo_pow10(n)
        if 0 <= n <= 308:
            exp = float("1e%d" % (32 * (n // 32))) * float("1e%d" % (n % 32))
            differs_from_literal += exp != float("1e%d" % n)
        elif -323 <= n <= 0:
            exp = float("1e-%d" % (32 * ((-n) // 32))) / float("1e%d" % ((-n) % 32))
        else:
            exp = float("inf") if n > 0 else 0.0
        assert got == exp and np.float64(got).view(np.uint64) == np.float64(exp).view(np.uint64), n
    assert differs_from_literal > 0  # the table product is not the same thing as the literal: the restatement must use the product
