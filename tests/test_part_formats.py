"""SURVEY.md 8(f).2-3, host side: the wire forms around the hot path -- blockHeader (block_header.go:104/:122), index blocks
(unmarshalBlockHeaders :261), metaindex rows (metaindex_row.go:61/:72/:129) and decimal.CalibrateScale (decimal.go:13) --
checked against an independent struct.pack of the documented layouts, the oracle, and the reference's own test vectors.
No GPU: these entry points are plain host code in libvmb200."""
import ctypes as C

import numpy as np
import pytest

import partgen
from victoriametrics_b200 import _lib, decimal as vdecimal, storage


def _rand_header(rng):
    return dict(min_ts=int(rng.integers(-2**62, 2**62)), max_ts=int(rng.integers(-2**62, 2**62)),
                first_value=int(rng.integers(-2**63, 2**63 - 1)), ts_off=int(rng.integers(0, 2**62)),
                val_off=int(rng.integers(0, 2**62)), ts_size=int(rng.integers(0, 131073)), val_size=int(rng.integers(0, 131073)),
                rows=int(rng.integers(1, 16385)), scale=int(rng.integers(-32768, 32768)), ts_mt=int(rng.integers(0, 7)),
                val_mt=int(rng.integers(0, 7)), precision_bits=int(rng.integers(1, 65)))


def _desc(h):
    d = np.zeros(1, dtype=storage.DESC_DTYPE)
    for k, v in h.items():
        d[k] = v
    return d


def test_block_header_marshal_matches_layout_and_oracle(oracle):
    rng = np.random.default_rng(3)
    for _ in range(300):
        h = _rand_header(rng)
        tsid = partgen.pack_tsid(*[int(x) for x in rng.integers(0, 2**31, 4)])
        want = partgen.pack_header(tsid, h)
        assert len(want) == 81  # block_header_test.go:10 TestMarshaledBlockHeaderSize
        got = storage.marshal_block_header(_desc(h)[0], tsid)
        assert got == want, h
        # the oracle's own writer agrees (tsid zero there)
        bh = oracle.BlockHeader()
        for k in ("min_ts", "max_ts", "first_value", "ts_size", "val_size", "rows", "scale", "ts_mt", "val_mt", "precision_bits"):
            setattr(bh, k, h[k])
        bh.ts_off, bh.val_off = h["ts_off"], h["val_off"]
        buf = (C.c_uint8 * 81)()
        oracle.lib().vmo_block_header_marshal(buf, C.byref(bh))
        assert bytes(buf)[24:] == got[24:]
        # and back: blockHeader.Unmarshal (block_header_test.go:20 round trip)
        back, ids = storage.unmarshal_block_headers(got, 1)
        for k, v in h.items():
            assert int(back[0][k]) == v, (k, h)
        assert ids[0].tobytes() == tsid


def test_index_block_unmarshal_checks():
    rng = np.random.default_rng(4)
    hs = [_rand_header(rng) for _ in range(50)]
    tsids = sorted(partgen.pack_tsid(1, 2, 3, int(m)) for m in rng.integers(0, 2**40, 50))
    data = b"".join(partgen.pack_header(t, h) for t, h in zip(tsids, hs))
    descs, ids = storage.unmarshal_block_headers(data, 50)
    assert [int(x) for x in descs["rows"]] == [h["rows"] for h in hs]
    assert [i.tobytes() for i in ids] == tsids
    # invalid number of block headers (block_header.go:281)
    with pytest.raises(_lib.VmbError):
        storage.unmarshal_block_headers(data, 49)
    with pytest.raises(_lib.VmbError):
        storage.unmarshal_block_headers(data[:-1], 50)
    # not sorted by tsid (:286)
    swapped = data[81:162] + data[:81] + data[162:]
    if tsids[0] != tsids[1]:
        with pytest.raises(_lib.VmbError):
            storage.unmarshal_block_headers(swapped, 50)
    # equal TSIDs are fine (several blocks of one series)
    same = b"".join(partgen.pack_header(tsids[0], h) for h in hs[:5])
    assert len(storage.unmarshal_block_headers(same, 5)[0]) == 5
    # blockHeader.validate (:230): zero rows, too many rows, bad marshal type, bad precision bits, oversized payloads
    for field, bad in (("rows", 0), ("rows", 16385), ("ts_mt", 7), ("val_mt", 9), ("precision_bits", 0), ("precision_bits", 65),
                       ("ts_size", 131073), ("val_size", 131073)):
        h = dict(hs[0])
        h[field] = bad
        with pytest.raises(_lib.VmbError):
            storage.unmarshal_block_headers(partgen.pack_header(tsids[0], h), 1)


def test_metaindex_rows_roundtrip_and_checks():
    rng = np.random.default_rng(5)
    tsids = sorted(partgen.pack_tsid(7, 1, 1, int(m)) for m in rng.integers(0, 2**40, 40))
    rows = [(t, int(rng.integers(1, 1000)), int(rng.integers(-2**62, 0)), int(rng.integers(0, 2**62)), int(rng.integers(0, 2**50)),
             int(rng.integers(1, 131073))) for t in tsids]
    data = b"".join(partgen.pack_metaindex_row(*r) for r in rows)
    assert len(data) == 56 * 40
    got = storage.unmarshal_metaindex_rows(data)
    assert len(got) == 40
    for g, r in zip(got, rows):
        assert (g["tsid"].tobytes(), int(g["block_headers_count"]), int(g["min_ts"]), int(g["max_ts"]),
                int(g["index_block_offset"]), int(g["index_block_size"])) == r
        assert storage.marshal_metaindex_row(g) == partgen.pack_metaindex_row(*r)  # metaindex_row_test.go:32 round trip
    for bad in (b"", data[:-1], data[:55]):  # zero rows (:154), truncated row
        with pytest.raises(_lib.VmbError):
            storage.unmarshal_metaindex_rows(bad)
    with pytest.raises(_lib.VmbError):  # BlockHeadersCount must be > 0 (:112)
        storage.unmarshal_metaindex_rows(partgen.pack_metaindex_row(tsids[0], 0, 0, 1, 0, 10))
    with pytest.raises(_lib.VmbError):  # too big IndexBlockSize (:115)
        storage.unmarshal_metaindex_rows(partgen.pack_metaindex_row(tsids[0], 1, 0, 1, 0, 131073))
    if tsids[0] != tsids[1]:
        with pytest.raises(_lib.VmbError):  # sorted by TSID (:160)
            storage.unmarshal_metaindex_rows(data[56:112] + data[:56])


def test_calibrate_scale_reference_vectors(kats):
    """decimal_test.go TestCalibrateScale vectors, both argument orders like the Go test helper"""
    assert len(kats["calibrate_scale"]) >= 30
    for a, b, ae, be, a_exp, b_exp, e_exp in kats["calibrate_scale"]:
        a2, b2, e = vdecimal.calibrate_scale(a, ae, b, be)
        assert (a2.tolist(), b2.tolist(), e) == (a_exp, b_exp, e_exp), (a, b, ae, be)
        b3, a3, e3 = vdecimal.calibrate_scale(b, be, a, ae)
        assert (a3.tolist(), b3.tolist(), e3) == (a_exp, b_exp, e_exp), (a, b, ae, be)


def test_calibrate_scale_random_vs_oracle(oracle):
    rng = np.random.default_rng(6)
    specials = [2**63 - 1, -2**63, 2**63 - 2]  # vInfPos, vInfNeg, vStaleNaN (decimal.go:403)
    for _ in range(400):
        na, nb = int(rng.integers(0, 6)), int(rng.integers(0, 6))
        mag = int(rng.integers(1, 63))
        a = [int(x) for x in rng.integers(-2**mag, 2**mag, na)]
        b = [int(x) for x in rng.integers(-2**mag, 2**mag, nb)]
        if na and rng.random() < 0.2:
            a[0] = specials[int(rng.integers(0, 3))]
        if nb and rng.random() < 0.2:
            b[-1] = specials[int(rng.integers(0, 3))]
        ae, be = int(rng.integers(-30, 30)), int(rng.integers(-30, 30))
        aa, bb = np.array(a, dtype=np.int64), np.array(b, dtype=np.int64)
        e_exp = oracle.lib().vmo_calibrate_scale(aa.ctypes.data_as(oracle.i64p), na, ae, bb.ctypes.data_as(oracle.i64p), nb, be)
        a2, b2, e = vdecimal.calibrate_scale(a, ae, b, be)
        assert (a2.tolist(), b2.tolist(), e) == (aa.tolist(), bb.tolist(), e_exp), (a, b, ae, be)
