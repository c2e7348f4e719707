"""Host-side write path of the product (vmb_marshal_int64, vmb_zstd_compress, vmb_float_to_decimal) checked against the
oracle and against the reference's own libzstd: the frames our writer produces must decode in the reference."""
import numpy as np
import pytest

from victoriametrics_b200 import decimal as vdecimal
from victoriametrics_b200 import encoding


def _streams():
    rng = np.random.default_rng(11)
    for n in (128, 129, 200, 1023, 1024, 1025, 4096, 16382, 16383, 16384, 40000, 81910):
        yield rng.integers(0, 256, n).astype(np.uint8)                      # incompressible
        yield rng.integers(0, 3, n).astype(np.uint8)                        # few symbols
        yield (rng.integers(0, 24, n) | ((np.arange(n) & 1) << 7)).astype(np.uint8)  # 2-byte-varint like, symbols > 128
        yield np.full(n, 7, dtype=np.uint8)                                 # RLE
        yield np.minimum(rng.geometric(0.2, n), 255).astype(np.uint8)       # skewed => long codes (length limiting)
        yield np.minimum(rng.geometric(0.02, n) + 100, 255).astype(np.uint8)


def test_zstd_writer_frames_decode_with_oracle_and_libzstd(oracle):
    have_ref = bool(oracle.lib().vmo_zstd_ref_available())
    nframes = 0
    for raw in _streams():
        c = encoding.zstd_compress(raw)
        rc, d = oracle.zstd_decompress(c)
        assert rc == 0 and np.array_equal(d, raw), (len(raw), raw[:8])
        if have_ref:
            ref = oracle.zstd_ref_decompress(c, len(raw))
            assert np.array_equal(ref, raw), len(raw)
        nframes += 1
    assert nframes >= 60


def test_zstd_writer_compresses_noisy_varints(oracle):
    rng = np.random.default_rng(5)
    v = np.cumsum(rng.integers(0, 1500, 8192))
    raw, _ = oracle.marshal_nearest_delta(v, 64, delta2=True)
    c = encoding.zstd_compress(raw)
    assert len(c) < 0.9 * len(raw)  # => MarshalTypeZSTDNearestDelta2 like the reference (ratio 0.86)
    if oracle.lib().vmo_zstd_ref_available():
        ref = oracle.zstd_ref_compress(raw, 5)
        assert len(c) < 1.03 * len(ref)  # Huffman-only is what libzstd itself emits here


@pytest.mark.parametrize("pb", [64, 32, 8, 4, 1])
def test_marshal_matches_oracle_types_and_roundtrips(oracle, pb):
    rng = np.random.default_rng(pb)
    cases = [
        np.full(100, 7),                                                # const
        1000 + 15000 * np.arange(8192),                                 # delta const
        np.cumsum(30000 + np.round(rng.normal(0, 1000, 1024))),         # counter, plain delta2 expected at pb=64
        np.cumsum(rng.integers(0, 1500, 8192)),                         # counter -> zstd
        np.round(rng.normal(5000, 300, 8192)),                          # gauge -> zstd
        np.round(rng.normal(0, 1e6, 300)),                              # gauge with negatives
        np.array([1, 20, 234]), np.array([5]), np.array([1, 2]),
    ]
    for vals in cases:
        vals = vals.astype(np.int64)
        b, mt, first = encoding.marshal_values(vals, pb)
        ob, omt, ofirst = oracle.marshal_int64_array(vals, pb) if oracle.lib().vmo_zstd_ref_available() else (None, None, None)
        assert first == int(vals[0])
        if omt is not None:
            # same family always; identical bytes whenever no zstd is involved
            fam = {1: "d2", 5: "d2", 4: "d", 6: "d", 2: "dc", 3: "c"}
            assert fam[mt] == fam[omt], (mt, omt)
            if mt in (2, 3, 5, 6) and omt == mt:
                assert np.array_equal(b, ob)
        # the reference-side decoder (oracle) accepts our bytes and returns what the oracle's own encoding returns
        rc, out = oracle.unmarshal_int64_array(b, mt, first, len(vals))
        assert rc == 0
        if omt is not None:
            rc2, out2 = oracle.unmarshal_int64_array(ob, omt, ofirst, len(vals))
            assert rc2 == 0 and np.array_equal(out, out2)
        if pb == 64:
            assert np.array_equal(out, vals)


def test_float_to_decimal_kats(kats):
    from conftest import gofloat
    for fa, da_exp, e_exp in kats["append_float_to_decimal"]:
        da, e = vdecimal.append_float_to_decimal([gofloat(s) for s in fa])
        assert e == e_exp and da.tolist() == da_exp, fa


def test_encoder_payloads_are_as_small_as_the_reference_encoders(oracle):
    """MarshalValues through the library's writer (Huffman-only zstd frames) vs the reference algorithm with its own libzstd
    at the reference's levels (encoding.go:371): same MarshalType, and no more than 3 % larger on any value kind"""
    import blockgen
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    rng = np.random.default_rng(31)
    for kind in ("counter", "counter_smooth", "gauge", "gauge_small", "counter_resets", "counter_big"):
        ref = ours = 0
        for _ in range(6):
            v = blockgen.gen_values(rng, kind, 8192)
            od, omt, ofirst = oracle.marshal_int64_array(v, 64)
            pd, pmt, pfirst = encoding.marshal_values(v, 64)
            assert (pmt, pfirst) == (omt, ofirst), kind
            ref += od.size
            ours += pd.size
        assert ours <= 1.03 * ref, (kind, ours, ref)
