"""Pins the oracle's from-spec zstd decoder against the reference's own libzstd 1.5.7:
(a) committed golden frames (tests/golden/zstd_frames.json, made by make_zstd_frames.py);
(b) live differential test through oracle/_ref when it is present."""
import base64
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def frames():
    with open(os.path.join(HERE, "golden", "zstd_frames.json")) as f:
        return json.load(f)


def test_golden_frames_decode(frames, oracle):
    assert len(frames) > 60
    kinds = set()
    for fr in frames:
        c = np.frombuffer(base64.b64decode(fr["frame"]), dtype=np.uint8)
        rc, d = oracle.zstd_decompress(c)
        assert rc == 0, fr["name"]
        assert len(d) == fr["raw_len"] and hashlib.sha256(d.tobytes()).hexdigest() == fr["sha256"], (fr["name"], fr["level"])
        kinds.add(c[4])  # frame header descriptor byte
    assert len(kinds) >= 2  # 1-byte and 2-byte content-size forms both seen


def test_corrupt_frames_are_errors(frames, oracle):
    c = np.frombuffer(base64.b64decode(frames[10]["frame"]), dtype=np.uint8).copy()
    bad_magic = c.copy()
    bad_magic[0] ^= 0xFF
    assert oracle.zstd_decompress(bad_magic)[0] < 0
    assert oracle.zstd_decompress(c[:len(c) // 2])[0] < 0
    assert oracle.zstd_decompress(c[:3])[0] < 0


def test_live_differential_vs_libzstd(oracle):
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref/libzstd_ref.so not built")
    rng = np.random.default_rng(7)
    n_checked = 0
    for trial in range(120):
        n = int(rng.integers(128, 20000))
        mode = trial % 6
        if mode == 0:
            raw = rng.integers(0, 256, n).astype(np.uint8)  # incompressible => Raw block
        elif mode == 1:
            raw = rng.integers(0, 4, n).astype(np.uint8)
        elif mode == 2:
            raw = np.repeat(rng.integers(0, 256, n // 50 + 1), 50)[:n].astype(np.uint8)
        elif mode == 3:
            raw = oracle.marshal_varint64s(np.round(rng.normal(0, 10 ** rng.uniform(-0.5, 4), n)).astype(np.int64))
        elif mode == 4:
            base = rng.integers(0, 256, 97).astype(np.uint8)
            raw = np.tile(base, n // 97 + 1)[:n].copy()
            raw[rng.integers(0, n, n // 100)] = 0
        else:
            raw = np.zeros(n, dtype=np.uint8)  # RLE block
        for level in (-5, 1, 2, 3, 4, 5):
            c = oracle.zstd_ref_compress(raw, level)
            rc, d = oracle.zstd_decompress(c)
            assert rc == 0 and np.array_equal(d, raw), (trial, mode, level, n)
            n_checked += 1
    assert n_checked == 720


def test_marshal_bytes_roundtrip_through_both_decoders(oracle):
    """config-1 style: marshal with libzstd, decode with the oracle decoder AND libzstd: identical int64s"""
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref/libzstd_ref.so not built")
    rng = np.random.default_rng(3)
    v = np.cumsum(rng.integers(0, 3, 8192)).astype(np.int64)  # smooth => ZSTD type 1 with sequences
    b, mt, first = oracle.marshal_int64_array(v)
    assert mt == 1
    rc, out = oracle.unmarshal_int64_array(b, mt, first, len(v))
    assert rc == 0 and np.array_equal(out, v)
    cs = oracle.lib().vmo_zstd_content_size(b.ctypes.data_as(oracle.u8p), len(b))
    ref = oracle.zstd_ref_decompress(b, cs)
    rc, mine = oracle.zstd_decompress(b)
    assert rc == 0 and np.array_equal(ref, mine)
