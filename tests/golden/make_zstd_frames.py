#!/usr/bin/env python3
"""Generate tests/golden/zstd_frames.json: zstd frames made by the reference's own libzstd 1.5.7 (through
oracle/_ref, exactly as gozstd.go:171 calls ZSTD_compressCCtx) on the codec's varint streams, plus the expected
decompressed bytes' sha256 and the raw stream.  Run in the build container only (needs oracle/_ref)."""
import base64
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def streams():
    rng = np.random.default_rng(20260921)
    for n in (130, 202, 1024, 8192):
        for sigma in (0.3, 1, 3, 10, 30):
            inc = 30000 + np.round(rng.normal(0, sigma, n)).astype(np.int64)
            b, _ = O.marshal_nearest_delta(np.cumsum(inc), 64, delta2=True)
            yield "delta2_n%d_s%g" % (n, sigma), b
        g = np.round(rng.normal(5000, 300, n)).astype(np.int64)
        b, _ = O.marshal_nearest_delta(g, 64, delta2=False)
        yield "gauge_n%d" % n, b
        u = np.cumsum(rng.integers(0, 1500, n))
        b, _ = O.marshal_nearest_delta(u, 64, delta2=True)
        yield "counter_u1500_n%d" % n, b
    # shapes that exercise RLE / repeat-offset / long-match paths
    yield "rle_like", np.array([7] * 5000 + [9] * 300 + [7] * 4000, dtype=np.uint8)
    yield "periodic", np.tile(np.arange(37, dtype=np.uint8), 400)
    yield "periodic_noise", (np.tile(np.arange(64, dtype=np.uint8), 200) ^ (rng.integers(0, 50, 12800) == 0).astype(np.uint8))
    yield "bigvarints", O.marshal_varint64s(rng.integers(-(1 << 62), 1 << 62, 3000))


def main():
    assert O.lib().vmo_zstd_ref_available(), "oracle/_ref/libzstd_ref.so missing: run `make -C oracle ref`"
    frames = []
    for name, raw in streams():
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        for level in ((-5, 1, 3, 5) if len(raw) < 3000 else (1, 5)):
            c = O.zstd_ref_compress(raw, level)
            frames.append({"name": name, "level": level, "raw_len": int(len(raw)),
                           "sha256": hashlib.sha256(raw.tobytes()).hexdigest(),
                           "frame": base64.b64encode(c.tobytes()).decode()})
    with open(os.path.join(HERE, "zstd_frames.json"), "w") as f:
        json.dump(frames, f, separators=(",", ":"))
    print("frames:", len(frames), "bytes:", os.path.getsize(os.path.join(HERE, "zstd_frames.json")))


if __name__ == "__main__":
    main()
