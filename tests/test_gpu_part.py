"""SURVEY.md 8(f).3 on the GPU: the compressed index files of a part go through the same zstd kernels as the block payloads.
  * vmb_zstd_decompress_batch == encoding.DecompressZSTD (compress.go:27) over frames written by the reference's libzstd
  * storage.Part.collect_blocks == metaindex.bin -> index blocks -> block headers -> descriptors over [timestamps.bin | values.bin],
    then the normal decode / rollup path: bit-exact against the oracle on the samples that were written."""
import ctypes as C

import numpy as np

from conftest import SEED0
import pytest

import blockgen
import partgen

pytestmark = pytest.mark.gpu


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


def _payloads(rng):
    hdr = lambda n: b"".join(partgen.pack_header(partgen.pack_tsid(1, 2, 3, 1000 + i // 3),
                                                 dict(min_ts=1_700_000_000_000 + i * 123_000, max_ts=1_700_000_000_000 + i * 123_000 + 122_000,
                                                      first_value=int(rng.integers(0, 10**9)), ts_off=i * 8, val_off=i * 9000,
                                                      ts_size=8, val_size=int(rng.integers(8000, 9000)), rows=8192, scale=-2, ts_mt=2,
                                                      val_mt=1, precision_bits=64)) for i in range(n))
    out = []
    for n in (1, 2, 7, 100, 808):                       # index-block shaped data (808 headers = 65448 B < maxBlockSize)
        out.append(np.frombuffer(hdr(n), dtype=np.uint8))
    for n in (1, 17, 255, 256, 4096, 65536, 131072, 163840):
        out.append(rng.integers(0, 256, n).astype(np.uint8))                         # incompressible: raw blocks
        out.append(rng.integers(0, 4, n).astype(np.uint8))                           # Huffman literals only
        out.append(np.full(n, 0x5A, dtype=np.uint8))                                 # RLE
        out.append(np.tile(rng.integers(0, 256, 37).astype(np.uint8), n // 37 + 1)[:n])  # long matches
        words = rng.integers(0, 256, (50, 6)).astype(np.uint8)
        out.append(words[rng.integers(0, 50, n // 6 + 1)].reshape(-1)[:n].copy())    # dictionary-like text: sequences + Huffman
    return out


def test_zstd_decompress_batch_matches_reference_libzstd(oracle):
    import victoriametrics_b200 as vm
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    rng = np.random.default_rng(SEED0 + 20)
    raws = _payloads(rng)
    frames = [oracle.zstd_ref_compress(r, int(lvl)) for r, lvl in zip(raws, rng.choice([1, 3, 5], len(raws)))]
    got = vm.encoding.decompress_zstd_batch(frames)
    assert len(got) == len(raws) >= 45
    for g, r in zip(got, raws):
        assert np.array_equal(g, r), len(r)
    # one frame at a time gives the same
    for k in (0, 4, len(raws) - 1):
        assert np.array_equal(vm.encoding.decompress_zstd_batch([frames[k]])[0], raws[k])


def test_zstd_decompress_batch_reports_bad_frames(oracle):
    import victoriametrics_b200 as vm
    from victoriametrics_b200 import _lib
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    rng = np.random.default_rng(SEED0 + 21)
    raws = [rng.integers(0, 7, 5000).astype(np.uint8) for _ in range(6)]
    frames = [oracle.zstd_ref_compress(r, 3) for r in raws]
    frames[1] = frames[1][: len(frames[1]) // 2].copy()        # truncated
    frames[3] = np.frombuffer(b"not a zstd frame at all....", dtype=np.uint8)
    big = oracle.zstd_ref_compress(np.zeros(163841, dtype=np.uint8), 1).copy()
    assert (big[4] >> 6) == 2                                    # Frame_Header_Descriptor: 4-byte Frame_Content_Size
    big[5:9] = np.frombuffer((1 << 28).to_bytes(4, "little"), dtype=np.uint8)  # declares 256 MiB: over the 128 MiB sanity cap
    frames[4] = big
    ctx = vm.default_context()
    offs = np.zeros(7, dtype=np.uint64)
    offs[1:] = np.cumsum([f.size for f in frames])
    arena = np.concatenate(frames)
    dst = np.zeros(64 * 1024, dtype=np.uint8)
    doffs, dlens, st = np.zeros(6, dtype=np.uint64), np.zeros(6, dtype=np.uint32), np.zeros(6, dtype=np.int32)
    rc = _lib.lib().vmb_zstd_decompress_batch(ctx.h, arena.ctypes.data_as(_lib.u8p), offs.ctypes.data_as(_lib.u64p), 6,
                                              dst.ctypes.data_as(_lib.u8p), dst.size, doffs.ctypes.data_as(_lib.u64p),
                                              dlens.ctypes.data_as(_lib.u32p), st.ctypes.data_as(_lib.i32p))
    assert rc == -6  # VMB_ERR_ZSTD (compress.go:31 returns the error)
    assert [int(s != 0) for s in st] == [0, 1, 0, 1, 1, 0], st
    for k in (0, 2, 5):  # the good frames of the batch are still delivered
        assert np.array_equal(dst[int(doffs[k]):int(doffs[k]) + int(dlens[k])], raws[k])
    with pytest.raises(_lib.VmbError):
        vm.encoding.decompress_zstd_batch(frames)


def test_zstd_decompress_batch_large_frames_and_content_checksums(oracle):
    """metaindex.bin has no size limit in the reference (metaindex_row.go:134): frames far above one block decompress (multi-block
    frames take the serial decoder); a frame that carries a Content_Checksum is verified (XXH64), a flipped content byte fails"""
    import victoriametrics_b200 as vm
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    rng = np.random.default_rng(SEED0 + 22)
    rows = rng.integers(0, 50, (20000, 56)).astype(np.uint8).reshape(-1)       # 1.1 MB, compressible, several zstd blocks
    noise = rng.integers(0, 256, 300000).astype(np.uint8)                      # raw blocks
    raws = [rows, noise, rng.integers(0, 3, 4000).astype(np.uint8)]
    frames = [oracle.zstd_ref_compress(r, 3) for r in raws]
    got = vm.encoding.decompress_zstd_batch(frames)
    for g, r in zip(got, raws):
        assert np.array_equal(g, r)
    ck = [oracle.zstd_ref_compress_checksum(r, 3) for r in raws]
    assert all((f[4] >> 2) & 1 for f in ck)                                     # Content_Checksum_Flag
    got = vm.encoding.decompress_zstd_batch(ck)
    for g, r in zip(got, raws):
        assert np.array_equal(g, r)
    bad = ck[1].copy()
    bad[len(bad) // 2] ^= 0x40                                                  # inside a Raw block: decodes, but to other content
    with pytest.raises(vm.VmbError):
        vm.encoding.decompress_zstd_batch([bad])
    bad2 = ck[0].copy()
    bad2[-1] ^= 1                                                               # the checksum itself
    with pytest.raises(vm.VmbError):
        vm.encoding.decompress_zstd_batch([bad2])


def _make_part(rng, nseries, max_index_block):
    t0 = 1_700_000_000_000
    series = []
    for s in range(nseries):
        tsid = partgen.pack_tsid(10 + s // 50, 1, s % 7, 5000 + s)
        nb = int(rng.choice([1, 1, 1, 2, 3]))
        blocks, cursor = [], t0 + int(rng.integers(0, 15)) * 1000
        for _ in range(nb):
            rows = int(rng.choice([1, 2, 40, 512, 2048, 8192]))
            kind = str(rng.choice(["regular", "jitter"]))
            ts = blockgen.gen_timestamps(rng, kind, rows, t0=cursor)
            vk = str(rng.choice(["counter", "counter_resets", "gauge", "const", "delta_const", "counter_smooth"]))
            blocks.append(blockgen.OBlock(ts, blockgen.gen_values(rng, vk, rows), int(rng.choice([-2, 0, 3])), 64, s))
            cursor = int(ts[-1]) + 15_000
        series.append((tsid, blocks))
    series.sort(key=lambda x: x[0])
    for i, (_, blocks) in enumerate(series):
        for b in blocks:
            b.series_idx = i
    return series, partgen.write_part(series, level=int(rng.choice([1, 3])), max_index_block=max_index_block)


@pytest.mark.parametrize("max_index_block", [partgen.MAX_BLOCK_SIZE, 81 * 5])
def test_part_collect_blocks_decode_and_rollup(oracle, max_index_block):
    import victoriametrics_b200 as vm
    from rollup_names import RF
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    rng = np.random.default_rng(SEED0 + 300 + max_index_block)
    series, files = _make_part(rng, 260, max_index_block)
    assert files["index_blocks"] >= (1 if max_index_block > 1000 else 50)
    part = vm.storage.Part(files["metaindex_bin"], files["index_bin"], files["timestamps_bin"], files["values_bin"])
    rows = part.metaindex_rows()
    assert len(rows) == files["index_blocks"] and rows.tobytes() is not None
    assert sum(int(r["block_headers_count"]) for r in rows) == len(files["headers"])
    descs, payload, tsids = part.collect_blocks()
    # every header field survives; offsets are rebased onto [timestamps.bin | values.bin]
    assert len(descs) == len(files["headers"]) and len(tsids) == len(series)
    nts = len(files["timestamps_bin"])
    for d, (tsid, h) in zip(descs, files["headers"]):
        for k in ("min_ts", "max_ts", "first_value", "ts_size", "val_size", "rows", "scale", "ts_mt", "val_mt", "precision_bits", "ts_off"):
            assert int(d[k]) == h[k], k
        assert int(d["val_off"]) == h["val_off"] + nts
    assert [t.tobytes() for t in tsids] == [s[0] for s in series]
    # decode: bit-exact against the oracle on what was written
    B = vm.storage.Blocks(descs, payload)
    ser, status = vm.storage.decode_blocks(B)
    assert not status.any()
    got = ser.to_lists()
    ser.close()
    assert len(got) == len(series)
    exp = []
    for s, (_, blocks) in enumerate(series):
        tss, vs = [], []
        for b in blocks:
            rc, ts, fv, _ = b.oracle_unmarshal()
            assert rc == 0
            tss.append(ts)
            vs.append(fv)
        ets, ev = np.concatenate(tss), np.concatenate(vs)
        assert np.array_equal(got[s][0], ets), s
        assert _same(got[s][1], ev), s
        exp.append((ets, ev))
    # and the query path on top: avg_over_time through the host pipeline
    t0 = 1_700_000_000_000
    start, end, step, window = t0 + 60_000, t0 + 3_000_000, 30_000, 300_000
    want = np.stack([oracle.rollup_do(RF["avg_over_time"], ev.copy(), ets.copy(), start, end, step, window)[0] for ets, ev in exp])
    out, _ = vm.promql.eval_rollup_func_host("avg_over_time", descs, payload, start, end, step, window)
    assert np.allclose(out, want, rtol=1e-12, atol=0, equal_nan=True)
    # a TSID + time-range filter keeps exactly the matching blocks (part_search.go:64)
    pick = [series[3][0], series[100][0]]
    d2, _, t2 = part.collect_blocks(tsids=pick, tr_min=t0, tr_max=t0 + 10**9)
    assert [t.tobytes() for t in t2] == pick and len(d2) == len(series[3][1]) + len(series[100][1])
    assert sorted(set(int(x) for x in d2["series_idx"])) == [0, 1]


def test_block_refs_feed_matches_collect_blocks(oracle):
    """vmb_blocks_upload_part: the tmpBlocksFile form of a query (marshaled headers + the part's two data files)"""
    import victoriametrics_b200 as vm
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    rng = np.random.default_rng(SEED0 + 777)
    series, files = _make_part(rng, 120, partgen.MAX_BLOCK_SIZE)
    # a query keeps some of the series: their headers, in series order, like netstorage.go groups BlockRefs by metric name
    keep = sorted(rng.choice(len(series), 70, replace=False).tolist())
    keep_tsids = {series[i][0] for i in keep}
    hdrs = b"".join(partgen.pack_header(tsid, h) for tsid, h in files["headers"] if tsid in keep_tsids)
    B = vm.storage.Blocks.from_block_refs(hdrs, files["timestamps_bin"], files["values_bin"])
    assert B.count == sum(len(series[i][1]) for i in keep)
    ser, status = vm.storage.decode_blocks(B)
    assert not status.any()
    got = ser.to_lists()
    ser.close()
    assert len(got) == len(keep)
    for g, i in zip(got, keep):
        tss, vs = [], []
        for b in series[i][1]:
            rc, ts, fv, _ = b.oracle_unmarshal()
            assert rc == 0
            tss.append(ts)
            vs.append(fv)
        assert np.array_equal(g[0], np.concatenate(tss)), i
        assert _same(g[1], np.concatenate(vs)), i
    # offsets outside the files / a damaged header are refused
    bad = bytearray(hdrs[:81])
    bad[56:64] = (len(files["values_bin"]) + 5).to_bytes(8, "big")  # ValuesBlockOffset
    with pytest.raises(vm.VmbError):
        vm.storage.Blocks.from_block_refs(bytes(bad), files["timestamps_bin"], files["values_bin"])
    with pytest.raises(ValueError):
        vm.storage.Blocks.from_block_refs(hdrs[:100], files["timestamps_bin"], files["values_bin"])


def test_zstd_decompress_empty_content_frame(oracle):
    import victoriametrics_b200 as vm
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    f = oracle.zstd_ref_compress(np.zeros(0, dtype=np.uint8), 3)
    one = oracle.zstd_ref_compress(np.array([9], dtype=np.uint8), 3)
    got = vm.encoding.decompress_zstd_batch([one, f, one])
    assert [g.tolist() for g in got] == [[9], [], [9]]


def _ll_table_log(frame):
    """accuracy log of the literal-lengths FSE table of a single-block frame with compressed literals (RFC 8878 3.1.1.3.2.1), or None"""
    b = bytes(frame)
    fhd = b[4]
    fcs, ss, did = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    pos = 5 + (0 if ss else 1) + [0, 1, 2, 4][did] + ([1, 2, 4, 8][fcs] if (fcs or ss) else 0)
    bh = b[pos] | (b[pos + 1] << 8) | (b[pos + 2] << 16)
    if not (bh & 1) or ((bh >> 1) & 3) != 2:
        return None  # not one compressed block
    blk = b[pos + 3: pos + 3 + (bh >> 3)]
    if (blk[0] & 3) != 2:
        return None
    sf = (blk[0] >> 2) & 3
    if sf in (0, 1):
        hs, comp = 3, ((blk[0] | (blk[1] << 8) | (blk[2] << 16)) >> 14) & 0x3ff
    elif sf == 2:
        hs, comp = 4, (int.from_bytes(blk[:4], "little") >> 18) & 0x3fff
    else:
        hs, comp = 5, (int.from_bytes(blk[:5], "little") >> 22) & 0x3ffff
    p = hs + comp
    b0 = blk[p]
    p += 1 if b0 < 128 else (2 if b0 < 255 else 3)
    if b0 == 0:
        return None
    modes = blk[p]
    return (blk[p + 1] & 15) + 5 if (modes >> 6) & 3 == 2 else None


def test_zstd_sequence_tables_above_256_states(oracle):
    """frames whose FSE tables have accuracy log 9 take the second k_zstd_seq_decode launch (full-size tables)"""
    import victoriametrics_b200 as vm
    if not oracle.lib().vmo_zstd_ref_available():
        pytest.skip("oracle/_ref (the reference's libzstd) was not built")
    rng = np.random.default_rng(SEED0 + 4321)
    srcs, frames, logs = [], [], []
    for trial in range(24):
        n = int(rng.choice([30000, 60000, 100000]))
        if trial % 2:
            a = np.repeat(rng.integers(0, 255, n // 7 + 1).astype(np.uint8), 7)[:n].copy()
            a[rng.random(n) < 0.05] = 1
        else:
            words = [bytes(rng.integers(97, 123, int(rng.integers(3, 12))).astype(np.uint8)) for _ in range(200)]
            a = np.frombuffer(b" ".join(words[int(i)] for i in rng.integers(0, 200, n // 6)), dtype=np.uint8)[:n].copy()
        f = oracle.zstd_ref_compress(a, int(rng.choice([1, 3])))
        srcs.append(a)
        frames.append(f)
        logs.append(_ll_table_log(f))
    assert any(l == 9 for l in logs), logs            # the shape this test is about
    small = [oracle.zstd_ref_compress(np.cumsum(rng.integers(0, 3, 8192)).astype(np.uint8), 1) for _ in range(8)]  # log <= 8 beside them
    got = vm.encoding.decompress_zstd_batch(frames + small)
    for g, a in zip(got[:len(srcs)], srcs):
        assert np.array_equal(g, a)
