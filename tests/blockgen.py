"""Test helpers: synthetic blocks marshaled by the ORACLE (i.e. byte streams as the reference would write them, zstd
frames by the reference's libzstd when oracle/_ref is present) and the oracle's view of the decoded result."""
import ctypes as C

import numpy as np

import oracle_lib as O


def gen_values(rng, kind, n):
    if kind == "const":
        return np.full(n, int(rng.integers(-5, 1000)), dtype=np.int64)
    if kind == "delta_const":
        return (int(rng.integers(-1000, 1000)) + int(rng.integers(-50, 50)) * np.arange(n)).astype(np.int64)
    if kind == "counter_smooth":   # tiny deltas -> zstd with sequences
        return np.cumsum(rng.integers(0, 3, n)).astype(np.int64)
    if kind == "counter":          # 2-byte varints -> zstd huffman-only
        return np.cumsum(rng.integers(0, 1500, n)).astype(np.int64)
    if kind == "counter_resets":
        inc = rng.integers(0, 1500, n)
        v = np.cumsum(inc)
        if n > 1:
            for r in rng.integers(1, n, max(n // 500, 1)):
                v[r:] -= v[r]
        return v.astype(np.int64)
    if kind == "counter_big":      # multi-byte varints, incompressible -> plain delta2
        return np.cumsum(rng.integers(0, 1 << 40, n)).astype(np.int64)
    if kind == "gauge":
        return np.round(rng.normal(5000, 300, n)).astype(np.int64)
    if kind == "gauge_wide":
        return rng.integers(-(1 << 62), 1 << 62, n).astype(np.int64)
    if kind == "gauge_small":
        return rng.integers(-3, 4, n).astype(np.int64)
    if kind == "special":          # decimal special values sprinkled in
        v = np.round(rng.normal(5000, 300, n)).astype(np.int64)
        idx = rng.integers(0, n, max(n // 50, 1))
        v[idx] = rng.choice([(1 << 63) - 1, -(1 << 63), (1 << 63) - 2, (1 << 63) - 3, -(1 << 63) + 1], len(idx))
        return v
    raise KeyError(kind)


def gen_timestamps(rng, kind, n, t0=1_700_000_000_000):
    if kind == "regular":
        return (t0 + 15000 * np.arange(n)).astype(np.int64)
    if kind == "jitter":
        return (t0 + 15000 * np.arange(n) + rng.integers(-50, 51, n)).astype(np.int64)
    if kind == "irregular":
        return (t0 + np.cumsum(rng.integers(1, 60000, n))).astype(np.int64)
    if kind == "dups":
        return (t0 + np.cumsum(rng.integers(0, 3, n)) * 1000).astype(np.int64)
    if kind == "single":
        return np.full(n, t0, dtype=np.int64)
    raise KeyError(kind)


VALUE_KINDS = ["const", "delta_const", "counter_smooth", "counter", "counter_resets", "counter_big", "gauge", "gauge_wide",
               "gauge_small", "special"]
TS_KINDS = ["regular", "jitter", "irregular", "dups"]


class OBlock:
    """a block marshaled by the oracle"""

    def __init__(self, ts, vals, scale, pb=64, series_idx=0):
        self.ts, self.vals, self.scale, self.pb, self.series_idx = ts, vals, scale, pb, series_idx
        self.vdata, self.vmt, self.first_value = O.marshal_int64_array(vals, pb)
        self.tdata, self.tmt, self.min_ts = O.marshal_int64_array(ts, pb)
        self.max_ts = int(ts[-1])
        self.rows = len(vals)

    def header(self):
        return dict(first_value=self.first_value, min_ts=self.min_ts, max_ts=self.max_ts, ts_size=self.tdata.size,
                    val_size=self.vdata.size, rows=self.rows, series_idx=self.series_idx, scale=self.scale,
                    ts_mt=self.tmt, val_mt=self.vmt, precision_bits=self.pb)

    def oracle_unmarshal(self, tr_min=-(1 << 63), tr_max=(1 << 63) - 1):
        """Block.UnmarshalData + AppendRowsWithTimeRangeFilter via the oracle -> (rc, ts, f64 values, int64 values)"""
        bh = O.BlockHeader()
        bh.min_ts, bh.max_ts, bh.first_value = self.min_ts, self.max_ts, self.first_value
        bh.ts_size, bh.val_size, bh.rows = self.tdata.size, self.vdata.size, self.rows
        bh.scale, bh.ts_mt, bh.val_mt, bh.precision_bits = self.scale, self.tmt, self.vmt, self.pb
        ts = np.empty(self.rows, dtype=np.int64)
        fv = np.empty(self.rows, dtype=np.float64)
        iv = np.empty(self.rows, dtype=np.int64)
        td = np.ascontiguousarray(self.tdata)
        vd = np.ascontiguousarray(self.vdata)
        n = O.lib().vmo_block_unmarshal(ts.ctypes.data_as(O.i64p), fv.ctypes.data_as(O.f64p), iv.ctypes.data_as(O.i64p),
                                        C.byref(bh), td.ctypes.data_as(O.u8p), vd.ctypes.data_as(O.u8p), tr_min, tr_max)
        if n < 0:
            return int(n), None, None, None
        return 0, ts[:n], fv[:n], iv[:n]


def random_blocks(rng, nblocks, rows_choices=(1, 2, 3, 31, 32, 33, 100, 511, 512, 513, 1000, 4096, 8191, 8192),
                  value_kinds=VALUE_KINDS, ts_kinds=TS_KINDS, pbs=(64,), scales=(-2, 0, 3, -9)):
    out = []
    for b in range(nblocks):
        n = int(rng.choice(rows_choices))
        vk = str(rng.choice(value_kinds))
        tk = str(rng.choice(ts_kinds))
        pb = int(rng.choice(pbs))
        out.append(OBlock(gen_timestamps(rng, tk, n), gen_values(rng, vk, n), int(rng.choice(scales)), pb, series_idx=b))
    return out


def to_blockset(blocks):
    from victoriametrics_b200 import storage
    bs = storage.BlockSet()
    for b in blocks:
        bs.add_marshaled(b.header(), b.tdata, b.vdata)
    return bs.finish()
