"""CPU-side checks of the drop-in boundary: libvmb200.so loads and exports every symbol include/vmb200.h declares;
the layout of the C structs matches the ctypes mirrors; no compute is attempted without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vmb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vmb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from victoriametrics_b200 import _lib
    L = C.CDLL(_lib.SO_PATH)
    names = header_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    _lib.lib()  # also validates the ctypes signature table


def test_struct_layouts():
    from victoriametrics_b200 import _lib
    assert C.sizeof(_lib.BlockDesc) == 64
    assert _lib.BlockDesc.rows.offset == 48 and _lib.BlockDesc.scale.offset == 56 and _lib.BlockDesc.precision_bits.offset == 60
    assert C.sizeof(_lib.RollupCfg) == 80
    assert _lib.RollupCfg.args.offset == 64


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from victoriametrics_b200 import Context, VmbError
    with pytest.raises(VmbError) as ei:
        Context(0)
    assert ei.value.code == -51 and "no CPU fallback" in str(ei.value)


def test_block_header_parse_matches_oracle(oracle):
    from victoriametrics_b200 import _lib
    bh = oracle.BlockHeader()
    for i in range(24):
        bh.tsid[i] = 200 - i
    bh.min_ts, bh.max_ts, bh.first_value = 1700000000000, 1700000123456, -987654321
    bh.ts_off, bh.val_off, bh.ts_size, bh.val_size, bh.rows = 1 << 40, 12345, 17, 13492, 8192
    bh.scale, bh.ts_mt, bh.val_mt, bh.precision_bits = -2, 2, 1, 64
    buf = (C.c_uint8 * 81)()
    oracle.lib().vmo_block_header_marshal(buf, C.byref(bh))
    d = _lib.BlockDesc()
    tsid = (C.c_uint8 * 24)()
    assert _lib.lib().vmb_block_desc_from_header(C.byref(d), buf, tsid) == 0
    assert bytes(tsid) == bytes(bh.tsid)
    for f in ("min_ts", "max_ts", "first_value", "ts_off", "val_off", "ts_size", "val_size", "rows", "scale", "ts_mt",
              "val_mt", "precision_bits"):
        assert getattr(d, f) == getattr(bh, f), f
    bh.rows = 0
    oracle.lib().vmo_block_header_marshal(buf, C.byref(bh))
    assert _lib.lib().vmb_block_desc_from_header(C.byref(d), buf, None) == -10


def test_rollup_points_and_name_tables():
    from victoriametrics_b200 import _lib, promql
    import rollup_names
    cfg = _lib.RollupCfg(1, 0, 1000, 9000, 1000, 0, 0, 0, 0, 0, None, None)
    assert _lib.lib().vmb_rollup_points(C.byref(cfg)) == 9
    assert promql._RF_ORDER == rollup_names.RF_IDS  # product enum order == oracle enum order (tests rely on it)
    assert promql.get_timestamps(5, 26, 10).tolist() == [5, 15, 25]
