"""-m gpu: the chunked host pipeline (vmb_eval_rollup_host: H2D / kernels / D2H overlapped on three streams) must give
the same bits as the unchunked device path, for several chunk sizes, incl. shared timestamp payloads, gathered (non
contiguous) payload layouts, multi-block series and corrupt blocks."""
import os

import numpy as np

from conftest import SEED0
import pytest

import blockgen

pytestmark = pytest.mark.gpu
T0 = 1_700_000_000_000


def f64bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def vm():
    import victoriametrics_b200 as v
    v.default_context()
    return v


def _device_reference(vm, descs, payload, func, start, end, step, window, nseries):
    import torch
    B = vm.storage.Blocks(descs, payload)
    P = 1 + (end - start) // step
    out = torch.empty((nseries, P), dtype=torch.float64, device="cuda")
    _, scanned = vm.promql.eval_rollup_func(func, B, start, end, step, window, out_dev_ptr=out.data_ptr())
    torch.cuda.synchronize()
    return out.cpu().numpy(), scanned


@pytest.mark.parametrize("chunk", ["7", "64", "100000"])
def test_pipeline_chunks_match_device_path(vm, chunk):
    rng = np.random.default_rng(SEED0 + 21)
    blocks = blockgen.random_blocks(rng, 300, rows_choices=(2, 50, 700, 8192),
                                    value_kinds=["counter", "counter_resets", "gauge", "const", "delta_const", "counter_smooth", "counter_big"],
                                    ts_kinds=["regular", "jitter"], scales=(-2, 0))
    descs, payload = blockgen.to_blockset(blocks)
    start, end, step, window = T0 + 300000, T0 + 15000 * 1500, 15000, 300000
    os.environ["VMB_PIPE_CHUNK_BLOCKS"] = chunk
    try:
        for func in ("rate", "avg_over_time"):
            exp, escanned = _device_reference(vm, descs, payload, func, start, end, step, window, len(blocks))
            got, scanned = vm.promql.eval_rollup_func_host(func, descs, payload, start, end, step, window)
            assert np.array_equal(f64bits(got), f64bits(exp)), (func, chunk)
            assert scanned == escanned
    finally:
        del os.environ["VMB_PIPE_CHUNK_BLOCKS"]


def test_pipeline_multiblock_series_and_scattered_payload(vm, oracle):
    """series of 3 consecutive blocks (concatenated like mergeSortBlocks' fast path) + a payload arena whose blocks are
    stored in reverse order (forces the host-side gather path)"""
    from rollup_names import RF
    rng = np.random.default_rng(SEED0 + 22)
    nser, per = 40, 3
    bs = vm.storage.BlockSet()
    series_ts, series_v = [], []
    raw_blocks = []
    for s in range(nser):
        n = per * 1000
        ts = T0 + 15000 * np.arange(n) + rng.integers(-40, 40, n)
        v = np.cumsum(rng.integers(0, 1500, n)).astype(np.int64)
        series_ts.append(ts.astype(np.int64))
        series_v.append(v)
        for k in range(per):
            sl = slice(k * 1000, (k + 1) * 1000)
            blk = vm.storage.Block(ts[sl], v[sl], scale=-2, series_idx=s)
            raw_blocks.append(blk.marshal_data())
    # normal layout
    for h, t, v in raw_blocks:
        bs.add_marshaled(h, t, v)
    descs, payload = bs.finish()
    start, end, step, window = T0 + 300000, T0 + 15000 * 2900, 30000, 300000
    os.environ["VMB_PIPE_CHUNK_BLOCKS"] = "16"
    try:
        got, _ = vm.promql.eval_rollup_func_host("increase", descs, payload, start, end, step, window, nseries=nser)
        # scattered layout: same blocks, payloads placed back to front
        pieces, pos, offs = [], 0, []
        for h, t, v in reversed(raw_blocks):
            offs.append((pos, pos + t.size))
            pieces += [t, v]
            pos += t.size + v.size
        offs.reverse()
        d2 = (vm._lib.BlockDesc * len(raw_blocks))()
        for i, ((h, t, v), (to, vo)) in enumerate(zip(raw_blocks, offs)):
            for k, val in h.items():
                setattr(d2[i], k, val)
            d2[i].ts_off, d2[i].val_off = to, vo
        got2, _ = vm.promql.eval_rollup_func_host("increase", d2, np.concatenate(pieces), start, end, step, window, nseries=nser)
    finally:
        del os.environ["VMB_PIPE_CHUNK_BLOCKS"]
    assert np.array_equal(f64bits(got), f64bits(got2))
    rc = vm.promql.get_rollup_configs("increase", start, end, step, window)
    for s in (0, 7, nser - 1):
        fv = oracle.decimal_to_float(series_v[s], -2)
        oracle.lib().vmo_remove_counter_resets(fv.ctypes.data_as(oracle.f64p), series_ts[s].ctypes.data_as(oracle.i64p), len(fv), 0)
        exp, _ = oracle.rollup_do(RF["increase"], fv, series_ts[s], start, end, step, window,
                                  samples_scanned_per_call=rc.samplesScannedPerCall)
        assert np.allclose(got[s], exp, rtol=1e-12, atol=0, equal_nan=True), s


def test_pipeline_reports_corrupt_blocks(vm):
    rng = np.random.default_rng(SEED0 + 23)
    blocks = blockgen.random_blocks(rng, 40, rows_choices=(100, 1000), value_kinds=["counter", "gauge"], ts_kinds=["jitter"])
    blocks[5].vdata = blocks[5].vdata[:-3].copy()
    descs, payload = blockgen.to_blockset(blocks)
    os.environ["VMB_PIPE_CHUNK_BLOCKS"] = "8"
    try:
        with pytest.raises(vm.VmbError) as ei:
            vm.promql.eval_rollup_func_host("rate", descs, payload, T0, T0 + 10 ** 6, 15000, 60000)
        assert ei.value.code == -53
    finally:
        del os.environ["VMB_PIPE_CHUNK_BLOCKS"]
