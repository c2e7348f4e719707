"""app/vmselect/promql mirror: rollupConfig.Do, getRollupConfigs, evalRollupFunc*, incremental aggregates.

Names follow the reference (rollup.go / eval.go / aggr_incremental.go); every compute call goes to libvmb200.
"""
import ctypes as C

import numpy as np

from . import _lib, decimal, storage
from ._lib import RollupCfg, check, lib

# enum vmb_rollup_func order (include/vmb200.h); keys are the MetricsQL names of rollup.go:24-108
_RF_ORDER = ["default_rollup", "rate", "delta", "avg_over_time", "min_over_time", "max_over_time", "sum_over_time",
             "count_over_time", "quantile_over_time", "first_over_time", "last_over_time", "range_over_time",
             "sum2_over_time", "stddev_over_time", "stdvar_over_time", "ideriv", "idelta", "deriv", "increase_pure",
             "changes", "changes_prometheus", "resets", "increases_over_time", "integrate", "lag", "lifetime",
             "scrape_interval", "tmin_over_time", "tmax_over_time", "tfirst_over_time", "tlast_over_time",
             "tlast_change_over_time", "mode_over_time", "mad_over_time", "outlier_iqr_over_time", "zscore_over_time",
             "ascent_over_time", "descent_over_time", "distinct_over_time", "geomean_over_time", "predict_linear",
             "holt_winters", "hoeffding_bound_lower", "hoeffding_bound_upper", "duration_over_time",
             "count_le_over_time", "count_gt_over_time", "count_eq_over_time", "count_ne_over_time",
             "share_le_over_time", "share_gt_over_time", "share_eq_over_time", "sum_le_over_time", "sum_gt_over_time",
             "sum_eq_over_time", "present_over_time", "absent_over_time", "stale_samples_over_time",
             "median_over_time", "rate_over_sum", "delta_prometheus", "rate_prometheus", "rollup_open", "rollup_close",
             "rollup_high", "rollup_low"]
ROLLUP_FUNCS = {n: i for i, n in enumerate(_RF_ORDER)}
for _a, _b in {"deriv_fast": "rate", "increase": "delta", "irate": "ideriv", "decreases_over_time": "resets",
               "timestamp": "tlast_over_time", "timestamp_with_name": "tlast_over_time",
               "increase_prometheus": "delta_prometheus", "iqr_over_time": "outlier_iqr_over_time"}.items():
    ROLLUP_FUNCS[_a] = ROLLUP_FUNCS[_b]

# rollup.go:199 rollupFuncsCanAdjustWindow
ROLLUP_FUNCS_CAN_ADJUST_WINDOW = {"default_rollup", "deriv", "deriv_fast", "ideriv", "irate", "rate", "rate_over_sum",
                                  "rollup", "rollup_candlestick", "rollup_deriv", "rollup_rate",
                                  "rollup_scrape_interval", "scrape_interval", "timestamp"}
# rollup.go:223 rollupFuncsRemoveCounterResets
ROLLUP_FUNCS_REMOVE_COUNTER_RESETS = {"increase", "increase_prometheus", "increase_pure", "irate", "rate",
                                      "rate_prometheus", "rollup_increase", "rollup_rate"}
# rollup.go:238 rollupFuncsSamplesScannedPerCall
ROLLUP_FUNCS_SAMPLES_SCANNED_PER_CALL = {
    "absent_over_time": 1, "count_over_time": 1, "default_rollup": 1, "delta": 2, "delta_prometheus": 2, "deriv_fast": 2,
    "first_over_time": 1, "idelta": 2, "ideriv": 2, "increase": 2, "increase_prometheus": 2, "increase_pure": 2,
    "irate": 2, "lag": 1, "last_over_time": 1, "lifetime": 2, "present_over_time": 1, "rate": 2, "rate_prometheus": 2,
    "scrape_interval": 2, "tfirst_over_time": 1, "timestamp": 1, "timestamp_with_name": 1, "tlast_over_time": 1}

RC_MAY_ADJUST_WINDOW, RC_IS_DEFAULT_ROLLUP, RC_REMOVE_COUNTER_RESETS, RC_DROP_STALE_NANS = 1, 2, 4, 8
RC_PRE = {None: 0, "delta": 16, "deriv": 32, "scrape_interval": 64}  # value preFuncs of the multi-output rollups
# rollup.go:147 rollupAggrFuncs: what aggr_over_time() accepts
ROLLUP_AGGR_FUNCS = {"absent_over_time", "ascent_over_time", "avg_over_time", "changes", "count_over_time",
                     "decreases_over_time", "default_rollup", "delta", "deriv", "deriv_fast", "descent_over_time",
                     "distinct_over_time", "first_over_time", "geomean_over_time", "idelta", "ideriv", "increase",
                     "increase_pure", "increases_over_time", "integrate", "irate", "iqr_over_time", "lag", "last_over_time",
                     "lifetime", "mad_over_time", "max_over_time", "median_over_time", "min_over_time", "mode_over_time",
                     "present_over_time", "range_over_time", "rate", "rate_over_sum", "resets", "scrape_interval",
                     "stale_samples_over_time", "stddev_over_time", "stdvar_over_time", "sum_over_time", "sum2_over_time",
                     "tfirst_over_time", "timestamp", "timestamp_with_name", "tlast_change_over_time", "tlast_over_time",
                     "tmax_over_time", "tmin_over_time", "zscore_over_time"}
AGGR_FUNCS = {"sum": 0, "min": 1, "max": 2, "avg": 3, "count": 4, "sum2": 5, "geomean": 6, "any": 7, "group": 8}


def get_timestamps(start, end, step):
    """eval.go:230 getTimestamps"""
    if step <= 0:
        raise ValueError("BUG: Step must be bigger than 0; got %d" % step)
    if start > end:
        raise ValueError("BUG: Start cannot exceed End; got %d vs %d" % (start, end))
    return start + step * np.arange(1 + (end - start) // step, dtype=np.int64)


class RollupConfig:
    """rollupConfig rollup.go:574.  Func is the MetricsQL function name."""

    def __init__(self, Func, Start, End, Step, Window=0, LookbackDelta=0, MayAdjustWindow=False, isDefaultRollup=False,
                 samplesScannedPerCall=0, args=None, args2=None, removeCounterResets=False, dropStaleNaNs=False,
                 minStalenessInterval=0, TagValue="", preFunc=None):
        self.Func, self.Start, self.End, self.Step, self.Window = Func, int(Start), int(End), int(Step), int(Window)
        self.LookbackDelta, self.MayAdjustWindow, self.isDefaultRollup = int(LookbackDelta), MayAdjustWindow, isDefaultRollup
        self.samplesScannedPerCall, self.args, self.args2 = samplesScannedPerCall, args, args2
        self.removeCounterResets, self.dropStaleNaNs = removeCounterResets, dropStaleNaNs
        self.minStalenessInterval = int(minStalenessInterval)
        self.TagValue, self.preFunc = TagValue, preFunc  # rollup.go:576 TagValue; preFunc in (None, "delta", "deriv", "scrape_interval")
        if self.Step <= 0 or self.Start > self.End or self.Window < 0:  # rollup.go:703-711 logger.Panicf("BUG: ...")
            raise ValueError("BUG: invalid rollupConfig: Step=%d Start=%d End=%d Window=%d" % (Step, Start, End, Window))
        self.Timestamps = get_timestamps(self.Start, self.End, self.Step)
        self._keep = []

    @property
    def points(self):
        return int(self.Timestamps.size)

    def _cfg(self):
        flags = (RC_MAY_ADJUST_WINDOW if self.MayAdjustWindow else 0) | (RC_IS_DEFAULT_ROLLUP if self.isDefaultRollup else 0) \
            | (RC_REMOVE_COUNTER_RESETS if self.removeCounterResets else 0) | (RC_DROP_STALE_NANS if self.dropStaleNaNs else 0) \
            | RC_PRE[self.preFunc]
        cfg = RollupCfg(ROLLUP_FUNCS[self.Func], flags, self.Start, self.End, self.Step, self.Window, self.LookbackDelta,
                        self.minStalenessInterval, self.samplesScannedPerCall, 0, None, None)
        self._keep = []
        for name, a in (("args", self.args), ("args2", self.args2)):
            if a is not None:
                arr = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.points,)))
                self._keep.append(arr)
                setattr(cfg, name, arr.ctypes.data_as(_lib.f64p))
        return cfg

    def do(self, values, timestamps, ctx=None):
        """rollupConfig.Do rollup.go:688 for ONE series (kept for compatibility/tests; the batched forms are the fast path)
        -> (dstValues np.float64[points], samplesScanned)"""
        out, scanned = self.do_many([timestamps], [values], ctx)
        return out[0], scanned

    def do_many(self, timestamps_list, values_list, ctx=None):
        s = storage.Series.from_host(timestamps_list, values_list, ctx)
        try:
            return self.do_series(s)
        finally:
            s.close()

    def do_series(self, series, out_dev_ptr=None):
        """rollup over a device batch -> ([nseries x points] np.float64 (or None when out_dev_ptr is given), samplesScanned)"""
        cfg = self._cfg()
        scanned = C.c_uint64(0)
        if out_dev_ptr is not None:
            check(lib().vmb_rollup(series.ctx.h, series.h, C.byref(cfg), C.c_void_p(int(out_dev_ptr)), 1, C.byref(scanned)))
            return None, scanned.value
        out = np.empty((series.count, self.points), dtype=np.float64)
        check(lib().vmb_rollup(series.ctx.h, series.h, C.byref(cfg), C.c_void_p(out.ctypes.data), 0, C.byref(scanned)))
        return out, scanned.value


def get_rollup_configs(func_name, start, end, step, window=0, lookback_delta=0, args=None, args2=None,
                       no_stale_markers=False, min_staleness_interval=0):
    """getRollupConfigs rollup.go:374 for the single-config functions + the preFunc / dropStaleNaNs decisions of
    eval.go:1855-1866, :1985.  (rollup*(), aggr_over_time(), *_values_over_time() produce several series per input
    and stay on the host: SURVEY.md 8(a) a23.)"""
    name = func_name.lower()
    if name not in ROLLUP_FUNCS:
        raise KeyError("unsupported rollup function %r" % func_name)
    drop_stale = not (no_stale_markers or name in ("default_rollup", "stale_samples_over_time"))
    return RollupConfig(name, start, end, step, window, lookback_delta,
                        MayAdjustWindow=name in ROLLUP_FUNCS_CAN_ADJUST_WINDOW, isDefaultRollup=name == "default_rollup",
                        samplesScannedPerCall=ROLLUP_FUNCS_SAMPLES_SCANNED_PER_CALL.get(name, 0), args=args, args2=args2,
                        removeCounterResets=name in ROLLUP_FUNCS_REMOVE_COUNTER_RESETS, dropStaleNaNs=drop_stale,
                        minStalenessInterval=min_staleness_interval)


def eval_rollup_func(func_name, blocks, start, end, step, window=0, lookback_delta=0, args=None, args2=None,
                     tr_min=storage.INT64_MIN, tr_max=storage.INT64_MAX, out_dev_ptr=None, rc=None):
    """evalRollupFuncNoCache eval.go:1680 -> evalRollupNoIncrementalAggregate eval.go:1845 on device-resident blocks:
    decode, per-series preamble, rollupConfig.Do for every series.
    -> ([nseries x points] np.float64 or None, samplesScanned).  `rc`: a RollupConfig to use instead of the one
    getRollupConfigs derives from func_name."""
    rc = rc or get_rollup_configs(func_name, start, end, step, window, lookback_delta, args, args2)
    cfg = rc._cfg()
    scanned = C.c_uint64(0)
    if out_dev_ptr is not None:
        check(lib().vmb_eval_rollup_device(blocks.ctx.h, blocks.h, tr_min, tr_max, C.byref(cfg), C.c_void_p(int(out_dev_ptr)),
                                           C.byref(scanned)))
        return None, scanned.value
    series, _ = storage.decode_blocks(blocks, tr_min, tr_max)
    try:
        return rc.do_series(series)[0], None
    finally:
        series.close()


def eval_rollup_func_with_subquery(func_name, inner_dev_ptr, nseries, sq_start, sq_end, sq_step, start, end, step, window,
                                   lookback_delta=0, args=None, args2=None, out_dev_ptr=None, ctx=None):
    """evalRollupFuncWithSubquery eval.go:910 with the inner result resident on the device: `inner_dev_ptr` = [nseries x Psq] float64
    on the subquery grid sq_start..sq_end step sq_step (already aligned by the caller, eval.go:932); every row loses its NaN points
    (removeNanValues), goes through the outer function's preFunc and rollupConfig.Do on the outer grid.
    -> ([nseries x points] np.float64 or None, samplesScanned)"""
    rc = get_rollup_configs(func_name, start, end, step, window, lookback_delta, args, args2)
    rc.dropStaleNaNs = False  # the subquery path has no dropStaleNaNs (eval.go:958-964)
    psq = 1 + (sq_end - sq_start) // sq_step
    series = storage.Series.from_matrix(inner_dev_ptr, nseries, psq, sq_start, sq_step, ctx)
    try:
        return rc.do_series(series, out_dev_ptr)
    finally:
        series.close()


def eval_rollup_func_host(func_name, descs, payload, start, end, step, window=0, lookback_delta=0, args=None, args2=None,
                          tr_min=storage.INT64_MIN, tr_max=storage.INT64_MAX, out=None, nseries=None, ctx=None, rc=None):
    """the whole path with HOST buffers in one call (vmb_eval_rollup_host): H2D, decode, rollup, D2H."""
    ctx = ctx or _lib.default_context()
    rc = rc or get_rollup_configs(func_name, start, end, step, window, lookback_delta, args, args2)
    cfg = rc._cfg()
    if isinstance(descs, np.ndarray):
        dptr, n = descs.ctypes.data_as(C.POINTER(_lib.BlockDesc)), descs.shape[0]
        if nseries is None:
            nseries = int(np.count_nonzero(np.diff(descs["series_idx"])) + 1) if n else 0
    else:
        dptr, n = descs, len(descs)
        if nseries is None:
            nseries = len({d.series_idx for d in descs})
    if out is None:
        out = np.empty((nseries, rc.points), dtype=np.float64)
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    scanned = C.c_uint64(0)
    check(lib().vmb_eval_rollup_host(ctx.h, dptr, n, payload.ctypes.data_as(_lib.u8p), payload.size, tr_min, tr_max,
                                     C.byref(cfg), out.ctypes.data_as(_lib.f64p), None, C.byref(scanned)))
    return out, scanned.value


def eval_rollup_aggr_host(aggr_name, func_name, descs, payload, group_ids, ngroups, start, end, step, window=0,
                          lookback_delta=0, args=None, args2=None, tr_min=storage.INT64_MIN, tr_max=storage.INT64_MAX,
                          out=None, ctx=None, rc=None):
    """aggr(rollup(m[d])) by (...) with HOST buffers in one call (vmb_eval_rollup_aggr_host): H2D, decode, rollup and the
    incremental aggregate on the GPU, D2H of the [ngroups x points] result only -> (np.float64[ngroups, points], samplesScanned)"""
    ctx = ctx or _lib.default_context()
    rc = rc or get_rollup_configs(func_name, start, end, step, window, lookback_delta, args, args2)
    cfg = rc._cfg()
    if isinstance(descs, np.ndarray):
        dptr, n = descs.ctypes.data_as(C.POINTER(_lib.BlockDesc)), descs.shape[0]
    else:
        dptr, n = descs, len(descs)
    if out is None:
        out = np.empty((int(ngroups), rc.points), dtype=np.float64)
    g = np.ascontiguousarray(group_ids, dtype=np.uint32)
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    scanned = C.c_uint64(0)
    check(lib().vmb_eval_rollup_aggr_host(ctx.h, dptr, n, payload.ctypes.data_as(_lib.u8p), payload.size, tr_min, tr_max,
                                          C.byref(cfg), AGGR_FUNCS[aggr_name.lower()], g.ctypes.data_as(_lib.u32p), int(ngroups),
                                          out.ctypes.data_as(_lib.f64p), None, C.byref(scanned)))
    return out, scanned.value


def eval_rollup_aggr_dist(aggr_name, func_name, blocks, group_ids, ngroups, start, end, step, window=0, lookback_delta=0, args=None,
                          args2=None, tr_min=storage.INT64_MIN, tr_max=storage.INT64_MAX, out=None, rc=None):
    """aggr(rollup(m[d])) by (...) over every rank of the ctx's communicator in ONE library call (vmb_eval_rollup_aggr_dist): this
    rank's device-resident blocks are folded on the GPU, the partial states are merged by the library's NCCL all-reduce, every rank
    finalizes -> (np.float64[ngroups, points], this rank's samplesScanned).  Without a communicator: the single-GPU result."""
    rc = rc or get_rollup_configs(func_name, start, end, step, window, lookback_delta, args, args2)
    cfg = rc._cfg()
    if out is None:
        out = np.empty((int(ngroups), rc.points), dtype=np.float64)
    g = np.ascontiguousarray(group_ids, dtype=np.uint32)
    scanned = C.c_uint64(0)
    check(lib().vmb_eval_rollup_aggr_dist(blocks.ctx.h, blocks.h, tr_min, tr_max, C.byref(cfg), AGGR_FUNCS[aggr_name.lower()],
                                          g.ctypes.data_as(_lib.u32p), int(ngroups), out.ctypes.data_as(_lib.f64p), C.byref(scanned)))
    return out, scanned.value


class IncrementalAggr:
    """incrementalAggrFuncContext aggr_incremental.go:73: aggr(rollup(m[d])) by (...) without keeping [series x points]
    on the host.  update() == updateTimeseries for every series of a device batch (per-GPU partial state);
    finalize() == finalizeTimeseries.  With torch.distributed initialised, finalize(all_reduce=True) merges the per-rank
    partial states with one NCCL all-reduce of values and one of counts."""

    def __init__(self, aggr_name, ngroups, points, device_alloc):
        """device_alloc(nbytes) -> object with .ptr (device address); e.g. a torch.empty(..., device='cuda') wrapper"""
        self.aggr = AGGR_FUNCS[aggr_name.lower()]
        self.name = aggr_name.lower()
        self.ngroups, self.points = int(ngroups), int(points)
        self.values = device_alloc(self.ngroups * self.points * 8)
        self.counts = device_alloc(self.ngroups * self.points * 8)

    def update(self, series, rc, group_ids, rolled_scratch_ptr=None):
        g = np.ascontiguousarray(group_ids, dtype=np.uint32)
        cfg = rc._cfg()
        scanned = C.c_uint64(0)
        check(lib().vmb_rollup_aggr_partial(series.ctx.h, series.h, C.byref(cfg), self.aggr, g.ctypes.data_as(_lib.u32p),
                                            self.ngroups, C.c_void_p(self.values.ptr), C.c_void_p(self.counts.ptr),
                                            C.c_void_p(rolled_scratch_ptr or 0), C.byref(scanned)))
        return scanned.value

    def update_blocks(self, blocks, rc, group_ids, tr_min=storage.INT64_MIN, tr_max=storage.INT64_MAX):
        """decode + preamble + rollup + fold of device-resident compressed blocks in one library call
        (vmb_eval_rollup_aggr_device); the decoded columns stay in a library-side cache"""
        g = np.ascontiguousarray(group_ids, dtype=np.uint32)
        cfg = rc._cfg()
        scanned = C.c_uint64(0)
        check(lib().vmb_eval_rollup_aggr_device(blocks.ctx.h, blocks.h, tr_min, tr_max, C.byref(cfg), self.aggr,
                                                g.ctypes.data_as(_lib.u32p), self.ngroups, C.c_void_p(self.values.ptr),
                                                C.c_void_p(self.counts.ptr), C.byref(scanned)))
        return scanned.value

    def update_host(self, descs, payload, rc, group_ids, ctx, tr_min=storage.INT64_MIN, tr_max=storage.INT64_MAX):
        """the same from HOST buffers through the chunked pipeline (vmb_eval_rollup_aggr_host_partial): the partial state of
        the batch replaces this object's state"""
        g = np.ascontiguousarray(group_ids, dtype=np.uint32)
        cfg = rc._cfg()
        scanned = C.c_uint64(0)
        if isinstance(descs, np.ndarray):
            dptr, n = descs.ctypes.data_as(C.POINTER(_lib.BlockDesc)), descs.shape[0]
        else:
            dptr, n = descs, len(descs)
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        check(lib().vmb_eval_rollup_aggr_host_partial(ctx.h, dptr, n, payload.ctypes.data_as(_lib.u8p), payload.size, tr_min,
                                                      tr_max, C.byref(cfg), self.aggr, g.ctypes.data_as(_lib.u32p), self.ngroups,
                                                      C.c_void_p(self.values.ptr), C.c_void_p(self.counts.ptr), None,
                                                      C.byref(scanned)))
        return scanned.value

    def finalize(self, ctx, all_reduce=None, out=None):
        """all_reduce(values_buf, counts_buf, op) is called between prepare and finalize when given; `out` may be a
        preallocated (ideally pinned, vmb_host_alloc) [ngroups x points] float64 array"""
        n = self.ngroups * self.points
        if all_reduce is not None:
            check(lib().vmb_aggr_prepare_allreduce(ctx.h, self.aggr, C.c_void_p(self.values.ptr), C.c_void_p(self.counts.ptr), n))
            ctx.synchronize()
            op = {"min": "min", "max": "max", "geomean": "prod"}.get(self.name, "sum")
            all_reduce(self.values, self.counts, op)
        if out is None:
            out = np.empty((self.ngroups, self.points), dtype=np.float64)
        assert out.dtype == np.float64 and out.size == n and out.flags.c_contiguous
        check(lib().vmb_aggr_finalize(ctx.h, self.aggr, C.c_void_p(self.values.ptr), C.c_void_p(self.counts.ptr), n,
                                      out.ctypes.data_as(_lib.f64p)))
        return out


# ---- multi-GPU protocol for aggr(rollup(...)) by (...)   -------------------------------------------------
# all-reduce operator and the identity vmb_aggr_prepare_allreduce writes into empty cells (count == 0), per aggregate
ALLREDUCE_OP = {"sum": "sum", "avg": "sum", "count": "sum", "sum2": "sum", "group": "sum", "min": "min", "max": "max",
                "geomean": "prod"}
ALLREDUCE_IDENTITY = {"sum": 0.0, "avg": 0.0, "count": 0.0, "sum2": 0.0, "group": 0.0, "min": float("inf"),
                      "max": float("-inf"), "geomean": 1.0}


def shard_series(nseries, rank, world):
    """series owned by `rank`: MetricID mod world (independent series => no exchange for decode + rollup)"""
    return np.arange(rank, nseries, world)


def dense_group_ids(group_keys):
    """dense group ids from the marshaled group-by label sets (aggr_incremental.go:113 marshalMetricNameSorted);
    every rank must call this on the SAME global key list so that ids agree across ranks -> (ids, ngroups)"""
    uniq = {}
    ids = np.empty(len(group_keys), dtype=np.uint32)
    for i, k in enumerate(group_keys):
        ids[i] = uniq.setdefault(k, len(uniq))
    return ids, len(uniq)


def torch_all_reduce(values_t, counts_t, op):
    """the `all_reduce` callback for IncrementalAggr.finalize: NCCL (or gloo) all-reduce of values with the aggregate's
    operator and of counts with sum"""
    import torch.distributed as dist
    ops = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "prod": dist.ReduceOp.PRODUCT}
    dist.all_reduce(values_t, op=ops[op])
    dist.all_reduce(counts_t, op=dist.ReduceOp.SUM)


# ---- multi-output rollups (getRollupConfigs rollup.go:416-504): one input series -> several output series --------------
def get_rollup_configs_multi(func_name, start, end, step, window=0, lookback_delta=0, tag=None, aggr_funcs=None, phis=None,
                             no_stale_markers=False, min_staleness_interval=0):
    """getRollupConfigs rollup.go:374 for rollup(), rollup_rate/deriv/increase/delta(), rollup_scrape_interval(),
    rollup_candlestick(), aggr_over_time() and quantiles_over_time(): -> [RollupConfig], one per output series, TagValue =
    the value of the `rollup` (or phi) label.  The value preFunc (deltaValues / derivValues / intervals) and
    removeCounterResets are shared by the configs and run once per decoded batch."""
    name = func_name.lower()
    rcr = name in ROLLUP_FUNCS_REMOVE_COUNTER_RESETS
    pre = None
    spc = ROLLUP_FUNCS_SAMPLES_SCANNED_PER_CALL.get(name, 0)
    if name in ("rollup", "rollup_rate", "rollup_deriv", "rollup_increase", "rollup_delta", "rollup_scrape_interval"):
        pre = {"rollup_rate": "deriv", "rollup_deriv": "deriv", "rollup_increase": "delta", "rollup_delta": "delta",
               "rollup_scrape_interval": "scrape_interval"}.get(name)
        funcs = {"min": "min_over_time", "max": "max_over_time", "avg": "avg_over_time"}
        if tag in (None, ""):
            pairs = [(funcs[t], t, None) for t in ("min", "max", "avg")]
        elif tag in funcs:
            pairs = [(funcs[tag], "", None)]
        else:
            raise ValueError("unexpected second arg for %s: %r; want `min`, `max` or `avg`" % (func_name, tag))
    elif name == "rollup_candlestick":
        funcs = {"open": "rollup_open", "close": "rollup_close", "low": "rollup_low", "high": "rollup_high"}
        if tag in (None, ""):
            pairs = [(funcs[t], t, None) for t in ("open", "close", "low", "high")]
        elif tag in funcs:
            pairs = [(funcs[tag], tag, None)]
        else:
            raise ValueError("unexpected second arg for %s: %r; want `open`, `close`, `low` or `high`" % (func_name, tag))
    elif name == "aggr_over_time":
        if not aggr_funcs:
            raise ValueError("aggr_over_time() needs at least one aggregate function name")
        pairs = []
        for f in aggr_funcs:
            f = f.lower()
            if f not in ROLLUP_AGGR_FUNCS:
                raise ValueError("%r cannot be used in `aggr_over_time` function" % f)
            rcr = rcr or f in ROLLUP_FUNCS_REMOVE_COUNTER_RESETS
            pairs.append((f, f, None))
    elif name == "quantiles_over_time":
        if not phis:
            raise ValueError("quantiles_over_time() needs at least one phi")
        pairs = [("quantile_over_time", repr(float(p)).rstrip("0").rstrip(".") if float(p) != int(p) else str(int(p)), float(p))
                 for p in phis]
    else:
        raise KeyError("%r is not a multi-output rollup function" % func_name)
    drop_stale = not no_stale_markers
    return [RollupConfig(f, start, end, step, window, lookback_delta,
                         MayAdjustWindow=name in ROLLUP_FUNCS_CAN_ADJUST_WINDOW, isDefaultRollup=False,
                         samplesScannedPerCall=spc, args=arg, removeCounterResets=rcr, dropStaleNaNs=drop_stale,
                         minStalenessInterval=min_staleness_interval, TagValue=t, preFunc=pre) for f, t, arg in pairs]


def eval_rollup_func_multi(func_name, blocks, start, end, step, window=0, lookback_delta=0, tr_min=storage.INT64_MIN,
                           tr_max=storage.INT64_MAX, **kw):
    """evalRollupNoIncrementalAggregate eval.go:1845 for a multi-output rollup on device-resident blocks: decode once, run
    the shared preFunc once, then one rollup per config -> ({TagValue: [nseries x points] np.float64}, samplesScanned)"""
    rcs = get_rollup_configs_multi(func_name, start, end, step, window, lookback_delta, **kw)
    series, _ = storage.decode_blocks(blocks, tr_min, tr_max)
    try:
        out, scanned = {}, 0
        for rc in rcs:
            m, sc = rc.do_series(series)
            out[rc.TagValue] = m
            scanned += sc
        return out, scanned
    finally:
        series.close()


# ---- topk / bottomk over the [series x points] matrix (aggr.go:646 newAggrFuncTopK) -------------------------------------
def topk(ks, vals_dev_ptr, nseries, points, device_alloc, group_ids=None, ngroups=1, reverse=False, ctx=None,
         all_gather=None, group_sizes=None, series_id_base=0):
    """topk(k, q) (reverse=True: bottomk) on a DEVICE matrix [nseries x points] of float64, masked in place: per group and
    point only the k best values survive (fillNaNsAtIdx aggr.go:786).  ks: scalar or one k per point.
    device_alloc(nbytes) -> object with .ptr.  Several processes (series sharded by rank): all_gather(buf, nbytes) ->
    (gathered_buf, nparts) over the candidate lists (all_gather="nccl": the library's own ncclAllGather over the ctx's communicator),
    group_sizes = series per group over ALL processes, series_id_base = global id of this process's row 0 (equal values rank by
    ascending global series id: exactly k survive per group and point).
    -> np.bool_[nseries]: rows that still hold a value (removeEmptySeries drops the others)"""
    ctx = ctx or _lib.default_context()
    g = np.zeros(nseries, dtype=np.uint32) if group_ids is None else np.ascontiguousarray(group_ids, dtype=np.uint32)
    if group_sizes is None:
        group_sizes = np.bincount(g, minlength=ngroups)
    gs = np.ascontiguousarray(group_sizes, dtype=np.uint32)
    kk = np.ascontiguousarray(np.broadcast_to(np.asarray(ks, dtype=np.float64), (points,)))
    kclean = np.where(np.isnan(kk) | (kk < 0), 0.0, kk)
    kmax = int(min(np.floor(kclean.max()) if points else 0, gs.max() if ngroups else 0))
    kmax = max(kmax, 1)
    if kmax > 64:
        raise ValueError("topk on the GPU supports k <= 64 (got %d)" % kmax)
    cells = int(ngroups) * int(points)
    cand = device_alloc(cells * kmax * 16)  # {value, global series id} per entry
    rev = 1 if reverse else 0
    check(lib().vmb_topk_candidates(ctx.h, C.c_void_p(int(vals_dev_ptr)), nseries, points, g.ctypes.data_as(_lib.u32p), int(ngroups),
                                    kmax, rev, int(series_id_base), C.c_void_p(cand.ptr)))
    if all_gather == "nccl":
        nparts = ctx.comm_size
        parts = device_alloc(nparts * cells * kmax * 16)
        check(lib().vmb_topk_allgather(ctx.h, C.c_void_p(cand.ptr), cells * kmax * 2, C.c_void_p(parts.ptr)))
        check(lib().vmb_topk_merge(ctx.h, C.c_void_p(parts.ptr), int(nparts), cells, kmax, rev, C.c_void_p(cand.ptr)))
    elif all_gather is not None:
        ctx.synchronize()
        parts, nparts = all_gather(cand, cells * kmax * 16)
        check(lib().vmb_topk_merge(ctx.h, C.c_void_p(parts.ptr), int(nparts), cells, kmax, rev, C.c_void_p(cand.ptr)))
    flags = np.zeros(max(nseries, 1), dtype=np.uint8)
    check(lib().vmb_topk_apply(ctx.h, C.c_void_p(int(vals_dev_ptr)), nseries, points, g.ctypes.data_as(_lib.u32p), int(ngroups),
                               gs.ctypes.data_as(_lib.u32p), C.c_void_p(cand.ptr), kmax, kk.ctypes.data_as(_lib.f64p), rev,
                               int(series_id_base), flags.ctypes.data_as(_lib.u8p)))
    return flags[:nseries].astype(bool)


# ---- post-rollup operations on device matrices (binary_op.go, aggr.go:1217 quantile, rollup_result_cache.go:618 mergeSeries) ----------
BINARY_OPS = {"+": 0, "-": 1, "*": 2, "/": 3, "%": 4, "^": 5, "atan2": 6, "==": 7, "!=": 8, ">": 9, "<": 10, ">=": 11, "<=": 12,
              "default": 13, "if": 14, "ifnot": 15}


def binary_op(op, left_dev_ptr, right_dev_ptr, npairs, points, dst_dev_ptr, left_rows=None, right_rows=None, is_bool=False, ctx=None):
    """newBinaryOpFunc binary_op.go:155: dst[i] = left[left_rows[i]] op right[right_rows[i]] element by element on DEVICE matrices;
    the row lists come from the host's tag matching (adjustBinaryOpTags), a scalar operand is a one-row matrix with rows all 0"""
    ctx = ctx or _lib.default_context()
    lr = None if left_rows is None else np.ascontiguousarray(left_rows, dtype=np.uint32)
    rr = None if right_rows is None else np.ascontiguousarray(right_rows, dtype=np.uint32)
    check(lib().vmb_binary_op(ctx.h, BINARY_OPS[op.lower()], int(bool(is_bool)), C.c_void_p(int(left_dev_ptr)),
                              lr.ctypes.data_as(_lib.u32p) if lr is not None else None, C.c_void_p(int(right_dev_ptr)),
                              rr.ctypes.data_as(_lib.u32p) if rr is not None else None, int(npairs), int(points), C.c_void_p(int(dst_dev_ptr))))


def merge_series(a_dev_ptr, a_rows, pa, b_dev_ptr, b_rows, pb, dst_dev_ptr, ctx=None):
    """mergeSeries rollup_result_cache.go:618 on DEVICE matrices: dst row i = a[a_rows[i]] ++ b[b_rows[i]], -1 = series missing (NaNs)"""
    ctx = ctx or _lib.default_context()
    ar = np.ascontiguousarray(a_rows, dtype=np.int64)
    br = np.ascontiguousarray(b_rows, dtype=np.int64)
    assert ar.size == br.size
    check(lib().vmb_matrix_merge_rows(ctx.h, C.c_void_p(int(a_dev_ptr)), ar.ctypes.data_as(_lib.i64p), int(pa), C.c_void_p(int(b_dev_ptr)),
                                      br.ctypes.data_as(_lib.i64p), int(pb), ar.size, C.c_void_p(int(dst_dev_ptr))))


def group_first_value(vals_dev_ptr, nseries, points, group_ids, ngroups, out_dev_ptr, ctx=None):
    """the right-hand side of a set operator reduced per tag-set group (vmb_group_first_value): out[g][j] = first non-NaN value of
    the group's rows at point j, in row order"""
    ctx = ctx or _lib.default_context()
    g = np.ascontiguousarray(group_ids, dtype=np.uint32)
    check(lib().vmb_group_first_value(ctx.h, C.c_void_p(int(vals_dev_ptr)), int(nseries), int(points), g.ctypes.data_as(_lib.u32p), int(ngroups),
                                      C.c_void_p(int(out_dev_ptr))))


def set_op(op, left_dev_ptr, left_groups, nleft, right_dev_ptr, right_groups, nright, ngroups, points, dst_dev_ptr, tmp_dev_ptr, ctx=None):
    """`and` (binaryOpAnd binary_op.go:430), `unless` (:610), `if` (:416), `ifnot` (:595), `default` (:463) between two DEVICE matrices
    whose rows the host has keyed by tag set (left_groups / right_groups: dense key ids < ngroups, every left key present on the
    right): the right side is reduced per key into tmp_dev_ptr [ngroups x points], then one element pass over the left rows"""
    el = {"and": "if", "if": "if", "unless": "ifnot", "ifnot": "ifnot", "default": "default"}[op.lower()]
    group_first_value(right_dev_ptr, nright, points, right_groups, ngroups, tmp_dev_ptr, ctx=ctx)
    binary_op(el, left_dev_ptr, tmp_dev_ptr, nleft, points, dst_dev_ptr, right_rows=np.asarray(left_groups, dtype=np.uint32), ctx=ctx)


TRANSFORM_FUNCS = {n: i for i, n in enumerate(
    ["abs", "ceil", "floor", "sqrt", "exp", "ln", "log2", "log10", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh",
     "acosh", "atanh", "deg", "rad", "sgn", "clamp", "clamp_min", "clamp_max", "round"])}
TRANSFORM_FUNCS.update({n: 32 + i for i, n in enumerate(
    ["running_sum", "running_min", "running_max", "running_avg", "range_sum", "range_min", "range_max", "range_avg", "range_first",
     "range_last", "keep_last_value", "keep_next_value", "remove_resets", "interpolate"])})


def _go_pow10(n):
    """Go's math.Pow10 (src/math/pow10.go): a PRODUCT / QUOTIENT of two table literals, not the literal 1eN itself for |n| >= 32"""
    if 0 <= n <= 308:
        return float("1e%d" % (n // 32 * 32)) * float("1e%d" % (n % 32))
    if -323 <= n <= 0:
        return float("1e-%d" % ((-n) // 32 * 32)) / float("1e%d" % ((-n) % 32))
    return 0.0 if n < 0 else float("inf")


def transform(name, dev_ptr, nrows, points, *scalar_args, ctx=None):
    """transform.go value functions in place on a DEVICE matrix [nrows x points] (vmb_transform).  scalar_args: the function's scalar
    arguments (numbers or per-point arrays, getScalar): clamp(min, max), clamp_min(min), clamp_max(max), round(nearest = 1)"""
    ctx = ctx or _lib.default_context()
    name = name.lower()
    bc = lambda x: np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (points,)))
    a1 = a2 = None
    if name == "clamp":
        a1, a2 = bc(scalar_args[0]), bc(scalar_args[1])
    elif name in ("clamp_min", "clamp_max"):
        a1 = bc(scalar_args[0])
    elif name == "round":
        a1 = bc(scalar_args[0] if scalar_args else 1.0)
        # p10 = math.Pow10(-e), (_, e) = decimal.FromFloat(nearest)  transform.go:2341
        uniq, inv = np.unique(a1, return_inverse=True)
        p10u = np.empty(uniq.size)
        for k, n in enumerate(uniq):
            if np.isnan(n) or np.isinf(n) or n == 0:
                p10u[k] = 1.0
                continue
            _, e = decimal.append_float_to_decimal(np.array([n], dtype=np.float64))
            p10u[k] = _go_pow10(-int(e))
        a2 = np.ascontiguousarray(p10u[inv])
    fp = lambda a: a.ctypes.data_as(_lib.f64p) if a is not None else None
    check(lib().vmb_transform(ctx.h, TRANSFORM_FUNCS[name], C.c_void_p(int(dev_ptr)), int(nrows), int(points), fp(a1), fp(a2)))


def aggr_quantile(phis, vals_dev_ptr, nseries, points, out_dev_ptr, group_ids=None, ngroups=1, ctx=None):
    """quantile(phi, q) by (...) / median (phi = 0.5)  aggr.go:1217 on a DEVICE matrix -> out_dev_ptr [ngroups x points]"""
    ctx = ctx or _lib.default_context()
    g = np.zeros(nseries, dtype=np.uint32) if group_ids is None else np.ascontiguousarray(group_ids, dtype=np.uint32)
    ph = np.ascontiguousarray(np.broadcast_to(np.asarray(phis, dtype=np.float64), (points,)))
    check(lib().vmb_aggr_quantile(ctx.h, C.c_void_p(int(vals_dev_ptr)), int(nseries), int(points), g.ctypes.data_as(_lib.u32p), int(ngroups),
                                  ph.ctypes.data_as(_lib.f64p), C.c_void_p(int(out_dev_ptr))))
