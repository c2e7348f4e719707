//go:build cgo && vmb200

package promql

import (
	"errors"

	"github.com/VictoriaMetrics/VictoriaMetrics/app/vmselect/netstorage"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/bytesutil"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/querytracer"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/storage"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/vmb200"
)

// errVMB200Declined: the GPU path does not take this query shape; the caller runs the stock Go code (rss.RunParallel + rc.Do).
var errVMB200Declined = errors.New("vmb200: query shape not handled on the GPU")

// newVMB200Cfg builds vmb_rollup_cfg from the first rollupConfig of getRollupConfigs (rollup.go:374) and the decisions
// evalRollupFuncNoCache makes around it (eval.go:1716-1760).  Function names come from the generated table
// (vmb200_func_ids.go); an unknown name, a multi-output rollup or a function with per-point arguments is declined.
func newVMB200Cfg(funcName string, rcs []*rollupConfig, preFuncRemovesResets, dropStale bool) (*vmb200.RollupCfg, error) {
	id, ok := vmb200FuncIDs[funcName]
	if !ok || vmb200FuncsWithArgs[funcName] || len(rcs) != 1 {
		return nil, errVMB200Declined
	}
	rc := rcs[0]
	var cfg vmb200.RollupCfg
	cfg.func_id = id
	cfg.start, cfg.end, cfg.step, cfg.window = rc.Start, rc.End, rc.Step, rc.Window
	cfg.lookback_delta = rc.LookbackDelta
	cfg.min_staleness_ms = minStalenessInterval.Milliseconds()
	cfg.samples_scanned_per_call = int32(rc.samplesScannedPerCall)
	if preFuncRemovesResets {
		cfg.flags |= vmb200.RemoveCounterResets
	}
	if dropStale {
		cfg.flags |= vmb200.DropStaleNaNs
	}
	if rc.MayAdjustWindow {
		cfg.flags |= vmb200.MayAdjustWindow
	}
	if rc.isDefaultRollup {
		cfg.flags |= vmb200.IsDefaultRollup
	}
	return &cfg, nil
}

// evalRollupNoIncrementalAggregateGPU == evalRollupNoIncrementalAggregate (eval.go:1845) for single-output rollups: one
// library call for all series instead of rss.RunParallel + rc.Do per series.
func evalRollupNoIncrementalAggregateGPU(funcName string, keepMetricNames bool, rss *netstorage.Results, rcs []*rollupConfig,
	removesResets, dropStale bool, sharedTimestamps []int64) ([]*timeseries, uint64, error) {
	descs, payload, names, tr := rss.CollectBlocks()
	if len(descs) == 0 {
		return nil, 0, nil
	}
	c := vmb200.Get()
	defer vmb200.Put(c)
	c.SetDedupInterval(storage.GetDedupInterval())
	points := len(sharedTimestamps)
	out := make([]float64, len(names)*points)
	cfg, err := newVMB200Cfg(funcName, rcs, removesResets, dropStale)
	if err != nil {
		return nil, 0, err // errVMB200Declined: evalRollupNoIncrementalAggregate runs
	}
	scanned, err := c.EvalRollup(descs, payload, tr.MinTimestamp, tr.MaxTimestamp, cfg, out)
	if err != nil {
		return nil, 0, err
	}
	tss := make([]*timeseries, len(names))
	for i, name := range names {
		var ts timeseries
		if err := ts.MetricName.Unmarshal(bytesutil.ToUnsafeBytes(name)); err != nil {
			return nil, 0, err
		}
		if !keepMetricNames && !rollupFuncsKeepMetricName[funcName] { // rollup.go:267-287
			ts.MetricName.ResetMetricGroup()
		}
		ts.Values = out[i*points : (i+1)*points : (i+1)*points]
		ts.Timestamps = sharedTimestamps
		ts.denyReuse = true
		tss[i] = &ts
	}
	return tss, scanned, nil
}

// vmb200AggrIDs: incremental aggregates (aggr_incremental.go:17 incrementalAggrFuncCallbacksMap) -> enum vmb_aggr_func.
var vmb200AggrIDs = map[string]int32{"sum": 0, "min": 1, "max": 2, "avg": 3, "count": 4, "sum2": 5, "geomean": 6, "any": 7, "group": 8}

// evalRollupWithIncrementalAggregateGPU == evalRollupWithIncrementalAggregate (eval.go:1804): the series are folded into one
// partial state per group on the GPU (the per-worker incrementalAggrContext of aggr_incremental.go:184); groupIDs come from
// the caller's marshalMetricNameSorted keys after removeGroupTags (aggr_incremental.go:107-113), dense, identical on every
// vmselect that takes part in a multi-GPU query.
func evalRollupWithIncrementalAggregateGPU(funcName, aggrName string, rss *netstorage.Results, rcs []*rollupConfig, removesResets,
	dropStale bool, groupIDs []uint32, ngroups int, points int) ([]float64, uint64, error) {
	aggrID, ok := vmb200AggrIDs[aggrName]
	if !ok {
		return nil, 0, errVMB200Declined
	}
	cfg, err := newVMB200Cfg(funcName, rcs, removesResets, dropStale)
	if err != nil {
		return nil, 0, err
	}
	descs, payload, _, tr := rss.CollectBlocks()
	c := vmb200.Get()
	defer vmb200.Put(c)
	c.SetDedupInterval(storage.GetDedupInterval())
	out := make([]float64, ngroups*points)
	scanned, err := c.EvalRollupAggr(descs, payload, tr.MinTimestamp, tr.MaxTimestamp, cfg, int(aggrID), groupIDs, ngroups, out)
	return out, scanned, err
}

// evalRollupGPUOrStock is what evalRollupFuncNoCache (eval.go:1680) calls instead of evalRollupNoIncrementalAggregate: the GPU
// path when a B200 is usable and the query shape is taken, the stock Go path otherwise (no B200, declined shape, or a library
// error -- the error is logged once per minute, the query still succeeds on the CPU).
func evalRollupGPUOrStock(qt *querytracer.Tracer, funcName string, keepMetricNames bool, rss *netstorage.Results, rcs []*rollupConfig,
	preFunc func(values []float64, timestamps []int64), removesResets, dropStale bool, sharedTimestamps []int64) ([]*timeseries, uint64, error) {
	if vmb200.Available() {
		tss, scanned, err := evalRollupNoIncrementalAggregateGPU(funcName, keepMetricNames, rss, rcs, removesResets, dropStale, sharedTimestamps)
		if err == nil {
			return tss, scanned, nil
		}
		if !errors.Is(err, errVMB200Declined) {
			vmb200.LogErrorRateLimited(err)
		}
	}
	return evalRollupNoIncrementalAggregate(qt, funcName, keepMetricNames, rss, rcs, preFunc, sharedTimestamps)
}
