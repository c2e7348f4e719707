//go:build cgo && vmb200

package promql

import (
	"github.com/VictoriaMetrics/VictoriaMetrics/app/vmselect/netstorage"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/bytesutil"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/storage"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/vmb200"
)

// vmb200FuncIDs maps rollup function names to VMB_RF_* (include/vmb200.h lists them in the order of rollupFuncs, rollup.go:24).
var vmb200FuncIDs = map[string]int{ /* "rate": C.VMB_RF_RATE, ... generated from include/vmb200.h */ }

// newVMB200Cfg builds vmb_rollup_cfg from the first rollupConfig of getRollupConfigs (rollup.go:374) and the decisions
// evalRollupFuncNoCache makes around it (eval.go:1716-1760).
func newVMB200Cfg(funcName string, rc *rollupConfig, preFuncRemovesResets, dropStale bool) *vmb200.RollupCfg {
	var cfg vmb200.RollupCfg
	cfg.func_id = int32(vmb200FuncIDs[funcName])
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
	return &cfg
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
	cfg := newVMB200Cfg(funcName, rcs[0], removesResets, dropStale)
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
