// Go-side baseline for the path libvmb200 replaces (SURVEY.md 8(d)): written for this repository, NOT part of the
// reference.  There is no Go toolchain in the build container, so this file has never been compiled here; it only uses
// exported lib/encoding + lib/decimal functions and the package-private rollup pieces whose names and signatures are
// cited below.
//
// Usage (in a VictoriaMetrics checkout):
//   cp vmb200_path_timing_test.go app/vmselect/promql/
//   go test ./app/vmselect/promql/ -run xxx -bench 'BenchmarkVMB200' -benchtime 5x -cpu 1,16,128
//
// Workload = bench.py's configs[1]: per series one block of 8192 rows, timestamps t0 + 15000*i (MarshalTypeDeltaConst),
// values node_cpu_seconds_total-like mantissas (increments U[0,1500] at scale -2, reset with p = 1e-4), then
// rate(m[5m]) on the grid start = t0+5m, end = t0+8191*15s, step = 15s (8172 points).  One benchmark iteration processes
// nSeries series: Block.UnmarshalData's work (UnmarshalTimestamps + UnmarshalValues, lib/storage/block.go:250),
// decimal.AppendDecimalToFloat (block.go:327), removeCounterResets (rollup.go:921), rollupConfig.Do (rollup.go:688).
// The reported metric "samples/s" is raw samples decoded and scanned per second, like bench.py's `value`.
package promql

import (
	"math/rand"
	"sync/atomic"
	"testing"

	"github.com/VictoriaMetrics/VictoriaMetrics/lib/decimal"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/encoding"
)

const (
	vmb200Rows    = 8192
	vmb200Scrape  = 15000
	vmb200T0      = int64(1700000000000)
	vmb200Scale   = int16(-2)
	vmb200NSeries = 2048 // distinct series per benchmark iteration (x 8192 rows = 16.8 M samples)
)

type vmb200Block struct {
	tsData, valData   []byte
	tsMT, valMT       encoding.MarshalType
	tsFirst, valFirst int64
}

func vmb200MakeBlocks() []vmb200Block {
	r := rand.New(rand.NewSource(1234))
	ts := make([]int64, vmb200Rows)
	for i := range ts {
		ts[i] = vmb200T0 + int64(i)*vmb200Scrape
	}
	blocks := make([]vmb200Block, vmb200NSeries)
	vals := make([]int64, vmb200Rows)
	for s := range blocks {
		v := r.Int63n(1e9)
		for i := range vals {
			if i > 0 {
				if r.Float64() < 1e-4 {
					v = 0
				}
				v += r.Int63n(1501)
			}
			vals[i] = v
		}
		b := &blocks[s]
		b.tsData, b.tsMT, b.tsFirst = encoding.MarshalTimestamps(nil, ts, 64)
		b.valData, b.valMT, b.valFirst = encoding.MarshalValues(nil, vals, 64)
	}
	return blocks
}

func BenchmarkVMB200DecodeRate5m(b *testing.B) {
	blocks := vmb200MakeBlocks()
	start := vmb200T0 + 5*60*1000
	end := vmb200T0 + int64(vmb200Rows-1)*vmb200Scrape
	rcProto := rollupConfig{
		Func:               rollupDerivFast, // "rate" (rollup.go:74)
		Start:              start,
		End:                end,
		Step:               vmb200Scrape,
		Window:             5 * 60 * 1000,
		MaxPointsPerSeries: 1e5,

		samplesScannedPerCall: 2, // rollupFuncsSamplesScannedPerCall["rate"] (rollup.go:256)
	}
	rcProto.Timestamps = rcProto.getTimestamps()

	var next atomic.Int64
	b.ReportAllocs()
	b.ResetTimer()
	b.RunParallel(func(pb *testing.PB) {
		rc := rcProto // rollupConfig.Do only reads rc
		var tsBuf, valBuf []int64
		var fvals, out []float64
		var sink float64
		for pb.Next() {
			// one pb.Next() = one series (a worker of netstorage.RunParallel)
			blk := &blocks[int(next.Add(1))%len(blocks)]
			var err error
			tsBuf, err = encoding.UnmarshalTimestamps(tsBuf[:0], blk.tsData, blk.tsMT, blk.tsFirst, vmb200Rows)
			if err != nil {
				panic(err)
			}
			valBuf, err = encoding.UnmarshalValues(valBuf[:0], blk.valData, blk.valMT, blk.valFirst, vmb200Rows)
			if err != nil {
				panic(err)
			}
			fvals = decimal.AppendDecimalToFloat(fvals[:0], valBuf, vmb200Scale)
			removeCounterResets(fvals, tsBuf, 0)
			out, _ = rc.Do(out[:0], fvals, tsBuf)
			sink += out[len(out)-1]
		}
		SinkLock.Lock()
		Sink += sink
		SinkLock.Unlock()
	})
	b.StopTimer()
	b.ReportMetric(float64(b.N)*vmb200Rows/b.Elapsed().Seconds(), "samples/s")
}
