//go:build cgo && vmb200

package netstorage

import (
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/storage"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/vmb200"
)

// CollectBlocks returns every block of every series of rss as descriptors + one payload arena, in series order, and the
// metric name of each series. It replaces the per-series packedTimeseries.Unpack fan-out (netstorage.go:425) when the
// rollup runs on the GPU: the time-range trim, mergeSortBlocks (:566) and DeduplicateSamples happen inside libvmb200
// (blocks of one series may be appended in any order; pts.brs order is fine).
//
// tr is the search time range (Results.tr, netstorage.go:64) that Unpack passes to AppendRowsWithTimeRangeFilter.
func (rss *Results) CollectBlocks() (descs []vmb200.BlockDesc, payload []byte, metricNames []string, tr storage.TimeRange) {
	tbf := rss.tbf
	for i := range rss.packedTimeseries {
		pts := &rss.packedTimeseries[i]
		for _, br := range pts.brs {
			ref := tbf.MustReadBlockRefAt(br.partRef, br.addr) // tmp_blocks_file.go:168
			descs, payload = ref.AppendRawTo(descs, payload, uint32(i))
		}
		metricNames = append(metricNames, pts.metricName)
	}
	return descs, payload, metricNames, rss.tr
}
