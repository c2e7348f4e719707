//go:build cgo && vmb200

package storage

import (
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/bytesutil"
	"github.com/VictoriaMetrics/VictoriaMetrics/lib/vmb200"
)

// AppendRawTo appends the block referenced by br, as it lies in the part files, to a descriptor array and a payload arena:
// the same two reads MustReadBlock performs (search.go:73-82), without UnmarshalData. seriesIdx is the dense index of the
// series inside the query (the blocks of one series must be appended consecutively, in any order).
func (br *BlockRef) AppendRawTo(descs []vmb200.BlockDesc, payload []byte, seriesIdx uint32) ([]vmb200.BlockDesc, []byte) {
	var hdr [81]byte
	h := br.bh.Marshal(hdr[:0]) // block_header.go:104
	var d vmb200.BlockDesc
	if err := vmb200.DescFromHeader(&d, h); err != nil {
		// the header was validated when the part was opened (block_header.go:155)
		panic(err)
	}
	tsOff := len(payload)
	payload = bytesutil.ResizeWithCopyMayOverallocate(payload, tsOff+int(br.bh.TimestampsBlockSize))
	br.p.timestampsFile.MustReadAt(payload[tsOff:], int64(br.bh.TimestampsBlockOffset))
	valOff := len(payload)
	payload = bytesutil.ResizeWithCopyMayOverallocate(payload, valOff+int(br.bh.ValuesBlockSize))
	br.p.valuesFile.MustReadAt(payload[valOff:], int64(br.bh.ValuesBlockOffset))
	vmb200.SetPlacement(&d, uint64(tsOff), uint64(valOff), seriesIdx)
	return append(descs, d), payload
}
