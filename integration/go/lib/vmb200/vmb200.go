//go:build cgo && vmb200

// Package vmb200 binds libvmb200.so (include/vmb200.h): the B200 implementation of the block codec and rollup executor.
//
// cgo rules followed throughout: only pointer-free Go memory ([]byte, []int64, []float64, []BlockDesc) is passed, for the
// duration of the call; the library copies what it needs to device memory and never retains host pointers.
package vmb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/vmb200/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/vmb200 -lvmb200
#include <stdlib.h>
#include "vmb200.h"
*/
import "C"

import (
	"fmt"
	"log"
	"runtime"
	"sync"
	"sync/atomic"
	"time"
	"unsafe"
)

// BlockDesc mirrors vmb_block_desc (64 bytes) == lib/storage/block_header.go:19 blockHeader + series index.
type BlockDesc = C.vmb_block_desc

// RollupCfg mirrors vmb_rollup_cfg == rollupConfig (app/vmselect/promql/rollup.go:574) + the preFunc decisions of
// getRollupConfigs (rollup.go:374).
type RollupCfg = C.vmb_rollup_cfg

// Flags of RollupCfg.flags.
const (
	RemoveCounterResets = C.VMB_RC_REMOVE_COUNTER_RESETS
	DropStaleNaNs       = C.VMB_RC_DROP_STALE_NANS
	MayAdjustWindow     = C.VMB_RC_MAY_ADJUST_WINDOW
	IsDefaultRollup     = C.VMB_RC_IS_DEFAULT_ROLLUP
)

// Ctx wraps vmb_ctx. A Ctx serialises its launches on one CUDA stream: use one per concurrent query worker.
type Ctx struct{ p *C.vmb_ctx }

func lastError(rc C.int, what string) error {
	return fmt.Errorf("%s: vmb200 error %d: %s", what, int(rc), C.GoString(C.vmb_last_error()))
}

// NewCtx creates a context on the given CUDA device; it fails when no sm_100 device is usable (there is no CPU fallback
// inside the library: the caller keeps using the stock Go path then).
func NewCtx(device int) (*Ctx, error) {
	var p *C.vmb_ctx
	if rc := C.vmb_ctx_create(C.int(device), &p); rc != 0 {
		return nil, lastError(rc, "vmb_ctx_create")
	}
	return &Ctx{p: p}, nil
}

// Close releases the device memory held by c.
func (c *Ctx) Close() {
	if c.p != nil {
		C.vmb_ctx_destroy(c.p)
		c.p = nil
	}
}

// SetDedupInterval passes storage.GetDedupInterval() (lib/storage/dedup.go:15), in milliseconds.
func (c *Ctx) SetDedupInterval(ms int64) {
	C.vmb_ctx_set_dedup_interval(c.p, C.int64_t(ms))
}

var (
	poolOnce sync.Once
	pool     chan *Ctx
)

// Enabled reports whether a B200 is usable; on first call it creates n contexts on device 0.
func Enabled(n int) bool {
	poolOnce.Do(func() {
		pool = make(chan *Ctx, n)
		for i := 0; i < n; i++ {
			c, err := NewCtx(0)
			if err != nil {
				break
			}
			pool <- c
		}
	})
	return cap(pool) > 0 && len(pool)+inUse() > 0
}

var inUseN int32

func inUse() int { return int(atomic.LoadInt32(&inUseN)) }

// Get / Put hand out pooled contexts.
func Get() *Ctx  { c := <-pool; atomic.AddInt32(&inUseN, 1); return c }
func Put(c *Ctx) { atomic.AddInt32(&inUseN, -1); pool <- c }

func descPtr(d []BlockDesc) *C.vmb_block_desc { return (*C.vmb_block_desc)(unsafe.Pointer(&d[0])) }
func bytePtr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&b[0]))
}

// EvalRollup == evalRollupNoIncrementalAggregate (eval.go:1845) for all series of a query: out is [nseries*points].
func (c *Ctx) EvalRollup(descs []BlockDesc, payload []byte, trMin, trMax int64, cfg *RollupCfg, out []float64) (uint64, error) {
	var scanned C.uint64_t
	rc := C.vmb_eval_rollup_host(c.p, descPtr(descs), C.size_t(len(descs)), bytePtr(payload), C.size_t(len(payload)),
		C.int64_t(trMin), C.int64_t(trMax), cfg, (*C.double)(unsafe.Pointer(&out[0])), nil, &scanned)
	runtime.KeepAlive(descs)
	runtime.KeepAlive(payload)
	runtime.KeepAlive(out)
	if rc != 0 {
		return 0, lastError(rc, "vmb_eval_rollup_host")
	}
	return uint64(scanned), nil
}

// EvalRollupAggr == evalRollupWithIncrementalAggregate (eval.go:1804): out is [ngroups*points], the only data that
// crosses PCIe on the way back. groupIDs[i] is the dense id of series i's group (aggr_incremental.go:113).
func (c *Ctx) EvalRollupAggr(descs []BlockDesc, payload []byte, trMin, trMax int64, cfg *RollupCfg, aggrID int,
	groupIDs []uint32, ngroups int, out []float64) (uint64, error) {
	var scanned C.uint64_t
	rc := C.vmb_eval_rollup_aggr_host(c.p, descPtr(descs), C.size_t(len(descs)), bytePtr(payload), C.size_t(len(payload)),
		C.int64_t(trMin), C.int64_t(trMax), cfg, C.int(aggrID), (*C.uint32_t)(unsafe.Pointer(&groupIDs[0])),
		C.uint32_t(ngroups), (*C.double)(unsafe.Pointer(&out[0])), nil, &scanned)
	runtime.KeepAlive(descs)
	runtime.KeepAlive(payload)
	runtime.KeepAlive(groupIDs)
	runtime.KeepAlive(out)
	if rc != 0 {
		return 0, lastError(rc, "vmb_eval_rollup_aggr_host")
	}
	return uint64(scanned), nil
}

// Blocks are the compressed blocks of a query resident in HBM (vmb_blocks).
type Blocks struct{ p *C.vmb_blocks }

func (b *Blocks) Close() {
	if b.p != nil {
		C.vmb_blocks_free(b.p)
		b.p = nil
	}
}

// UploadBlockRefs == BlockRef.Init + BlockRef.MustReadBlock (lib/storage/search.go:38,73) for every block of a query that
// lives in one part: headers is what tmpBlocksFile.WriteBlockRefData (tmp_blocks_file.go:110) kept per block -- the
// marshaled 81-byte blockHeaders, in series order --, timestampsBin / valuesBin are the part's mmap'ed data files.
func (c *Ctx) UploadBlockRefs(headers, timestampsBin, valuesBin []byte) (*Blocks, error) {
	var b *C.vmb_blocks
	rc := C.vmb_blocks_upload_part(c.p, bytePtr(headers), C.size_t(len(headers)/81), bytePtr(timestampsBin), C.size_t(len(timestampsBin)),
		bytePtr(valuesBin), C.size_t(len(valuesBin)), &b)
	runtime.KeepAlive(headers)
	runtime.KeepAlive(timestampsBin)
	runtime.KeepAlive(valuesBin)
	if rc != 0 {
		return nil, lastError(rc, "vmb_blocks_upload_part")
	}
	return &Blocks{p: b}, nil
}

// EvalRollupResident == the per-series closure of evalRollupNoIncrementalAggregate (eval.go:1855-1866) for every series of
// b, with the result left in device memory (dOut: a device pointer to [nseries*points] float64) for the operators that
// follow in the expression tree (vmb_binary_op, vmb_aggr_quantile, vmb_topk_*, vmb_series_from_matrix for a subquery).
func (c *Ctx) EvalRollupResident(b *Blocks, trMin, trMax int64, cfg *RollupCfg, dOut unsafe.Pointer) (uint64, error) {
	var scanned C.uint64_t
	rc := C.vmb_eval_rollup_device(c.p, b.p, C.int64_t(trMin), C.int64_t(trMax), cfg, (*C.double)(dOut), &scanned)
	if rc != 0 {
		return 0, lastError(rc, "vmb_eval_rollup_device")
	}
	return uint64(scanned), nil
}

// ---- multi-GPU: one vmselect process per GPU; the partial states of aggr(rollup) by (...) are merged by the library's NCCL
// all-reduce (include/vmb200.h "multi-GPU").  The 128-byte unique id travels over vmselect's own RPC.

// CommUniqueID == ncclGetUniqueId on the process that owns rank 0.
func CommUniqueID() ([128]byte, error) {
	var id [128]byte
	if rc := C.vmb_comm_get_unique_id((*C.uint8_t)(unsafe.Pointer(&id[0]))); rc != 0 {
		return id, lastError(rc, "vmb_comm_get_unique_id")
	}
	return id, nil
}

// CommInit joins the communicator of nranks processes as `rank`.
func (c *Ctx) CommInit(id [128]byte, nranks, rank int) error {
	if rc := C.vmb_ctx_comm_init(c.p, (*C.uint8_t)(unsafe.Pointer(&id[0])), C.int(nranks), C.int(rank)); rc != 0 {
		return lastError(rc, "vmb_ctx_comm_init")
	}
	return nil
}

// Available reports whether a usable B200 was found (Enabled created at least one context for the pool behind Get / Put).
func Available() bool { return Enabled(runtime.GOMAXPROCS(0)) }

// LogErrorRateLimited logs a library error at most once per minute (the query falls back to the stock Go path).
func LogErrorRateLimited(err error) {
	now := time.Now().Unix()
	if last := atomic.LoadInt64(&lastErrLog); now-last >= 60 && atomic.CompareAndSwapInt64(&lastErrLog, last, now) {
		log.Printf("vmb200: falling back to the CPU path: %s", err)
	}
}

var lastErrLog int64

// DescFromHeader == blockHeader.Unmarshal + validate (block_header.go:122, :230) on the 81-byte wire form.
func DescFromHeader(d *BlockDesc, header []byte) error {
	if len(header) != 81 {
		return fmt.Errorf("too short block header; got %d bytes; want 81 bytes", len(header))
	}
	if rc := C.vmb_block_desc_from_header(d, bytePtr(header), nil); rc != 0 {
		return lastError(rc, "vmb_block_desc_from_header")
	}
	return nil
}

// SetPlacement rebases a descriptor onto the caller's payload arena and assigns its dense series index (the cgo field
// types are local to this package, so other packages go through this setter).
func SetPlacement(d *BlockDesc, tsOff, valOff uint64, seriesIdx uint32) {
	d.ts_off = C.uint64_t(tsOff)
	d.val_off = C.uint64_t(valOff)
	d.series_idx = C.uint32_t(seriesIdx)
}

// UnmarshalValues == encoding.UnmarshalValues (lib/encoding/encoding.go:111), one column per call (compat / tests).
func (c *Ctx) UnmarshalValues(dst []int64, src []byte, mt byte, firstValue int64, itemsCount int) ([]int64, error) {
	n := len(dst)
	dst = append(dst, make([]int64, itemsCount)...)
	rc := C.vmb_unmarshal_int64(c.p, (*C.int64_t)(unsafe.Pointer(&dst[n])), C.size_t(itemsCount), bytePtr(src),
		C.size_t(len(src)), C.int(mt), C.int64_t(firstValue))
	runtime.KeepAlive(src)
	if rc != 0 {
		return nil, fmt.Errorf("cannot unmarshal %d values from len(src)=%d bytes: vmb200 error %d", itemsCount, len(src), int(rc))
	}
	return dst, nil
}

// AppendDecimalToFloat == decimal.AppendDecimalToFloat (lib/decimal/decimal.go:100).
func (c *Ctx) AppendDecimalToFloat(dst []float64, va []int64, e int16) []float64 {
	n := len(dst)
	dst = append(dst, make([]float64, len(va))...)
	if len(va) > 0 {
		C.vmb_decimal_to_float(c.p, (*C.double)(unsafe.Pointer(&dst[n])), (*C.int64_t)(unsafe.Pointer(&va[0])),
			C.size_t(len(va)), C.int16_t(e))
		runtime.KeepAlive(va)
	}
	return dst
}

// DecompressZSTDBatch == encoding.DecompressZSTD (lib/encoding/compress.go:27) for many frames at once: frame i is
// frames[offs[i]:offs[i+1]]; the result holds frame i at dst[dstOffs[i] : dstOffs[i]+dstLens[i]].
func (c *Ctx) DecompressZSTDBatch(frames []byte, offs []uint64) (dst []byte, dstOffs []uint64, dstLens []uint32, err error) {
	n := len(offs) - 1
	var bound C.uint64_t
	if rc := C.vmb_zstd_decompress_bound(bytePtr(frames), (*C.uint64_t)(unsafe.Pointer(&offs[0])), C.size_t(n), &bound); rc != 0 {
		return nil, nil, nil, lastError(rc, "vmb_zstd_decompress_bound")
	}
	dst = make([]byte, int(bound)+1)
	dstOffs = make([]uint64, n)
	dstLens = make([]uint32, n)
	rc := C.vmb_zstd_decompress_batch(c.p, bytePtr(frames), (*C.uint64_t)(unsafe.Pointer(&offs[0])), C.size_t(n), bytePtr(dst),
		C.size_t(len(dst)), (*C.uint64_t)(unsafe.Pointer(&dstOffs[0])), (*C.uint32_t)(unsafe.Pointer(&dstLens[0])), nil)
	runtime.KeepAlive(frames)
	runtime.KeepAlive(offs)
	if rc != 0 {
		return nil, nil, nil, lastError(rc, "vmb_zstd_decompress_batch")
	}
	return dst, dstOffs, dstLens, nil
}

// UnmarshalIndexBlock == unmarshalBlockHeaders (lib/storage/block_header.go:261) on one decompressed index block.
func UnmarshalIndexBlock(dst []BlockDesc, tsids []byte, data []byte, count int) ([]BlockDesc, []byte, error) {
	n := len(dst)
	dst = append(dst, make([]BlockDesc, count)...)
	t := len(tsids)
	tsids = append(tsids, make([]byte, 24*count)...)
	rc := C.vmb_index_block_unmarshal(&dst[n], bytePtr(tsids[t:]), C.size_t(count), bytePtr(data), C.size_t(len(data)))
	runtime.KeepAlive(data)
	if rc != 0 {
		return dst[:n], tsids[:t], lastError(rc, "vmb_index_block_unmarshal")
	}
	return dst, tsids, nil
}
