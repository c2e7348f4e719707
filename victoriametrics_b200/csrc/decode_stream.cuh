// Value-major varint / delta decoder (included by decode.cu).
//
// One warp decodes one nearest-delta(2) stream (lib/encoding/nearest_delta.go:53, nearest_delta2.go:57,
// int.go:182-284) in 512-byte tiles:
//   1. byte-major: every lane loads 16 bytes (coalesced), finds the varint terminators (bytes < 0x80) with bit tricks and
//      publishes their positions in shared memory (compacted with one warp scan of the per-lane counts);
//   2. value-major: the T values of the tile are dealt out in contiguous runs of c = ceil(T/32) (rounded up to odd:
//      bank-conflict-free) values per lane; a lane assembles each of its varints from the staged bytes (branch-free for
//      <= 4 bytes, loop for longer ones), zig-zag decodes it and keeps running sums;
//   3. one warp scan of the (count, sum, sum-of-prefix-sums) triples -- associative under wrapping int64 arithmetic, so the
//      result is bit-identical to the sequential Go loop -- then every lane replays its run and emits final values.
// Compared with assigning values to the lane that holds their terminator byte this keeps all 32 lanes equally busy
// whatever the varint widths are (no 16-step predicated parse/emit).
#pragma once

#define DS_TILE 512
#define DS_BYTES_WORDS ((16 + DS_TILE + 16) / 4)

struct DecodeSmem {  // per warp
    uint32_t bytes[DS_BYTES_WORDS];  // [0,16): last 16 bytes of the previous tile, [16, 16+512): this tile, then padding
    unsigned short pos[DS_TILE];     // terminator positions (byte offset inside the tile) in stream order
    long long val[DS_TILE + 32];     // decoded varints of the tile
};

__device__ __forceinline__ uint32_t ds_load_u32(const uint32_t* bytes, int byte_off) {  // byte_off relative to tile start, >= -16
    int a = byte_off + 16;
    uint32_t lo = bytes[a >> 2], hi = bytes[(a >> 2) + 1];
    return __funnelshift_r(lo, hi, (uint32_t)(a & 3) * 8);
}

template <class E>
__device__ int decode_delta_stream_v2(const uint8_t* __restrict__ src, uint32_t len, uint32_t n, int64_t first, bool delta2,
                                      E& em, DecodeSmem* sm) {
    const int lane = lane_id();
    if (n < (delta2 ? 2u : 1u)) return VMB_ERR_ROWS;  // Go: logger.Panicf("BUG: itemsCount ...")
    const uint32_t nvar = n - 1;
    if (len < nvar) return VMB_ERR_SHORT_SRC;  // int.go:183
    if (lane == 0) em.emit(0, first, first);
    uint32_t N = 0;            // varints decoded so far
    uint64_t D1 = 0;           // running first-order delta (delta2)
    uint64_t V = (uint64_t)first;
    int start_rel = 0;         // where the value in progress starts, relative to the current tile start (<= 0)
    int err = 0;
    uint32_t err_at = 0xffffffffu;  // index of the varint that raised err: the sequential Go loop reports the FIRST bad varint
                                    // (int.go:196-284) and never looks past the nvar-th one
    if (lane < 4) sm->bytes[lane] = 0;
    __syncwarp();

    // the 16 bytes of a lane are fetched one tile ahead (aligned words + funnel shift): the HBM round trip of tile t+1
    // overlaps the parse of tile t
    uint32_t nw0 = 0, nw1 = 0, nw2 = 0, nw3 = 0, nw4 = 0;
    auto fetch = [&](uint32_t tile) {
        const uint32_t o = tile + (uint32_t)lane * 16u;
        if (o < len) {
            uintptr_t a = (uintptr_t)(src + o);
            const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
            nw0 = w[0]; nw1 = w[1]; nw2 = w[2]; nw3 = w[3];
            nw4 = (a & 3) ? w[4] : 0u;
        }
    };
    fetch(0);
    for (uint32_t tile = 0; tile < len; tile += DS_TILE) {
        // ---- 1. this lane's 16 bytes (already in flight), stage them, find terminators
        const uint32_t o = tile + (uint32_t)lane * 16u;
        const uint32_t valid = o >= len ? 0u : (len - o >= 16u ? 16u : len - o);
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (valid) {
            const uint32_t sh = (uint32_t)((uintptr_t)(src + o) & 3) * 8;
            c0 = __funnelshift_r(nw0, nw1, sh);
            c1 = __funnelshift_r(nw1, nw2, sh);
            c2 = __funnelshift_r(nw2, nw3, sh);
            c3 = __funnelshift_r(nw3, nw4, sh);
        }
        if (tile + DS_TILE < len) fetch(tile + DS_TILE);
        uint32_t* sb = sm->bytes + 4 + lane * 4;
        sb[0] = c0; sb[1] = c1; sb[2] = c2; sb[3] = c3;
        uint32_t m = term_mask4(c0) | (term_mask4(c1) << 4) | (term_mask4(c2) << 8) | (term_mask4(c3) << 12);
        m &= valid >= 16 ? 0xffffu : ((1u << valid) - 1u);
        uint32_t cnt = (uint32_t)__popc(m);
        uint32_t inc = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(VMB_FULL, inc, off);
            if (lane >= off) inc += t;
        }
        const uint32_t T = __shfl_sync(VMB_FULL, inc, 31);
        {
            uint32_t j = inc - cnt, mm = m;
            while (mm) {
                int k = __ffs((int)mm) - 1;
                sm->pos[j++] = (unsigned short)(lane * 16 + k);
                mm &= mm - 1;
            }
        }
        __syncwarp();
        // ---- 2. value-major parse: lane owns values [k0, k1)
        uint32_t c = (T + 31) >> 5;
        c |= 1u;  // odd stride: conflict-free shared-memory access
        const uint32_t k0 = min(T, (uint32_t)lane * c), k1 = min(T, k0 + c);
        uint64_t s1 = 0, s2 = 0;
        int s = (k0 == 0 || k0 >= T) ? start_rel : (int)sm->pos[k0 - 1] + 1;  // first byte of the lane's first varint
#pragma unroll 2
        for (uint32_t k = k0; k < k1; k++) {
            const int e = (int)sm->pos[k];
            const int vl = e - s + 1;  // varint length in bytes
            uint64_t u;
            if (vl <= 4) {
                uint32_t w = ds_load_u32(sm->bytes, s);
                uint32_t x = (w & 0x7fu) | ((w & 0x7f00u) >> 1) | ((w & 0x7f0000u) >> 2) | ((w & 0x7f000000u) >> 3);
                u = x & (0xffffffffu >> (32 - 7 * vl));
            } else if (vl <= 10 && s >= -16) {
                u = 0;
                for (int b = 0; b < vl; b++) {
                    uint32_t byte = (sm->bytes[(s + b + 16) >> 2] >> (8 * ((s + b + 16) & 3))) & 0xffu;
                    if (b == 9) {  // 10th byte: int.go:269-275
                        if (byte > 1u && N + k < err_at) { err = VMB_ERR_VARINT_TOO_BIG; err_at = N + k; }
                        u |= (uint64_t)1 << 63;
                    } else {
                        u |= (uint64_t)(byte & 0x7fu) << (7 * b);
                    }
                }
            } else {
                if (N + k < err_at) { err = VMB_ERR_VARINT_TOO_LONG; err_at = N + k; }  // int.go:277
                u = 0;
            }
            long long v = (long long)(u >> 1) ^ -(long long)(u & 1);  // zig-zag decode int.go:82
            sm->val[k] = v;
            s1 += (uint64_t)v;
            s2 += s1;
            s = e + 1;
        }
        // ---- 3. scan (count, s1, s2); combine(A then B): s2 = s2A + s2B + cntB * s1A
        uint32_t icnt = k1 - k0;
        uint64_t is1 = s1, is2 = s2;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t acnt = __shfl_up_sync(VMB_FULL, icnt, off);
            uint64_t as1 = shfl_up_u64(is1, off);
            uint64_t as2 = shfl_up_u64(is2, off);
            if (lane >= off) {
                if (delta2) is2 = as2 + is2 + (uint64_t)icnt * as1;
                is1 += as1;
                icnt += acnt;
            }
        }
        uint32_t ecnt = __shfl_up_sync(VMB_FULL, icnt, 1);
        uint64_t es1 = shfl_up_u64(is1, 1), es2 = shfl_up_u64(is2, 1);
        if (lane == 0) { ecnt = 0; es1 = 0; es2 = 0; }
        // ---- 4. emit
        {
            uint32_t pos = 1u + N + k0;
            uint64_t d1 = D1 + es1;
            uint64_t v = delta2 ? (V + es2 + (uint64_t)ecnt * D1) : (V + es1);
            for (uint32_t k = k0; k < k1; k++) {
                uint64_t pv = v;
                uint64_t x = (uint64_t)sm->val[k];
                if (delta2) {
                    d1 += x;
                    v += d1;
                } else {
                    v += x;
                }
                if (pos < n) em.emit(pos, (int64_t)v, (int64_t)pv);
                pos++;
            }
        }
        // ---- tile carries
        uint64_t ts1 = shfl_u64(is1, 31), ts2 = shfl_u64(is2, 31);
        if (delta2) {
            V += ts2 + (uint64_t)T * D1;
            D1 += ts1;
        } else {
            V += ts1;
        }
        N += T;
        int last_end = T ? (int)sm->pos[T - 1] : -1;
        __syncwarp();  // everyone is done reading pos / bytes / val
        if (T) start_rel = last_end + 1 - DS_TILE;
        else {
            // a whole tile without a terminator: a varint of > 512 bytes (int.go:277); a partial last tile without one is a
            // truncated varint and is reported by the stream-level checks below
            start_rel -= DS_TILE;
            if (start_rel < -16) {
                if (tile + DS_TILE <= len && N < err_at) { err = VMB_ERR_VARINT_TOO_LONG; err_at = N; }
                start_rel = -16;
            }
        }
        if (lane == 31) { sm->bytes[0] = c0; sm->bytes[1] = c1; sm->bytes[2] = c2; sm->bytes[3] = c3; }
        __syncwarp();
    }
    // ---- stream-level checks (uniform)
    if (err_at >= nvar) err = 0;  // a malformed varint behind the last one that is read is only "unexpected tail"
    {
        uint32_t first = err ? err_at : 0xffffffffu;
#pragma unroll
        for (int off = 16; off; off >>= 1) first = min(first, __shfl_xor_sync(VMB_FULL, first, off));
        const uint32_t owner = __ballot_sync(VMB_FULL, err != 0 && err_at == first);
        err = owner ? __shfl_sync(VMB_FULL, err, __ffs((int)owner) - 1) : 0;
    }
    int werr = err;
    if (werr == 0) {
        bool ends_ok = len == 0 || src[len - 1] < 0x80;
        if (N < nvar) werr = VMB_ERR_SHORT_SRC;              // int.go:199 "cannot unmarshal varint from empty data"
        else if (N > nvar || !ends_ok) werr = VMB_ERR_TAIL;  // nearest_delta.go:65 unexpected tail
    }
    return werr;
}
