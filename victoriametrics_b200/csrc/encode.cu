// Write path on the GPU: encoding.MarshalValues / MarshalTimestamps for many equal-length columns at once
// (Block.MarshalData lib/storage/block.go:192 as the merge path calls it, lib/storage/merge.go:227-235).
//
//   lib/encoding/encoding.go:119   marshalInt64Array: type selection
//   lib/encoding/encoding.go:289   isConst, :311 isDeltaConst, :331 isGauge
//   lib/encoding/nearest_delta.go:15, nearest_delta2.go:15   delta / delta-of-delta (+ :83 nearestDelta for precisionBits < 64)
//   lib/encoding/int.go:107        MarshalVarInt64s (zig-zag LEB128)
//
// One warp per column, two kernels:
//   k_marshal_plan:  the three detection scans as ONE pass of warp reductions (every early return of isGauge is an "exists"
//                    condition, the reset count an order-independent sum), the MarshalType, and the byte size of the varint
//                    stream (a warp sum of per-value lengths); for precisionBits < 64 the lossy deltas -- a sequential state
//                    machine (trailingZeros, v) -- are produced by one lane into a scratch column first;
//   k_marshal_pack:  per 32-value chunk the lengths are scanned over the warp and every lane writes its varint at its offset.
// The byte offsets between the two come from an exclusive prefix sum of the sizes (host, ncols entries).  The zstd stage that
// follows for streams of >= 128 bytes (encoding.go:152-167, including the 0.9 rule that may turn type 1 -> 5 / 4 -> 6) runs on
// host threads with the library's zstd writer (marshal.inc), on the bytes this kernel produced.
#pragma once

struct MarshalParams {
    const int64_t* vals;      // [ncols x rows]
    int64_t* deltas;          // scratch [ncols x rows] (precisionBits < 64 only, else nullptr)
    uint8_t* out;             // varint streams, column c at offs[c]
    const uint64_t* offs;     // [ncols] (pack)
    uint32_t* sizes;          // [ncols] (plan)
    uint8_t* mts;             // [ncols]: 3 const, 2 delta-const, 4 gauge -> nearest delta, 1 counter -> nearest delta2
    int64_t* firsts;          // [ncols]
    uint32_t ncols, rows;
    uint32_t pb;              // precisionBits 1..64
};

namespace {

__device__ __forceinline__ uint32_t varint_len(uint64_t u) {  // bytes of the LEB128 form, int.go:107
    const uint32_t bits = u ? 64u - (uint32_t)__clzll((long long)u) : 1u;
    return (bits + 6u) / 7u;
}
__device__ __forceinline__ uint64_t zz64(int64_t v) { return (uint64_t)((v << 1) ^ (v >> 63)); }

__device__ __forceinline__ uint32_t enc_bitlen(uint64_t x) { return x ? 64u - (uint32_t)__clzll((long long)x) : 0u; }
// getTrailingZeros nearest_delta.go:134
__device__ __forceinline__ uint32_t enc_trailing_zeros(int64_t v, uint32_t pb) {
    const uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    const uint32_t vb = enc_bitlen(a);
    return vb <= pb ? 0u : vb - pb;
}
// nearestDelta nearest_delta.go:83 (uint8 arithmetic of the trailing-zeros state kept)
__device__ __forceinline__ void enc_nearest_delta(int64_t next, int64_t prev, uint32_t pb, uint32_t ptz, int64_t* dout, uint32_t* tzout) {
    const int64_t d = (int64_t)((uint64_t)next - (uint64_t)prev);
    const uint32_t dec = ptz ? ptz - 1u : 0u;
    if (d == 0) { *dout = 0; *tzout = dec; return; }
    const uint64_t origin = next < 0 ? (uint64_t)0 - (uint64_t)next : (uint64_t)next;
    const uint32_t ob = enc_bitlen(origin);
    if (ob <= pb) { *dout = d; *tzout = dec; return; }
    const uint32_t tz = ob - pb;
    if (tz > ((ptz + 4u) & 0xffu)) { *dout = d; *tzout = (ptz + 2u) & 0xffu; return; }
    if (((tz + 4u) & 0xffu) < ptz) { *dout = d; *tzout = (ptz - 2u) & 0xffu; return; }
    const bool minus = d < 0;
    const uint64_t ad = minus ? (uint64_t)0 - (uint64_t)d : (uint64_t)d;
    const uint64_t mask = tz >= 64u ? 0ull : (~(uint64_t)0 << tz);
    const uint64_t nd = ad & mask;
    *dout = (int64_t)(minus ? (uint64_t)0 - nd : nd);
    *tzout = tz;
}

// the i-th value of the varint stream of a column (i >= 1): lossless forms straight from the column, lossy ones from the scratch
__device__ __forceinline__ int64_t enc_stream_value(const int64_t* a, const int64_t* dl, uint32_t i, bool delta2, bool lossy) {
    if (lossy) return dl[i];
    if (!delta2 || i == 1) return (int64_t)((uint64_t)a[i] - (uint64_t)a[i - 1]);
    return (int64_t)((uint64_t)a[i] - 2ull * (uint64_t)a[i - 1] + (uint64_t)a[i - 2]);  // next - v - d1, nearest_delta2.go:30
}

}  // namespace

__global__ void __launch_bounds__(128) k_marshal_plan(MarshalParams P) {
    const int lane = lane_id();
    const uint32_t wpg = gridDim.x * (blockDim.x >> 5);
    for (uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < P.ncols; c += wpg) {
        const int64_t* a = P.vals + (size_t)c * P.rows;
        const uint32_t n = P.rows;
        const int64_t a0 = a[0];
        // ---- detection: isConst, isDeltaConst, isGauge in one pass
        const uint64_t d1 = n >= 2 ? (uint64_t)a[1] - (uint64_t)a0 : 0ull;
        bool all_eq = true, dconst = n >= 2, imm = false;
        uint32_t resets = 0;
        for (uint32_t i = 1 + lane; i < n; i += 32) {
            const int64_t v = a[i], pv = a[i - 1];
            all_eq &= v == a0;
            dconst &= ((uint64_t)v - (uint64_t)pv) == d1;
            if (v < pv) {
                if (v < 0 || v > (pv >> 3)) imm = true;  // encoding.go:349-357
                resets++;
            }
        }
        all_eq = __all_sync(VMB_FULL, all_eq);
        dconst = __all_sync(VMB_FULL, dconst);
        imm = __any_sync(VMB_FULL, imm);
#pragma unroll
        for (int o = 16; o; o >>= 1) resets += __shfl_xor_sync(VMB_FULL, resets, o);
        bool gauge = false;
        if (n >= 2) gauge = a0 < 0 || imm || (resets > 2 && resets > (n >> 3));
        uint32_t mt, size = 0;
        if (all_eq) mt = 3;                                     // MarshalTypeConst encoding.go:124
        else if (dconst) {
            mt = 2;                                             // MarshalTypeDeltaConst :130
            size = varint_len(zz64((int64_t)d1));
        } else {
            mt = gauge ? 4u : 1u;
            const bool delta2 = !gauge;
            uint32_t pb = P.pb;
            if (gauge && pb < 6) pb += 2;                        // encoding.go:141
            const bool lossy = pb < 64;
            int64_t* dl = lossy ? P.deltas + (size_t)c * P.rows : nullptr;
            if (lossy) {
                // the state machine of nearest_delta.go:36-42 / nearest_delta2.go:39-46, sequential by construction
                if (lane == 0) {
                    if (!delta2) {
                        int64_t v = a0;
                        uint32_t tz = enc_trailing_zeros(v, pb);
                        for (uint32_t i = 1; i < n; i++) {
                            int64_t d;
                            enc_nearest_delta(a[i], v, pb, tz, &d, &tz);
                            v = (int64_t)((uint64_t)v + (uint64_t)d);
                            dl[i] = d;
                        }
                    } else {
                        int64_t dd = (int64_t)d1, v = a[1];
                        dl[1] = dd;
                        uint32_t tz = enc_trailing_zeros(v, pb);
                        for (uint32_t i = 2; i < n; i++) {
                            int64_t d2;
                            enc_nearest_delta((int64_t)((uint64_t)a[i] - (uint64_t)v), dd, pb, tz, &d2, &tz);
                            dd = (int64_t)((uint64_t)dd + (uint64_t)d2);
                            v = (int64_t)((uint64_t)v + (uint64_t)dd);
                            dl[i] = d2;
                        }
                    }
                }
                __syncwarp();
            }
            for (uint32_t i = 1 + lane; i < n; i += 32) size += varint_len(zz64(enc_stream_value(a, dl, i, delta2, lossy)));
#pragma unroll
            for (int o = 16; o; o >>= 1) size += __shfl_xor_sync(VMB_FULL, size, o);
        }
        if (lane == 0) {
            P.mts[c] = (uint8_t)mt;
            P.sizes[c] = size;
            P.firsts[c] = a0;
        }
    }
}

__global__ void __launch_bounds__(128) k_marshal_pack(MarshalParams P) {
    const int lane = lane_id();
    const uint32_t wpg = gridDim.x * (blockDim.x >> 5);
    for (uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < P.ncols; c += wpg) {
        const uint32_t mt = P.mts[c];
        if (mt == 3) continue;
        const int64_t* a = P.vals + (size_t)c * P.rows;
        uint8_t* out = P.out + P.offs[c];
        const uint32_t n = P.rows;
        if (mt == 2) {
            if (lane == 0) {
                uint64_t u = zz64((int64_t)((uint64_t)a[1] - (uint64_t)a[0]));
                uint32_t k = 0;
                while (u >= 0x80) { out[k++] = (uint8_t)(u | 0x80); u >>= 7; }
                out[k] = (uint8_t)u;
            }
            continue;
        }
        const bool delta2 = mt == 1;
        uint32_t pb = P.pb;
        if (!delta2 && pb < 6) pb += 2;
        const bool lossy = pb < 64;
        const int64_t* dl = lossy ? P.deltas + (size_t)c * P.rows : nullptr;
        uint32_t base = 0;
        for (uint32_t i0 = 1; i0 < n; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint64_t u = 0;
            uint32_t len = 0;
            if (i < n) {
                u = zz64(enc_stream_value(a, dl, i, delta2, lossy));
                len = varint_len(u);
            }
            uint32_t inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(VMB_FULL, inc, o);
                if (lane >= o) inc += t;
            }
            uint8_t* w = out + base + inc - len;
            for (uint32_t k = 0; k + 1 < len; k++) {
                w[k] = (uint8_t)(u | 0x80);
                u >>= 7;
            }
            if (len) w[len - 1] = (uint8_t)u;
            base += __shfl_sync(VMB_FULL, inc, 31);
        }
    }
}

// ---- decimal.AppendFloatToDecimal (lib/decimal/decimal.go:173-257) for many equal-length columns: one warp per column.
// FromFloat per value (:437, the same code as the host encoder: marshal.inc is compiled for both sides), the minimum exponent over
// the non-special values (:203-211), the down-shift that keeps every up-scaled mantissa inside int64 (:213-224), the rescale
// (:231-249); the all-zeros / all-ones fast paths (:177-184) come out of the same reductions.
__global__ void __launch_bounds__(128) k_float_to_decimal(const double* __restrict__ src, int64_t* __restrict__ dst, int16_t* __restrict__ ea,
                                                          int16_t* __restrict__ scales, uint32_t ncols, uint32_t rows) {
    const int lane = lane_id();
    const uint32_t wpg = gridDim.x * (blockDim.x >> 5);
    for (uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < ncols; c += wpg) {
        const double* f = src + (size_t)c * rows;
        int64_t* v = dst + (size_t)c * rows;
        int16_t* e = ea + (size_t)c * rows;
        bool zeros = true, ones = true;
        int min_exp = 32767;
        for (uint32_t i = lane; i < rows; i += 32) {
            const double x = f[i];
            const unsigned long long b = (unsigned long long)__double_as_longlong(x);
            zeros &= b == 0ull;
            ones &= b == 0x3ff0000000000000ull;
            int64_t vi;
            int16_t ei;
            vmb_host::from_float(x, &vi, &ei);
            v[i] = vi;
            e[i] = ei;
            if (ei < min_exp && !vmb_host::special(vi)) min_exp = ei;
        }
        zeros = __all_sync(VMB_FULL, zeros);
        ones = __all_sync(VMB_FULL, ones);
#pragma unroll
        for (int o = 16; o; o >>= 1) min_exp = min(min_exp, __shfl_xor_sync(VMB_FULL, min_exp, o));
        __syncwarp();
        if (zeros || ones) {
            for (uint32_t i = lane; i < rows; i += 32) v[i] = ones ? 1 : 0;
            if (lane == 0) scales[c] = 0;
            continue;
        }
        int down = 0;
        for (uint32_t i = lane; i < rows; i += 32) {
            const int up = (int16_t)(e[i] - min_exp);
            const int d = (int16_t)(up - vmb_host::max_up_exponent(v[i]));
            down = max(down, d);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) down = max(down, __shfl_xor_sync(VMB_FULL, down, o));
        const int16_t mexp = (int16_t)(min_exp + down);
        for (uint32_t i = lane; i < rows; i += 32) {
            int64_t x = v[i];
            if (vmb_host::special(x)) continue;
            int adj = (int16_t)(e[i] - mexp);
            while (adj > 0) { x = (int64_t)((uint64_t)x * 10u); adj--; }
            while (adj < 0) { x /= 10; adj++; }
            v[i] = x;
        }
        if (lane == 0) scales[c] = mexp;
    }
}

void launch_marshal_plan(const MarshalParams& P, cudaStream_t st) {
    if (!P.ncols) return;
    uint32_t grid = (P.ncols + 3) / 4;
    if (grid > 148u * 16u) grid = 148u * 16u;
    k_marshal_plan<<<grid, 128, 0, st>>>(P);
}
void launch_marshal_pack(const MarshalParams& P, cudaStream_t st) {
    if (!P.ncols) return;
    uint32_t grid = (P.ncols + 3) / 4;
    if (grid > 148u * 16u) grid = 148u * 16u;
    k_marshal_pack<<<grid, 128, 0, st>>>(P);
}
