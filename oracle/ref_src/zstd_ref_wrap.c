/* ORACLE support (test infrastructure): thin C wrapper over the reference's own libzstd 1.5.7
 * (vendor/github.com/valyala/gozstd/libzstd_linux_amd64.a), calling it exactly as gozstd does
 * (gozstd.go:171 ZSTD_compressCCtx, gozstd.go:331 ZSTD_decompressDCtx).  Built into oracle/_ref/. */
#include "zstd.h"
size_t ref_zstd_compress(void* dst, size_t cap, const void* src, size_t n, int level) {
    ZSTD_CCtx* c = ZSTD_createCCtx();
    size_t r = ZSTD_compressCCtx(c, dst, cap, src, n, level);
    ZSTD_freeCCtx(c);
    return r;
}
size_t ref_zstd_decompress(void* dst, size_t cap, const void* src, size_t n) {
    ZSTD_DCtx* d = ZSTD_createDCtx();
    size_t r = ZSTD_decompressDCtx(d, dst, cap, src, n);
    ZSTD_freeDCtx(d);
    return r;
}
unsigned ref_zstd_is_error(size_t r) { return ZSTD_isError(r); }
unsigned ref_zstd_version(void) { return ZSTD_versionNumber(); }
