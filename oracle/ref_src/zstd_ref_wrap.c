/* ORACLE support (test infrastructure): thin C wrapper over the reference's own libzstd 1.5.7
 * (vendor/github.com/valyala/gozstd/libzstd_linux_amd64.a), calling it exactly as gozstd does
 * (gozstd.go:171 ZSTD_compressCCtx, gozstd.go:331 ZSTD_decompressDCtx).  Contexts are reused per thread, like gozstd's
 * sync.Pool of cctx/dctx wrappers (gozstd.go:93, :222).  Built into oracle/_ref/. */
#include "zstd.h"
static __thread ZSTD_CCtx* t_cctx;
static __thread ZSTD_DCtx* t_dctx;
size_t ref_zstd_compress(void* dst, size_t cap, const void* src, size_t n, int level) {
    if (!t_cctx) t_cctx = ZSTD_createCCtx();
    return ZSTD_compressCCtx(t_cctx, dst, cap, src, n, level);
}
size_t ref_zstd_decompress(void* dst, size_t cap, const void* src, size_t n) {
    if (!t_dctx) t_dctx = ZSTD_createDCtx();
    return ZSTD_decompressDCtx(t_dctx, dst, cap, src, n);
}
/* a frame WITH a content checksum (ZSTD_c_checksumFlag): gozstd never writes one; used to test that the decoders verify it */
size_t ref_zstd_compress_checksum(void* dst, size_t cap, const void* src, size_t n, int level) {
    if (!t_cctx) t_cctx = ZSTD_createCCtx();
    ZSTD_CCtx_reset(t_cctx, ZSTD_reset_session_and_parameters);
    ZSTD_CCtx_setParameter(t_cctx, ZSTD_c_compressionLevel, level);
    ZSTD_CCtx_setParameter(t_cctx, ZSTD_c_checksumFlag, 1);
    size_t r = ZSTD_compress2(t_cctx, dst, cap, src, n);
    ZSTD_CCtx_reset(t_cctx, ZSTD_reset_session_and_parameters);
    return r;
}
unsigned ref_zstd_is_error(size_t r) { return ZSTD_isError(r); }
unsigned ref_zstd_version(void) { return ZSTD_versionNumber(); }
