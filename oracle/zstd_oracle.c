/* oracle/zstd_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h).  PLACEHOLDER: filled in with the
 * Zstandard frame decoder restatement in the zstd milestone. */
#include "oracle.h"
int orc_zstd_decode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *out_len) {
    (void)src; (void)src_len; (void)dst; (void)dst_cap; *out_len = 0;
    return ORC_DECOMPRESS_FAILED;
}
