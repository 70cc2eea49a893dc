/* oracle/entry_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * One ZPack entry read the way zpack_read_file does it (lib/zpack_read.c:326-471):
 * guards, method dispatch, decode, XXH3-64 verify over uncomp_size bytes.
 */
#include "oracle.h"
#include <string.h>

int orc_read_entry(int method, const uint8_t *comp, size_t comp_size, uint8_t *dst,
                   size_t max_size, size_t uncomp_size, uint64_t expect_hash, uint64_t *digest) {
    size_t produced = 0;
    int rc;
    if (comp_size == 0) return ORC_OK;                              /* zpack_read.c:328 — no hash */
    if (max_size < uncomp_size) return ORC_BUFFER_TOO_SMALL;        /* :329 */
    switch (method) {
    case 0:                                                         /* :352-368 */
        if (uncomp_size > comp_size) return 18;                     /* ZPACK_ERROR_FILE_SIZE_INVALID */
        memcpy(dst, comp, uncomp_size);
        break;
    case 1:                                                         /* :370-390 */
        rc = orc_zstd_decode(comp, comp_size, dst, max_size, &produced);
        if (rc != ORC_OK) return ORC_DECOMPRESS_FAILED;
        break;
    case 2:                                                         /* :396-453 */
        rc = orc_lz4f_decode(comp, comp_size, dst, max_size, &produced);
        if (rc != ORC_OK) return rc;
        break;
    default:
        return 19;                                                  /* ZPACK_ERROR_COMP_METHOD_INVALID */
    }
    uint64_t h = orc_xxh3_64(dst, uncomp_size);                     /* :466 */
    if (digest) *digest = h;
    return h == expect_hash ? ORC_OK : ORC_HASH_MISMATCH;
}

/* Batch loop for the CPU baseline when oracle/_ref is absent: same slicing as ref_driver.c. */
long orc_unpack_range(const uint8_t *archive, const uint64_t *offset, const uint64_t *comp,
                      const uint64_t *uncomp, const uint64_t *hash, const uint8_t *method,
                      size_t first, size_t end, size_t stride, uint8_t *out, size_t out_cap,
                      uint64_t *bytes_done) {
    long bad = 0;
    uint64_t total = 0, dg;
    for (size_t i = first; i < end; i += stride) {
        int rc = orc_read_entry(method[i], archive + offset[i], comp[i], out, out_cap, uncomp[i], hash[i], &dg);
        if (rc != ORC_OK) ++bad;
        total += uncomp[i];
    }
    if (bytes_done) *bytes_done = total;
    return bad;
}
