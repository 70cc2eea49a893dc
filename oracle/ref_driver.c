/* oracle/ref_driver.c — TEST INFRASTRUCTURE ONLY.
 *
 * Thin C loops around the UNMODIFIED reference's public API (linked from _ref/libzpack_ref.so) so
 * that the CPU baseline is timed without Python in the loop: each host thread calls
 * ref_unpack_range() on its own slice with its own dctx, which lib/zpack.h:335-341 allows for a
 * buffer-mode reader.  Prototypes are restated from /root/reference/lib/zpack.h:383,728,742;
 * zpack_file_entry is 48 bytes (lib/zpack.h:71-80).
 */
#include <stddef.h>
#include <stdint.h>

int zpack_read_file(void *reader, void *entry, uint8_t *buffer, size_t max_size, void *dctx);
void *zpack_create_dctx(int method);
void zpack_free_dctx(int method, void *dctx);

/* entries first, first+stride, ... < end; returns the number of entries that did NOT return ZPACK_OK */
long ref_unpack_range(void *reader, uint8_t *entries, size_t first, size_t end, size_t stride,
                      uint8_t *out, size_t out_cap, int method, uint64_t *bytes_done) {
    void *dctx = zpack_create_dctx(method);
    long bad = 0;
    uint64_t total = 0;
    for (size_t i = first; i < end; i += stride) {
        uint8_t *e = entries + 48 * i;
        int rc = zpack_read_file(reader, e, out, out_cap, dctx);
        if (rc != 0) ++bad;
        total += *(uint64_t *)(e + 24); /* uncomp_size */
    }
    zpack_free_dctx(method, dctx);
    if (bytes_done) *bytes_done = total;
    return bad;
}

/* The same, BASELINE.md §4's shape: entry i is decoded into ITS OWN slice of one preallocated output,
 * out_base + i * slot_bytes (so the decoded bytes really land in host DRAM, like the GPU arm's D2H target). */
long ref_unpack_slices(void *reader, uint8_t *entries, size_t first, size_t end, size_t stride,
                       uint8_t *out_base, size_t slot_bytes, int method, uint64_t *bytes_done) {
    void *dctx = zpack_create_dctx(method);
    long bad = 0;
    uint64_t total = 0;
    for (size_t i = first; i < end; i += stride) {
        uint8_t *e = entries + 48 * i;
        int rc = zpack_read_file(reader, e, out_base + i * slot_bytes, slot_bytes, dctx);
        if (rc != 0) ++bad;
        total += *(uint64_t *)(e + 24); /* uncomp_size */
    }
    zpack_free_dctx(method, dctx);
    if (bytes_done) *bytes_done = total;
    return bad;
}
