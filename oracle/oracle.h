/* oracle/oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the algorithms on ZPack's per-entry hot path, written from the
 * published formats and checked against (a) the reference's golden vectors and upstream
 * known-answer tables and (b) the unmodified reference compiled into oracle/_ref/.
 * Nothing under zpack_b200/ may include, link or call this; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do, and only as the checker.
 *
 * Status codes deliberately reuse the reference's enum zpack_result numbering
 * (/root/reference/lib/zpack.h:189-218) so parity tests compare integers directly.
 */
#ifndef ZPB200_ORACLE_H
#define ZPB200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_OK = 0,
    ORC_BUFFER_TOO_SMALL = 12,   /* ZPACK_ERROR_BUFFER_TOO_SMALL   */
    ORC_DECOMPRESS_FAILED = 13,  /* ZPACK_ERROR_DECOMPRESS_FAILED  */
    ORC_COMPRESS_FAILED = 14,    /* ZPACK_ERROR_COMPRESS_FAILED    */
    ORC_HASH_MISMATCH = 15,      /* ZPACK_ERROR_FILE_HASH_MISMATCH */
    ORC_FILE_INCOMPLETE = 17     /* ZPACK_ERROR_FILE_INCOMPLETE    */
};

/* XXH3-64, seed 0, default secret (xxHash 0.8.0; externals/xxHash/xxhash.h:3866). */
uint64_t orc_xxh3_64(const void *data, size_t len);

/* XXH32 seed-able (LZ4 frame header/ block / content checksums; externals/lz4/lib/xxhash.c). */
uint32_t orc_xxh32(const void *data, size_t len, uint32_t seed);

/* LZ4 block decode with an optional prefix window directly before dst
 * (externals/lz4/lib/lz4.c:1737-2165, prefix mode :2408-2414).
 * Returns decoded size, or -1 on malformed input / overflow. `prefix` bytes before dst are readable. */
long orc_lz4_block_decode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
                          size_t prefix);

/* Decode every LZ4 frame (and skippable frame) in [src,src+src_len) into dst, the way
 * zpack_read_file drives LZ4F_decompress (lib/zpack_read.c:396-453; lz4frame.c:1384-1879).
 * *out_len receives the bytes produced.  Returns an ORC_ code. */
int orc_lz4f_decode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
                    size_t *out_len);

/* Greedy single-probe hash-table LZ4 block compressor following lz4.c:851-1240 (level < 3).
 * `dict_len` bytes immediately before src are a linked-block prefix the table may reference;
 * `table` is the caller-owned 4096-entry position table that persists across linked blocks
 * (positions are offsets from `base`), `accel` the acceleration (1 = default).
 * Returns compressed size, or 0 when the output would not fit dst_cap (=> stored block). */
size_t orc_lz4_block_encode(const uint8_t *base, size_t src_off, size_t src_len,
                            uint8_t *dst, size_t dst_cap, uint32_t *table, int accel);

/* Whole-frame writer as zpack_compress_file does it (lib/zpack_write.c:192-214):
 * block-linked 64 KB blocks, no checksums, no content size.  `independent` != 0 emits
 * B.Indep=1 (FLG 0x60) with a fresh table per block (what the GPU packer writes).
 * Returns frame size or 0 if dst_cap is too small. */
size_t orc_lz4f_encode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
                       int level, int independent);
size_t orc_lz4f_bound(size_t src_len);

/* Zstandard frame decoder (all frames + skippable frames in the buffer), following
 * externals/zstd/doc/zstd_compression_format.md and the lib/decompress sources.
 * Returns ORC_OK / ORC_DECOMPRESS_FAILED / ORC_BUFFER_TOO_SMALL. */
int orc_zstd_decode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
                    size_t *out_len);
/* the same with a caller-owned context of orc_zstd_ctx_size() bytes (one per thread) */
int orc_zstd_decode_ctx(void *ctx, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
                        size_t *out_len);
size_t orc_zstd_ctx_size(void);

/* One ZPack entry, exactly as zpack_read_file (lib/zpack_read.c:326-471) would treat it:
 * dispatch on method (0 none, 1 zstd, 2 lz4), decode, then XXH3 verify.  *digest gets the
 * digest of dst[0..uncomp_size). */
int orc_read_entry(int method, const uint8_t *comp, size_t comp_size, uint8_t *dst,
                   size_t max_size, size_t uncomp_size, uint64_t expect_hash, uint64_t *digest);

#ifdef __cplusplus
}
#endif
#endif
