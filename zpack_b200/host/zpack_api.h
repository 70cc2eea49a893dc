/* zpack_api.h — the ZPack library interface as this drop-in exports it.
 *
 * Declares, in this project's own words, the ABI of the reference's public header
 * (/root/reference/lib/zpack.h): the same 52 Linux entry points (:237-742), the same caller-allocated
 * struct layouts (:71-184 — callers memset them to zero and own them, tests/read_archive.c:93-94) and
 * the same return codes (:189-218).  A program compiled against the reference header links and runs
 * against libzpack.so built from zpack_host.cpp; tests/test_host_lib.py checks the layouts against the
 * reference header when it is available and runs the reference's own test programs on the result.
 */
#ifndef ZPACK_B200_HOST_API_H
#define ZPACK_B200_HOST_API_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t zpack_u8;
typedef uint16_t zpack_u16;
typedef uint32_t zpack_u32;
typedef uint64_t zpack_u64;
typedef zpack_u8 zpack_bool;

/* container constants (docs/specs.md:18-79; lib/zpack.h:36-52) */
enum {
    ZPK_SIG_HEADER = 0x154b505a, ZPK_SIG_DATA = 0x144b505a, ZPK_SIG_CDR = 0x134b505a, ZPK_SIG_EOCDR = 0x124b505a,
    ZPK_HEADER_BYTES = 6, ZPK_SIG_BYTES = 4, ZPK_CDR_HEADER_BYTES = 20, ZPK_ENTRY_FIXED_BYTES = 35,
    ZPK_EOCDR_BYTES = 12, ZPK_MIN_ARCHIVE_BYTES = 42, ZPK_MAX_NAME = 65535, ZPK_VERSION_MIN = 1, ZPK_VERSION_MAX = 1
};

typedef enum { ZPACK_COMPRESSION_NONE = 0, ZPACK_COMPRESSION_ZSTD = 1, ZPACK_COMPRESSION_LZ4 = 2 } zpack_compression_method;

typedef struct zpack_file_entry_s {   /* lib/zpack.h:71-80, 48 bytes */
    char *filename;
    zpack_u64 offset, comp_size, uncomp_size, hash;
    zpack_u8 comp_method;
} zpack_file_entry;

typedef struct zpack_reader_s {       /* lib/zpack.h:85-110 */
    zpack_u16 version;
    zpack_file_entry *file_entries;
    zpack_u64 file_count, comp_size, uncomp_size;
    size_t file_size;
    void *zstd_dctx;                  /* unused here */
    void *lz4f_dctx;                  /* unused here: the GPU context is process-wide */
    size_t last_return;               /* per-entry status of the last GPU call */
    zpack_u64 cdr_offset, eocdr_offset;
    zpack_u8 *buffer;
    zpack_bool buffer_shared;
    FILE *file;
} zpack_reader;

typedef struct zpack_compress_options_s { zpack_compression_method method; int level; } zpack_compress_options;

typedef struct zpack_file_s {         /* lib/zpack.h:125-134 */
    char *filename;
    zpack_u8 *buffer;
    zpack_u64 size;
    zpack_compress_options *options;
    void *cctx;
} zpack_file;

typedef struct zpack_writer_s {       /* lib/zpack.h:139-164 */
    zpack_u8 *buffer;
    size_t buffer_capacity;
    FILE *file;
    size_t file_size, write_offset;
    zpack_file_entry *file_entries;
    zpack_u64 fe_capacity, file_count;
    void *zstd_cctx, *lz4f_cctx;
    size_t last_return;
    zpack_u64 cdr_offset, eocdr_offset;
} zpack_writer;

typedef struct zpack_stream_s {       /* lib/zpack.h:169-184 */
    zpack_u8 *next_in;  size_t avail_in, total_in;
    zpack_u8 *next_out; size_t avail_out, total_out;
    size_t read_back;
    void *xxh3_state;                 /* library-owned: here the staging state of the GPU-backed stream */
} zpack_stream;

enum zpack_result {                   /* lib/zpack.h:189-218 */
    ZPACK_OK, ZPACK_ERROR_ARCHIVE_NOT_LOADED, ZPACK_ERROR_WRITER_NOT_OPENED, ZPACK_ERROR_OPEN_FAILED,
    ZPACK_ERROR_SEEK_FAILED, ZPACK_ERROR_FILE_TOO_SMALL, ZPACK_ERROR_SIGNATURE_INVALID, ZPACK_ERROR_READ_FAILED,
    ZPACK_ERROR_BLOCK_SIZE_INVALID, ZPACK_ERROR_VERSION_INCOMPATIBLE, ZPACK_ERROR_MALLOC_FAILED,
    ZPACK_ERROR_FILE_NOT_FOUND, ZPACK_ERROR_BUFFER_TOO_SMALL, ZPACK_ERROR_DECOMPRESS_FAILED,
    ZPACK_ERROR_COMPRESS_FAILED, ZPACK_ERROR_FILE_HASH_MISMATCH, ZPACK_ERROR_FILE_OFFSET_INVALID,
    ZPACK_ERROR_FILE_INCOMPLETE, ZPACK_ERROR_FILE_SIZE_INVALID, ZPACK_ERROR_COMP_METHOD_INVALID,
    ZPACK_ERROR_WRITE_FAILED, ZPACK_ERROR_STREAM_INVALID, ZPACK_ERROR_HASH_FAILED, ZPACK_ERROR_FILENAME_TOO_LONG,
    ZPACK_ERROR_NOT_AVAILABLE
};

/* low-level section readers (lib/zpack.h:237-331) */
int zpack_read_header_memory(const zpack_u8 *buffer, zpack_u16 *version);
int zpack_read_header(FILE *fp, zpack_u16 *version);
int zpack_read_data_header_memory(const zpack_u8 *buffer);
int zpack_read_data_header(FILE *fp);
int zpack_read_eocdr_memory(const zpack_u8 *buffer, zpack_u64 *cdr_offset);
int zpack_read_eocdr(FILE *fp, zpack_u64 eocdr_offset, zpack_u64 *cdr_offset);
int zpack_read_cdr_header_memory(const zpack_u8 *buffer, zpack_u64 *count, zpack_u64 *block_size);
int zpack_read_file_entry_memory(const zpack_u8 *buffer, zpack_u64 *size_left, zpack_file_entry *entry, size_t *entry_size);
int zpack_read_file_entries_memory(const zpack_u8 *buffer, zpack_file_entry **entries, zpack_u64 header_count,
                                   zpack_u64 block_size, zpack_u64 *count, zpack_u64 *total_cs, zpack_u64 *total_us);
int zpack_read_cdr_memory(const zpack_u8 *buffer, size_t size_left, zpack_file_entry **entries, zpack_u64 *count,
                          zpack_u64 *total_cs, zpack_u64 *total_us);
int zpack_read_cdr(FILE *fp, zpack_u64 cdr_offset, zpack_file_entry **entries, zpack_u64 *count, zpack_u64 *total_cs,
                   zpack_u64 *total_us);

/* reader (lib/zpack.h:350-470) */
int zpack_read_archive_memory(zpack_reader *reader);
int zpack_read_archive(zpack_reader *reader);
int zpack_read_raw_file(zpack_reader *reader, zpack_file_entry *entry, zpack_u8 *buffer, size_t max_size);
int zpack_read_file(zpack_reader *reader, zpack_file_entry *entry, zpack_u8 *buffer, size_t max_size, void *dctx);
int zpack_read_raw_file_stream(zpack_reader *reader, zpack_file_entry *entry, zpack_stream *stream, size_t *in_size);
int zpack_read_file_stream(zpack_reader *reader, zpack_file_entry *entry, zpack_stream *stream, void *dctx);
int zpack_init_reader(zpack_reader *reader, const char *path);
int zpack_init_reader_cfile(zpack_reader *reader, FILE *fp);
int zpack_init_reader_memory(zpack_reader *reader, const zpack_u8 *buffer, size_t size);
int zpack_init_reader_memory_shared(zpack_reader *reader, zpack_u8 *buffer, size_t size);
void zpack_reset_reader_dctx(zpack_reader *reader);
void zpack_close_reader(zpack_reader *reader);

/* extension: every listed entry in ONE GPU batch; out slot i = out + dst_off[i], capacity dst_cap[i] */
int zpack_read_files(zpack_reader *reader, zpack_file_entry *entries, zpack_u64 n, zpack_u8 *out, zpack_u64 out_size,
                     const zpack_u64 *dst_off, const zpack_u64 *dst_cap, int *status);

/* writer (lib/zpack.h:494-630) */
int zpack_init_writer(zpack_writer *writer, const char *path);
int zpack_init_writer_cfile(zpack_writer *writer, FILE *fp);
int zpack_init_writer_heap(zpack_writer *writer, size_t initial_size);
int zpack_write_header(zpack_writer *writer);
int zpack_write_header_ex(zpack_writer *writer, zpack_u16 version);
int zpack_write_data_header(zpack_writer *writer);
int zpack_write_files(zpack_writer *writer, zpack_file *files, zpack_u64 file_count);
int zpack_write_files_from_archive(zpack_writer *writer, zpack_reader *reader, zpack_file_entry *entries, zpack_u64 file_count);
int zpack_write_file_stream(zpack_writer *writer, zpack_compress_options *options, zpack_stream *stream, void *cctx);
int zpack_write_file_stream_end(zpack_writer *writer, char *filename, zpack_compress_options *options,
                                zpack_stream *stream, void *cctx);
int zpack_write_cdr(zpack_writer *writer);
int zpack_write_cdr_ex(zpack_writer *writer, zpack_file_entry *entries, zpack_u64 file_count);
int zpack_write_eocdr(zpack_writer *writer);
int zpack_write_eocdr_ex(zpack_writer *writer, zpack_u64 cdr_offset);
int zpack_write_archive(zpack_writer *writer, zpack_file *files, zpack_u64 file_count);
void zpack_close_writer(zpack_writer *writer);

/* stream (lib/zpack.h:644-656) */
int zpack_init_stream(zpack_stream *stream);
void zpack_reset_stream(zpack_stream *stream);
void zpack_close_stream(zpack_stream *stream);

/* utils (lib/zpack.h:671-742) */
size_t zpack_get_dstream_in_size(zpack_compression_method method);
size_t zpack_get_dstream_out_size(zpack_compression_method method);
size_t zpack_get_cstream_in_size(zpack_compression_method method);
size_t zpack_get_cstream_out_size(zpack_compression_method method);
zpack_file_entry *zpack_get_file_entry(const char *filename, zpack_file_entry *file_entries, zpack_u64 file_count);
zpack_bool zpack_read_stream_done(zpack_stream *stream, zpack_file_entry *entry);
void *zpack_create_cctx(zpack_compression_method method);
void *zpack_create_dctx(zpack_compression_method method);
void zpack_free_cctx(zpack_compression_method method, void *cctx);
void zpack_free_dctx(zpack_compression_method method, void *dctx);

#ifdef __cplusplus
}
#endif
#endif
