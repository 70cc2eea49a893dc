/* oracle/lz4_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * LZ4 block + frame codec restated from the published formats
 * (externals/lz4/doc/lz4_Block_format.md, lz4_Frame_format.md) and from the behaviour of
 * lz4 1.9.3 as ZPack drives it.  Parity is PINNED by tests/test_oracle.py: golden archive
 * tests/workdir/archive_lz4.zpk (tests/archive.h:72-91) and differential runs against
 * oracle/_ref (LZ4F_decompress / LZ4F_compressFrame of the unmodified reference).
 */
#include "oracle.h"
#include <string.h>

#define MAGIC_FRAME     0x184D2204u
#define MAGIC_SKIP_BASE 0x184D2A50u

static uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}
static uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }
static void wr32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

/* ------------------------------------------------------------------ block decode
 * Follows LZ4_decompress_generic (lz4.c:1737-2165) in its safe / endOnInputSize / full-block
 * configuration.  `dst_cap` plays the role of the reference's `outputSize` (= the frame's
 * maxBlockSize: lz4frame.c:1683,1708).  End-of-block rules the reference enforces:
 *   - literals that end within 8 bytes of the input end must end exactly on it (lz4.c:2055-2077)
 *   - a match may not end inside the last 5 bytes of capacity (lz4.c:2139)
 *   - offset must stay inside prefix + produced bytes (lz4.c:2093)
 * Deliberate divergence: offset 0 is rejected here (1.9.3 copies zeros, lz4.c:303-318); every
 * such stream fails the digest in the reference, so both report an error for the entry.
 */
long orc_lz4_block_decode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
                          size_t prefix) {
    size_t ip = 0, op = 0;
    if (src_len == 0) return -1;                                   /* lz4.c:1782-1786 */
    for (;;) {
        unsigned token = src[ip++];
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= src_len) return -1;
                b = src[ip++];
                lit += b;
            } while (b == 255);
        }
        if (lit > src_len - ip || lit > dst_cap - op) return -1;
        int last = (ip + lit == src_len);
        if (!last && ip + lit + 8 > src_len) return -1;            /* 2 + 1 + LASTLITERALS */
        memcpy(dst + op, src + ip, lit);
        ip += lit; op += lit;
        if (last) return (long)op;

        size_t off = (size_t)src[ip] | (size_t)src[ip + 1] << 8;
        ip += 2;
        size_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (ip >= src_len) return -1;
                b = src[ip++];
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (off == 0 || off > op + prefix) return -1;
        if (ml > dst_cap - op || op + ml + 5 > dst_cap) return -1; /* last 5 bytes are literals */
        for (size_t i = 0; i < ml; ++i) dst[op + i] = dst[op + i - off]; /* overlap-safe */
        op += ml;
        if (ip >= src_len) return -1;
    }
}

/* ------------------------------------------------------------------ frame decode
 * One pass over every frame in the entry, reproducing the outcome of the loop at
 * lib/zpack_read.c:414-450 around LZ4F_decompress (lz4frame.c:1384-1879):
 *   bad magic / flags / header checksum / oversized block / bad block  -> DECOMPRESS_FAILED
 *   input ends inside a frame                                          -> FILE_INCOMPLETE
 *   output capacity reached inside a frame                             -> BUFFER_TOO_SMALL
 * (when both run out the reference tests avail_out first, zpack_read.c:446-449).
 */
static int need_more(size_t op, size_t cap) { return op < cap ? ORC_FILE_INCOMPLETE : ORC_BUFFER_TOO_SMALL; }

int orc_lz4f_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len) {
    size_t ip = 0, op = 0;
    int rc = ORC_OK;
    while (ip < n && op < cap) {                                   /* zpack_read.c:414 */
        if (n - ip < 7) { rc = need_more(op, cap); break; }         /* minFHSize */
        uint32_t magic = rd32(src + ip);
        if ((magic & 0xFFFFFFF0u) == MAGIC_SKIP_BASE) {             /* lz4frame.c:1125-1136 */
            if (n - ip < 8) { rc = need_more(op, cap); break; }
            size_t skip = rd32(src + ip + 4);
            if (skip > n - ip - 8) { ip = n; rc = need_more(op, cap); break; }
            ip += 8 + skip;
            continue;
        }
        if (magic != MAGIC_FRAME) { rc = ORC_DECOMPRESS_FAILED; break; }
        unsigned flg = src[ip + 4];
        if ((flg >> 1) & 1) { rc = ORC_DECOMPRESS_FAILED; break; }  /* reserved bit */
        if (((flg >> 6) & 3) != 1) { rc = ORC_DECOMPRESS_FAILED; break; }
        int indep = (flg >> 5) & 1, bsum = (flg >> 4) & 1, has_size = (flg >> 3) & 1,
            csum = (flg >> 2) & 1, has_dict = flg & 1;
        size_t hsize = 7 + (has_size ? 8 : 0) + (has_dict ? 4 : 0);
        if (n - ip < hsize) { ip = n; rc = need_more(op, cap); break; }
        unsigned bd = src[ip + 5];
        if ((bd >> 7) || ((bd >> 4) & 7) < 4 || (bd & 15)) { rc = ORC_DECOMPRESS_FAILED; break; }
        size_t max_block = (size_t)1 << (8 + 2 * ((bd >> 4) & 7));  /* 64K,256K,1M,4M */
        if ((uint8_t)(orc_xxh32(src + ip + 4, hsize - 5, 0) >> 8) != src[ip + hsize - 1]) {
            rc = ORC_DECOMPRESS_FAILED; break;                      /* lz4frame.c:294-298,1184 */
        }
        uint64_t remaining = has_size ? rd64(src + ip + 6) : 0;
        ip += hsize;
        size_t frame_start = op;
        int done = 0;
        while (!done) {
            if (n - ip < 4) { ip = n; rc = need_more(op, cap); break; }
            uint32_t bh = rd32(src + ip);
            ip += 4;
            if (bh == 0) {                                          /* EndMark */
                if (has_size && remaining != 0) { rc = ORC_DECOMPRESS_FAILED; break; }
                if (csum) {
                    if (n - ip < 4) { ip = n; rc = need_more(op, cap); break; }
                    if (rd32(src + ip) != orc_xxh32(dst + frame_start, op - frame_start, 0)) {
                        rc = ORC_DECOMPRESS_FAILED; break;
                    }
                    ip += 4;
                }
                done = 1;
                break;
            }
            size_t bsz = bh & 0x7FFFFFFFu;
            if (bsz > max_block) { rc = ORC_DECOMPRESS_FAILED; break; }
            if (bh & 0x80000000u) {                                 /* stored block: lz4frame.c:1534-1572 */
                size_t take = bsz;
                if (take > n - ip) take = n - ip;
                if (take > cap - op) take = cap - op;
                memcpy(dst + op, src + ip, take);
                if (bsum && take == bsz) {
                    if (n - ip - bsz < 4) { op += take; ip = n; rc = need_more(op, cap); break; }
                    if (rd32(src + ip + bsz) != orc_xxh32(src + ip, bsz, 0)) {
                        rc = ORC_DECOMPRESS_FAILED; break;
                    }
                }
                op += take; remaining -= take;
                if (take < bsz) { ip += take; rc = need_more(op, cap); break; }
                ip += bsz + (bsum ? 4 : 0);
                continue;
            }
            if (op == cap) { rc = ORC_BUFFER_TOO_SMALL; break; }    /* lz4frame.c:1527-1530 */
            if (n - ip < bsz + (bsum ? 4 : 0)) { ip = n; rc = need_more(op, cap); break; }
            if (bsum && rd32(src + ip + bsz) != orc_xxh32(src + ip, bsz, 0)) {
                rc = ORC_DECOMPRESS_FAILED; break;
            }
            size_t prefix = indep ? 0 : op - frame_start;
            if (cap - op >= max_block) {
                long got = orc_lz4_block_decode(src + ip, bsz, dst + op, max_block, prefix);
                if (got < 0) { rc = ORC_DECOMPRESS_FAILED; break; }
                op += (size_t)got; remaining -= (uint64_t)got;
            } else {
                /* not enough room for a worst-case block: the reference decodes into its
                 * private buffer and flushes what fits (lz4frame.c:1687-1745). */
                static __thread uint8_t tmp[(4u << 20) + 65536];
                size_t keep = prefix > 65536 ? 65536 : prefix;
                memcpy(tmp, dst + op - keep, keep);
                long got = orc_lz4_block_decode(src + ip, bsz, tmp + keep, max_block, keep);
                if (got < 0) { rc = ORC_DECOMPRESS_FAILED; break; }
                size_t fit = (size_t)got > cap - op ? cap - op : (size_t)got;
                memcpy(dst + op, tmp + keep, fit);
                op += fit; remaining -= (uint64_t)got;
                if (fit < (size_t)got) { ip += bsz + (bsum ? 4 : 0); rc = ORC_BUFFER_TOO_SMALL; break; }
            }
            ip += bsz + (bsum ? 4 : 0);
        }
        if (rc != ORC_OK) break;
    }
    *out_len = op;
    return rc;
}

/* ------------------------------------------------------------------ block encode
 * LZ4_compress_generic_validated (lz4.c:851-1240) in the configuration LZ4F uses for
 * levels < 3: byU32 table of 4096 entries, limitedOutput, prefix mode.  Positions in `table`
 * are offsets from `base` so the table carries over linked blocks (lz4frame.c:661-664,777-781).
 */
#define HASHLOG 12
/* Linked blocks: byU32 table, 12-bit hash of 5 bytes (lz4.c:706-716).  Independent blocks below
 * LZ4_64Klimit go through LZ4_compress_fast_extState_fastReset with a byU16 table: 13-bit hash of
 * 4 bytes (lz4.c:697-704,1266-1282) — flagged by bit 31 of `accel` (internal to this file). */
static int g_u16_mode;
static uint32_t hash5(const uint8_t *p) {
    if (g_u16_mode) return (rd32(p) * 2654435761u) >> (32 - (HASHLOG + 1));
    return (uint32_t)(((rd64(p) << 24) * 889523592379ull) >> (64 - HASHLOG));
}
static size_t count_eq(const uint8_t *a, const uint8_t *b, const uint8_t *a_lim) {
    const uint8_t *s = a;
    while (a < a_lim && *a == *b) { ++a; ++b; }
    return (size_t)(a - s);
}

size_t orc_lz4_block_encode(const uint8_t *base, size_t src_off, size_t src_len,
                            uint8_t *dst, size_t dst_cap, uint32_t *table, int accel) {
    const uint8_t *src = base + src_off;
    const uint8_t *ip = src, *anchor = src, *iend = src + src_len;
    const uint8_t *mflimit1 = iend - 12 + 1, *matchlimit = iend - 5;
    uint8_t *op = dst, *olimit = dst + dst_cap;
    if (accel < 1) accel = 1;
    if (src_len < 13) goto tail;                                    /* LZ4_minLength */

    table[hash5(ip)] = (uint32_t)(ip - base);
    ++ip;
    uint32_t fwd_h = hash5(ip);
    for (;;) {
        const uint8_t *match;
        const uint8_t *fwd_ip = ip;
        int step = 1, tries = accel << 6;                           /* LZ4_skipTrigger */
        for (;;) {
            uint32_t h = fwd_h;
            uint32_t cur = (uint32_t)(fwd_ip - base);
            uint32_t cand = table[h];
            ip = fwd_ip;
            fwd_ip += step;
            step = tries++ >> 6;
            if (fwd_ip > mflimit1) goto tail;
            match = base + cand;
            fwd_h = hash5(fwd_ip);
            table[h] = cur;
            if (cand + 65535 < cur) continue;
            if (rd32(match) == rd32(ip)) break;
        }
        while (ip > anchor && match > base && ip[-1] == match[-1]) { --ip; --match; }

        size_t lit = (size_t)(ip - anchor);
        uint8_t *token = op++;
        if (op + lit + 8 + lit / 255 > olimit) return 0;
        if (lit >= 15) {
            size_t r = lit - 15;
            *token = 0xF0;
            for (; r >= 255; r -= 255) *op++ = 255;
            *op++ = (uint8_t)r;
        } else {
            *token = (uint8_t)(lit << 4);
        }
        memcpy(op, anchor, lit);
        op += lit;
    next_match:
        op[0] = (uint8_t)(ip - match); op[1] = (uint8_t)((ip - match) >> 8);
        op += 2;
        {
            size_t mc = count_eq(ip + 4, match + 4, matchlimit);
            ip += mc + 4;
            if (op + 6 + (mc + 240) / 255 > olimit) return 0;
            if (mc >= 15) {
                *token += 15;
                mc -= 15;
                for (; mc >= 255; mc -= 255) *op++ = 255;
                *op++ = (uint8_t)mc;
            } else {
                *token += (uint8_t)mc;
            }
        }
        anchor = ip;
        if (ip >= mflimit1) break;
        table[hash5(ip - 2)] = (uint32_t)(ip - 2 - base);
        {
            uint32_t h = hash5(ip), cur = (uint32_t)(ip - base), cand = table[h];
            match = base + cand;
            table[h] = cur;
            if (cand + 65535 >= cur && rd32(match) == rd32(ip)) {
                token = op++;
                *token = 0;
                goto next_match;
            }
        }
        fwd_h = hash5(++ip);
    }
tail: {
        size_t run = (size_t)(iend - anchor);
        if (op + run + 1 + (run + 255 - 15) / 255 > olimit) return 0;
        if (run >= 15) {
            size_t r = run - 15;
            *op++ = 0xF0;
            for (; r >= 255; r -= 255) *op++ = 255;
            *op++ = (uint8_t)r;
        } else {
            *op++ = (uint8_t)(run << 4);
        }
        memcpy(op, anchor, run);
        op += run;
    }
    return (size_t)(op - dst);
}

/* ------------------------------------------------------------------ frame encode
 * zpack_compress_file's LZ4 arm (lib/zpack_write.c:192-214): zeroed preferences => 64 KB
 * blocks, linked, no checksums, no content size; header bytes 04 22 4D 18 40 40 C0.
 * Blocks that do not shrink are stored with bit 31 set (lz4frame.c:740-763).
 */
size_t orc_lz4f_bound(size_t src_len) {                             /* lz4frame.c:324-349, prefs NULL */
    size_t block = 65536, max_src = src_len + (block - 1);
    size_t full = max_src / block, partial = max_src & (block - 1);
    size_t last = src_len == 0 ? partial : 0;
    size_t nblocks = full + (last > 0);
    return 8 * nblocks + block * full + last + 8;
}

size_t orc_lz4f_encode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level,
                       int independent) {
    static uint32_t table[2 << HASHLOG];
    size_t op = 0, nblocks = (n + 65535) / 65536;
    if (cap < 7 + nblocks * 4 + n + 4) return 0;
    wr32(dst, MAGIC_FRAME);
    dst[4] = independent ? 0x60 : 0x40;
    dst[5] = 0x40;
    dst[6] = (uint8_t)(orc_xxh32(dst + 4, 2, 0) >> 8);
    op = 7;
    int accel = level < 0 ? -level + 1 : 1;                         /* lz4frame.c:768,779 */
    memset(table, 0, sizeof table);
    for (size_t off = 0; off < n; off += 65536) {
        size_t len = n - off < 65536 ? n - off : 65536;
        size_t c;
        if (independent) {
            memset(table, 0, sizeof table);
            g_u16_mode = 1;                                         /* len <= 64 KB < LZ4_64Klimit */
            c = orc_lz4_block_encode(src + off, 0, len, dst + op + 4, len - 1, table, accel);
            g_u16_mode = 0;
        } else {
            c = orc_lz4_block_encode(src, off, len, dst + op + 4, len - 1, table, accel);
        }
        if (c == 0) {
            wr32(dst + op, (uint32_t)len | 0x80000000u);
            memcpy(dst + op + 4, src + off, len);
            c = len;
        } else {
            wr32(dst + op, (uint32_t)c);
        }
        op += 4 + c;
    }
    wr32(dst + op, 0);
    return op + 4;
}
