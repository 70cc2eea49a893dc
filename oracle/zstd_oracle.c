/* oracle/zstd_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Zstandard decoder restated from the published format (externals/zstd/doc/zstd_compression_format.md)
 * with the acceptance rules of zstd 1.5.0 as ZPack drives it (ZSTD_decompressDCtx, lib/zpack_read.c:380):
 *   multi-frame loop + skippable frames   externals/zstd/lib/decompress/zstd_decompress.c:907-996
 *   frame header                          zstd_decompress.c:419-493
 *   frame / block loop                    zstd_decompress.c:819-905, zstd_decompress_block.c:56-70
 *   literals section                      zstd_decompress_block.c:79-235
 *   Huffman weights + table, X1 decode    common/entropy_common.c:264-329, huf_decompress.c:147-441
 *   FSE table description + table build   common/entropy_common.c:64-210, zstd_decompress_block.c:368-485
 *   sequence header / decode / execute    zstd_decompress_block.c:577-654, 937-1039, 804-893, 1090-1210
 * Parity is PINNED by tests/test_oracle.py: golden archive tests/workdir/archive_zstd.zpk, the frames in
 * tests/golden/zstd_cases.npz written by the unmodified reference at levels 1-19, and differential runs
 * against oracle/_ref (ZSTD_decompress) on the synthetic corpus, truncations and bit flips.
 * Deliberate simplification: a backward bitstream that is read past its beginning is an error here at once;
 * the library keeps decoding garbage and fails later (final check zstd_decompress_block.c:1195 or the digest).
 */
#include "oracle.h"
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define ZERR (-1)
#define BLOCK_MAX (128u * 1024u)

static u32 rd16(const u8 *p) { return (u32)p[0] | (u32)p[1] << 8; }
static u32 rd24(const u8 *p) { return rd16(p) | (u32)p[2] << 16; }
static u32 rd32(const u8 *p) { return rd16(p) | rd16(p + 2) << 16; }
static u64 rd64(const u8 *p) { return (u64)rd32(p) | (u64)rd32(p + 4) << 32; }
static int highbit(u32 v) { int r = 0; while (v >>= 1) ++r; return r; }

/* bits [pos, pos+n) of a little-endian byte string, n <= 56; bits outside [0, 8*len) read as zero */
static u64 bits_at(const u8 *src, size_t len, int64_t pos, int n) {
    u64 v = 0;
    for (int i = 0; i < n; ++i) {
        int64_t b = pos + i;
        if (b >= 0 && (u64)b < (u64)len * 8 && (src[b >> 3] >> (b & 7) & 1)) v |= (u64)1 << i;
    }
    return v;
}

/* ---- backward bitstream (bitstream.h:277-322): `left` = unread bits below the end marker */
typedef struct { const u8 *src; size_t len; int64_t left; } rbits;
static int rb_init(rbits *b, const u8 *src, size_t len) {
    if (len == 0 || src[len - 1] == 0) return ZERR;
    b->src = src; b->len = len;
    b->left = (int64_t)(len - 1) * 8 + highbit(src[len - 1]);
    return 0;
}
static u64 rb_read(rbits *b, int n) { b->left -= n; return bits_at(b->src, b->len, b->left, n); }

/* ---- XXH64 (zstd's optional frame checksum; externals/zstd/lib/common/xxhash.c) */
static u64 rotl64(u64 v, int r) { return (v << r) | (v >> (64 - r)); }
static u64 xxh64(const u8 *p, size_t len) {
    const u64 P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull,
              P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
    const u8 *end = p + len;
    u64 h;
    if (len >= 32) {
        u64 v[4] = {P1 + P2, P2, 0, 0 - P1};
        do {
            for (int i = 0; i < 4; ++i) { v[i] = rotl64(v[i] + rd64(p) * P2, 31) * P1; p += 8; }
        } while (p + 32 <= end);
        h = rotl64(v[0], 1) + rotl64(v[1], 7) + rotl64(v[2], 12) + rotl64(v[3], 18);
        for (int i = 0; i < 4; ++i) h = (h ^ (rotl64(v[i] * P2, 31) * P1)) * P1 + P4;
    } else {
        h = P5;
    }
    h += len;
    for (; p + 8 <= end; p += 8) h = rotl64(h ^ (rotl64(rd64(p) * P2, 31) * P1), 27) * P1 + P4;
    if (p + 4 <= end) { h = rotl64(h ^ ((u64)rd32(p) * P1), 23) * P2 + P3; p += 4; }
    for (; p < end; ++p) h = rotl64(h ^ (*p * P5), 11) * P1;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* ---- FSE */
typedef struct { u16 base; u8 sym; u8 nbits; } fse_cell;
typedef struct { fse_cell cell[512]; int log; } fse_table;

/* normalized counts from a table description (entropy_common.c:64-210); returns bytes used or ZERR */
static int fse_read_ncount(const u8 *src, size_t len, int16_t *norm, int *max_sym, int *log_out, int max_log_abs) {
    if (len == 0) return ZERR;
    for (int s = 0; s <= *max_sym; ++s) norm[s] = 0;      /* symbols skipped by repeat flags stay at zero */
    int64_t bp = 0;
    int log = (int)bits_at(src, len, bp, 4) + 5; bp += 4;
    if (log > max_log_abs) return ZERR;
    int remaining = (1 << log) + 1, threshold = 1 << log, nbits = log + 1, sym = 0, prev0 = 0;
    while (remaining > 1 && sym <= *max_sym) {
        if (prev0) {
            for (;;) {
                int rep = (int)bits_at(src, len, bp, 2); bp += 2;
                sym += rep;
                if (rep != 3) break;
            }
            if (sym > *max_sym) break;     /* entropy_common.c: charnum past the limit ends the loop */
        }
        int max = (2 * threshold - 1) - remaining, count;
        u32 v = (u32)bits_at(src, len, bp, nbits);
        if ((int)(v & (u32)(threshold - 1)) < max) { count = (int)(v & (u32)(threshold - 1)); bp += nbits - 1; }
        else { count = (int)(v & (u32)(2 * threshold - 1)); if (count >= threshold) count -= max; bp += nbits; }
        --count;                                     /* -1: "less than one" */
        remaining -= count < 0 ? -count : count;
        norm[sym++] = (int16_t)count;
        prev0 = !count;
        while (remaining < threshold) { --nbits; threshold >>= 1; }
    }
    if (remaining != 1 || sym > *max_sym + 1) return ZERR;
    size_t used = (size_t)((bp + 7) >> 3);
    if (used > len) return ZERR;
    for (int s = sym; s <= *max_sym; ++s) norm[s] = 0;
    *max_sym = sym - 1;
    *log_out = log;
    return (int)used;
}

/* decode table from normalized counts (zstd_decompress_block.c:368-485 / fse_decompress.c:71-130) */
static int fse_build(fse_table *t, const int16_t *norm, int max_sym, int log) {
    int size = 1 << log, high = size - 1;
    u16 next[256];
    for (int s = 0; s <= max_sym; ++s) {
        if (norm[s] == -1) { t->cell[high--].sym = (u8)s; next[s] = 1; }
        else next[s] = (u16)norm[s];
    }
    int step = (size >> 1) + (size >> 3) + 3, mask = size - 1, pos = 0;
    for (int s = 0; s <= max_sym; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            t->cell[pos].sym = (u8)s;
            do pos = (pos + step) & mask; while (pos > high);
        }
    if (pos != 0) return ZERR;
    for (int u = 0; u < size; ++u) {
        u16 n = next[t->cell[u].sym]++;
        int nb = log - highbit(n);
        t->cell[u].nbits = (u8)nb;
        t->cell[u].base = (u16)((n << nb) - size);
    }
    t->log = log;
    return 0;
}
static void fse_rle(fse_table *t, int sym) { t->cell[0].sym = (u8)sym; t->cell[0].nbits = 0; t->cell[0].base = 0; t->log = 0; }

/* ---- Huffman (X1 semantics: any valid stream decodes identically with the library's X2 tables) */
typedef struct { u8 sym[4096]; u8 len[4096]; int log; } huf_table;

static int huf_build(huf_table *h, const u8 *w, int nsym) {   /* weights for symbols 0..nsym-2, last implied */
    u32 total = 0;
    for (int i = 0; i < nsym - 1; ++i) { if (w[i] > 12) return ZERR; if (w[i]) total += 1u << (w[i] - 1); }
    if (total == 0) return ZERR;
    int log = highbit(total) + 1;
    if (log > 12) return ZERR;
    u32 rest = (1u << log) - total;
    if (rest & (rest - 1)) return ZERR;                   /* entropy_common.c:307-313: must be a power of two */
    u8 weights[256];
    memcpy(weights, w, (size_t)(nsym - 1));
    weights[nsym - 1] = (u8)(highbit(rest) + 1);
    u32 rank[16] = {0};
    for (int i = 0; i < nsym; ++i) rank[weights[i]]++;
    if (rank[1] < 2 || (rank[1] & 1)) return ZERR;        /* entropy_common.c:321 */
    u32 start[16], acc = 0;
    for (int r = 1; r <= log; ++r) { start[r] = acc; acc += rank[r] << (r - 1); }
    for (int s = 0; s < nsym; ++s) {
        int wt = weights[s];
        if (!wt) continue;
        u32 n = 1u << (wt - 1);
        for (u32 k = 0; k < n; ++k) { h->sym[start[wt] + k] = (u8)s; h->len[start[wt] + k] = (u8)(log + 1 - wt); }
        start[wt] += n;
    }
    h->log = log;
    return 0;
}

/* tree description (entropy_common.c:264-329); returns bytes used or ZERR */
static int huf_read_tree(huf_table *h, const u8 *src, size_t len) {
    if (len == 0) return ZERR;
    u8 w[256];
    int nw;
    size_t used;
    u32 hb = src[0];
    if (hb >= 128) {                                       /* 4-bit weights, direct */
        nw = (int)hb - 127;
        used = 1 + (size_t)(nw + 1) / 2;
        if (used > len) return ZERR;
        for (int i = 0; i < nw; ++i) w[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    } else {                                               /* FSE-compressed weights, two interleaved states */
        used = 1 + hb;
        if (used > len) return ZERR;
        int16_t norm[256];
        int max_sym = 255, log;
        int hs = fse_read_ncount(src + 1, hb, norm, &max_sym, &log, 15);
        if (hs < 0 || log > 6) return ZERR;
        fse_table t;
        if (fse_build(&t, norm, max_sym, log)) return ZERR;
        rbits b;
        if (rb_init(&b, src + 1 + hs, hb - (size_t)hs)) return ZERR;
        u32 s1 = (u32)rb_read(&b, log), s2 = (u32)rb_read(&b, log);
        if (b.left < 0) return ZERR;
        nw = 0;
        for (;;) {                                         /* fse_decompress.c tail loop: ends on overflow */
            if (nw > 253) return ZERR;
            w[nw++] = t.cell[s1].sym;
            s1 = t.cell[s1].base + (u32)rb_read(&b, t.cell[s1].nbits);
            if (b.left < 0) { w[nw++] = t.cell[s2].sym; break; }
            if (nw > 253) return ZERR;
            w[nw++] = t.cell[s2].sym;
            s2 = t.cell[s2].base + (u32)rb_read(&b, t.cell[s2].nbits);
            if (b.left < 0) { w[nw++] = t.cell[s1].sym; break; }
        }
    }
    if (huf_build(h, w, nw + 1)) return ZERR;
    return (int)used;
}

static int huf_stream(const huf_table *h, const u8 *src, size_t len, u8 *dst, size_t n) {
    rbits b;
    if (rb_init(&b, src, len)) return ZERR;
    for (size_t i = 0; i < n; ++i) {
        u32 idx = (u32)bits_at(b.src, b.len, b.left - h->log, h->log);
        dst[i] = h->sym[idx];
        b.left -= h->len[idx];
    }
    return b.left == 0 ? 0 : ZERR;                          /* huf_decompress.c: BIT_endOfDStream */
}

/* ---- frame state */
typedef struct {
    huf_table huf; int huf_valid;
    fse_table ll, of, ml; int seq_valid;
    u64 rep[3];
    u8 lit[BLOCK_MAX + 32];
} zctx;

static const int16_t LL_DEF[36] = {4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1};
static const int16_t ML_DEF[53] = {1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1};
static const int16_t OF_DEF[29] = {1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1};
static const u32 LL_BASE[36] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,1024,2048,4096,8192,16384,32768,65536};
static const u8 LL_BITS[36] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16};
static const u32 ML_BASE[53] = {3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,32771,65539};
static const u8 ML_BITS[53] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16};

/* one of the three sequence tables (zstd_decompress_block.c:529-575); returns bytes used or ZERR */
static int seq_table(fse_table *t, int mode, const u8 *src, size_t len, int max_sym, int max_log,
                     const int16_t *def, int def_n, int def_log, int have_prev) {
    switch (mode) {
    case 0: return fse_build(t, def, def_n - 1, def_log) ? ZERR : 0;
    case 1: if (len == 0 || src[0] > max_sym) return ZERR; fse_rle(t, src[0]); return 1;
    case 2: {
        int16_t norm[64];
        int ms = max_sym, log;
        int hs = fse_read_ncount(src, len, norm, &ms, &log, 15);
        if (hs < 0 || log > max_log) return ZERR;
        if (fse_build(t, norm, ms, log)) return ZERR;
        return hs;
    }
    default: return have_prev ? 0 : ZERR;
    }
}

/* literals section (zstd_decompress_block.c:79-235); returns bytes used or ZERR; literals land in z->lit */
static int literals(zctx *z, const u8 *src, size_t len, size_t *lit_size) {
    if (len < 3) return ZERR;                                             /* MIN_CBLOCK_SIZE */
    int type = src[0] & 3, fmt = (src[0] >> 2) & 3;
    if (type == 0 || type == 1) {                                         /* raw / RLE */
        size_t lh, ls;
        if (fmt == 0 || fmt == 2) { lh = 1; ls = src[0] >> 3; }
        else if (fmt == 1) { lh = 2; ls = rd16(src) >> 4; }
        else { lh = 3; ls = rd24(src) >> 4; }
        if (type == 0) {
            if (lh + ls > len) return ZERR;
            memcpy(z->lit, src + lh, ls);
            *lit_size = ls;
            return (int)(lh + ls);
        }
        if (fmt == 3 && len < 4) return ZERR;
        if (ls > BLOCK_MAX) return ZERR;
        memset(z->lit, src[lh], ls);
        *lit_size = ls;
        return (int)lh + 1;
    }
    if (type == 3 && !z->huf_valid) return ZERR;                          /* treeless without a previous tree */
    if (len < 5) return ZERR;
    size_t lh, ls, cs;
    int single = 0;
    u32 lhc = rd32(src);
    if (fmt <= 1) { single = !fmt; lh = 3; ls = (lhc >> 4) & 0x3FF; cs = (lhc >> 14) & 0x3FF; }
    else if (fmt == 2) { lh = 4; ls = (lhc >> 4) & 0x3FFF; cs = lhc >> 18; }
    else { lh = 5; ls = (lhc >> 4) & 0x3FFFF; cs = (lhc >> 22) + ((size_t)src[4] << 10); }
    if (ls > BLOCK_MAX || cs + lh > len) return ZERR;
    const u8 *p = src + lh;
    size_t left = cs;
    if (type == 2) {
        int hs = huf_read_tree(&z->huf, p, left);
        if (hs < 0) return ZERR;
        if ((size_t)hs >= left) return ZERR;                              /* huf_decompress.c: hSize >= cSrcSize */
        p += hs; left -= (size_t)hs;
        z->huf_valid = 1;
    }
    if (single) {
        if (huf_stream(&z->huf, p, left, z->lit, ls)) return ZERR;
    } else {
        if (left < 10) return ZERR;                                       /* jump table + 1 byte per stream */
        size_t s1 = rd16(p), s2 = rd16(p + 2), s3 = rd16(p + 4);
        if (6 + s1 + s2 + s3 > left) return ZERR;
        size_t s4 = left - 6 - s1 - s2 - s3, seg = (ls + 3) / 4;
        if (3 * seg > ls) return ZERR;                                    /* huf_decompress.c:386: opStart4 > oend */
        const u8 *q = p + 6;
        if (huf_stream(&z->huf, q, s1, z->lit, seg) || huf_stream(&z->huf, q + s1, s2, z->lit + seg, seg) ||
            huf_stream(&z->huf, q + s1 + s2, s3, z->lit + 2 * seg, seg) ||
            huf_stream(&z->huf, q + s1 + s2 + s3, s4, z->lit + 3 * seg, ls - 3 * seg))
            return ZERR;
    }
    *lit_size = ls;
    return (int)(lh + cs);
}

/* compressed block (zstd_decompress_block.c:1456-1525, 1090-1210); returns decoded size, ZERR, or -2 = output full */
static long block(zctx *z, const u8 *src, size_t len, u8 *dst_base, size_t frame_start, size_t op, size_t cap) {
    if (len >= BLOCK_MAX) return ZERR;
    size_t lit_size = 0;
    int used = literals(z, src, len, &lit_size);
    if (used < 0) return ZERR;
    src += used; len -= (size_t)used;
    /* sequences header (zstd_decompress_block.c:577-654) */
    if (len < 1) return ZERR;
    const u8 *ip = src, *iend = src + len;
    int nseq = *ip++;
    size_t start = op, lit_pos = 0;
    if (nseq == 0) {
        if (len != 1) return ZERR;
    } else {
        if (nseq > 0x7F) {
            if (nseq == 0xFF) { if (ip + 2 > iend) return ZERR; nseq = (int)rd16(ip) + 0x7F00; ip += 2; }
            else { if (ip >= iend) return ZERR; nseq = ((nseq - 0x80) << 8) + *ip++; }
        }
        if (ip + 1 > iend) return ZERR;
        int modes = *ip++;
        int n;
        if ((n = seq_table(&z->ll, modes >> 6, ip, (size_t)(iend - ip), 35, 9, LL_DEF, 36, 6, z->seq_valid)) < 0) return ZERR;
        ip += n;
        if ((n = seq_table(&z->of, (modes >> 4) & 3, ip, (size_t)(iend - ip), 31, 8, OF_DEF, 29, 5, z->seq_valid)) < 0) return ZERR;
        ip += n;
        if ((n = seq_table(&z->ml, (modes >> 2) & 3, ip, (size_t)(iend - ip), 52, 9, ML_DEF, 53, 6, z->seq_valid)) < 0) return ZERR;
        ip += n;
        z->seq_valid = 1;
        rbits b;
        if (rb_init(&b, ip, (size_t)(iend - ip))) return ZERR;
        u32 sl = (u32)rb_read(&b, z->ll.log), so = (u32)rb_read(&b, z->of.log), sm = (u32)rb_read(&b, z->ml.log);
        if (b.left < 0) return ZERR;
        for (int i = 0; i < nseq; ++i) {
            int oc = z->of.cell[so].sym, mc = z->ml.cell[sm].sym, lc = z->ll.cell[sl].sym;
            u64 ofv = oc ? ((u64)1 << oc) + rb_read(&b, oc) : 1;       /* zstd_compression_format.md:  offset_value */
            u64 ml = ML_BASE[mc] + rb_read(&b, ML_BITS[mc]);
            u64 ll = LL_BASE[lc] + rb_read(&b, LL_BITS[lc]);
            if (b.left < 0) return ZERR;
            u64 off;
            if (ofv > 3) { off = ofv - 3; z->rep[2] = z->rep[1]; z->rep[1] = z->rep[0]; z->rep[0] = off; }
            else {                                                       /* repeat offsets (block.c:971-987) */
                u64 idx = ofv - 1 + (ll == 0);
                if (idx == 0) off = z->rep[0];
                else {
                    off = idx == 3 ? z->rep[0] - 1 : z->rep[idx];
                    if (!off) off = 1;                                   /* library: corrupted input, forced to 1 */
                    if (idx != 1) z->rep[2] = z->rep[1];
                    z->rep[1] = z->rep[0];
                    z->rep[0] = off;
                }
            }
            if (ll > lit_size - lit_pos) return ZERR;
            if (ll + ml > cap - op) return -2;
            memcpy(dst_base + op, z->lit + lit_pos, (size_t)ll);
            op += (size_t)ll; lit_pos += (size_t)ll;
            if (off > op - frame_start) return ZERR;
            for (u64 k = 0; k < ml; ++k) dst_base[op + k] = dst_base[op + k - off];
            op += (size_t)ml;
            if (i + 1 < nseq) {                                          /* state update order: LL, ML, OF */
                sl = z->ll.cell[sl].base + (u32)rb_read(&b, z->ll.cell[sl].nbits);
                sm = z->ml.cell[sm].base + (u32)rb_read(&b, z->ml.cell[sm].nbits);
                so = z->of.cell[so].base + (u32)rb_read(&b, z->of.cell[so].nbits);
                if (b.left < 0) return ZERR;
            }
        }
        /* the library updates the states once more and then wants every bit consumed (block.c:1195) */
        int64_t tail = z->ll.cell[sl].nbits + z->ml.cell[sm].nbits + z->of.cell[so].nbits;
        if (b.left > tail) return ZERR;
    }
    size_t rest = lit_size - lit_pos;
    if (rest > cap - op) return -2;
    memcpy(dst_base + op, z->lit + lit_pos, rest);
    op += rest;
    return (long)(op - start);
}

static zctx g_z;   /* test infrastructure: single-threaded callers use the shared context, others pass their own */

int orc_zstd_decode_ctx(void *ctx, const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len) {
    zctx *z = (zctx *)ctx;
    size_t ip = 0, op = 0;
    *out_len = 0;
    while (n - ip >= 4) {                                                  /* zstd_decompress.c:907-996 */
        u32 magic = rd32(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (n - ip < 8) return ORC_DECOMPRESS_FAILED;
            u64 skip = (u64)rd32(src + ip + 4) + 8;
            if (skip > n - ip) return ORC_DECOMPRESS_FAILED;
            ip += (size_t)skip;
            continue;
        }
        if (magic != 0xFD2FB528u) return ORC_DECOMPRESS_FAILED;
        /* frame header (zstd_decompress.c:419-493) */
        if (n - ip < 5 + 3) return ORC_DECOMPRESS_FAILED;                  /* FRAMEHEADERSIZE_MIN 6?: magic + FHD + >=1, + block header */
        u32 fhd = src[ip + 4];
        u32 did_code = fhd & 3, csum = (fhd >> 2) & 1, single = (fhd >> 5) & 1, fcs_code = fhd >> 6;
        if (fhd & 8) return ORC_DECOMPRESS_FAILED;                          /* reserved bit */
        static const u8 did_len[4] = {0, 1, 2, 4}, fcs_len[4] = {0, 2, 4, 8};
        size_t hsize = 5 + (single ? 0 : 1) + did_len[did_code] + fcs_len[fcs_code] + ((single && !fcs_code) ? 1 : 0);
        if (n - ip < hsize + 3) return ORC_DECOMPRESS_FAILED;
        size_t p = ip + 5;
        u64 window = 0;
        if (!single) {
            u32 wd = src[p++];
            u32 wlog = (wd >> 3) + 10;
            if (wlog > 31) return ORC_DECOMPRESS_FAILED;                    /* ZSTD_WINDOWLOG_MAX on 64-bit */
            window = ((u64)1 << wlog) + (((u64)1 << wlog) >> 3) * (wd & 7);
        }
        u32 dict_id = 0;
        for (u32 k = 0; k < did_len[did_code]; ++k) dict_id |= (u32)src[p++] << (8 * k);
        u64 fcs = 0;
        int have_fcs = 1;
        if (fcs_code == 0) { if (single) fcs = src[p++]; else have_fcs = 0; }
        else if (fcs_code == 1) { fcs = rd16(src + p) + 256; p += 2; }
        else if (fcs_code == 2) { fcs = rd32(src + p); p += 4; }
        else { fcs = rd64(src + p); p += 8; }
        if (single) window = fcs;
        if (dict_id) return ORC_DECOMPRESS_FAILED;                          /* no dictionary is ever loaded */
        (void)window;
        ip += hsize;
        z->huf_valid = 0; z->seq_valid = 0;
        z->rep[0] = 1; z->rep[1] = 4; z->rep[2] = 8;
        size_t frame_start = op;
        for (;;) {                                                          /* zstd_decompress.c:850-889 */
            if (n - ip < 3) return ORC_DECOMPRESS_FAILED;
            u32 bh = rd24(src + ip);
            ip += 3;
            u32 last = bh & 1, type = (bh >> 1) & 3, bsz = bh >> 3;
            if (type == 3) return ORC_DECOMPRESS_FAILED;
            size_t csz = type == 1 ? 1 : bsz;
            if (csz > n - ip) return ORC_DECOMPRESS_FAILED;
            if (type == 0) {
                if (bsz > cap - op) return ORC_DECOMPRESS_FAILED;           /* dstSize_tooSmall is an error code too */
                memcpy(dst + op, src + ip, bsz);
                op += bsz;
            } else if (type == 1) {
                if (bsz > cap - op) return ORC_DECOMPRESS_FAILED;
                memset(dst + op, src[ip], bsz);
                op += bsz;
            } else {
                long got = block(z, src + ip, csz, dst, frame_start, op, cap);
                if (got < 0) return ORC_DECOMPRESS_FAILED;
                op += (size_t)got;
            }
            ip += csz;
            if (last) break;
        }
        if (have_fcs && (u64)(op - frame_start) != fcs) return ORC_DECOMPRESS_FAILED;
        if (csum) {
            if (n - ip < 4) return ORC_DECOMPRESS_FAILED;
            if (rd32(src + ip) != (u32)xxh64(dst + frame_start, op - frame_start)) return ORC_DECOMPRESS_FAILED;
            ip += 4;
        }
    }
    if (ip != n) return ORC_DECOMPRESS_FAILED;                              /* "input not entirely consumed" */
    *out_len = op;
    return ORC_OK;
}

int orc_zstd_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len) {
    return orc_zstd_decode_ctx(&g_z, src, n, dst, cap, out_len);
}
size_t orc_zstd_ctx_size(void) { return sizeof(zctx); }
