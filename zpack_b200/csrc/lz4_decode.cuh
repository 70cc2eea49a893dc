// lz4_decode.cuh — LZ4 frame + block decoder for a lane group (kernel K1 of SURVEY.md §2.2).
//
// Replaces the LZ4 arm of zpack_read_file (/root/reference/lib/zpack_read.c:396-453), i.e. the
// loop around LZ4F_decompress (externals/lz4/lib/lz4frame.c:1384-1879) and the block decoder
// LZ4_decompress_safe_usingDict (externals/lz4/lib/lz4.c:1737-2165, :2404-2417).
//
// One group decodes one dependency chain front to back: every frame of the entry, every block
// of a frame in order, writing straight into the entry's final place in HBM so that a linked
// block's 64 KB window is simply "earlier output" (the reference's prefix mode, lz4.c:2408-2414).
// Control state is replicated per lane (see common.cuh); token / length / offset bytes are
// broadcast loads that hit L1, literal and match bytes move one byte per lane for short runs and
// as aligned 16-byte stores for long ones.  Overlapping matches (offset < length, lz4.c:2122)
// never read a byte produced by the same sequence: the source is periodic with period `offset`,
// so lane i reads  dst[op - offset + (i mod offset)], which was complete before the sequence.
//
// Untrusted input: every read is bounded by the entry's compressed range and every write by
// dst_cap; the checks mirror lz4frame.c:1156-1186,1511 and lz4.c:1811-1812,1853,2031,2139.
#pragma once
#include "common.cuh"
#include "xxh3.cuh"

#define LZ4F_MAGIC      0x184D2204u
#define LZ4F_MAGIC_SKIP 0x184D2A50u

// status values are enum zpack_result numbers (include/zpack_b200.h ZPB_ST_*)
#define ST_OK 0
#define ST_TOO_SMALL 12
#define ST_DECODE_FAILED 13
#define ST_HASH_MISMATCH 15
#define ST_OFFSET_INVALID 16
#define ST_INCOMPLETE 17
#define ST_SIZE_INVALID 18
#define ST_METHOD_INVALID 19
#define ST_NOT_AVAILABLE 24

// XXH32 (lz4's private xxhash.c) — only the frame-header check byte and the optional
// block/content checksums use it; evaluated redundantly by every lane (tiny or rare).
__device__ __noinline__ u32 xxh32_dev(const u8 *p, u64 len, u32 seed) {
    const u32 P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, P4 = 0x27D4EB2Fu, P5 = 0x165667B1u;
    const u8 *end = p + len;
    u32 h;
    if (len >= 16) {
        u32 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = __funnelshift_l(v1 + ld32u(p) * P2, v1 + ld32u(p) * P2, 13) * P1;
            v2 = __funnelshift_l(v2 + ld32u(p + 4) * P2, v2 + ld32u(p + 4) * P2, 13) * P1;
            v3 = __funnelshift_l(v3 + ld32u(p + 8) * P2, v3 + ld32u(p + 8) * P2, 13) * P1;
            v4 = __funnelshift_l(v4 + ld32u(p + 12) * P2, v4 + ld32u(p + 12) * P2, 13) * P1;
            p += 16;
        } while (p + 16 <= end);
        h = __funnelshift_l(v1, v1, 1) + __funnelshift_l(v2, v2, 7) + __funnelshift_l(v3, v3, 12) +
            __funnelshift_l(v4, v4, 18);
    } else {
        h = seed + P5;
    }
    h += (u32)len;
    while (p + 4 <= end) { u32 t = h + ld32u(p) * P3; h = __funnelshift_l(t, t, 17) * P4; p += 4; }
    while (p < end) { u32 t = h + *p * P5; h = __funnelshift_l(t, t, 11) * P1; ++p; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

// Decode one LZ4 block.  bsrc[0..bsz) -> bdst[0..), at most `cap` bytes, `prefix` bytes of
// window available immediately before bdst.  `hard_cap` is the frame's maxBlockSize (the
// reference decodes every block with that capacity, lz4frame.c:1683).  Returns produced bytes
// or a negative status.  `hs` / `hbase_off` let the entry digest trail the output front.
template <int G>
ZPB_DEVINL int lz4_block_decode(const Group<G> &g, const u8 *bsrc, u32 bsz, u8 *bdst, u32 cap,
                                u32 hard_cap, u64 prefix, Xxh3Stream<G> &hs, u64 out_before) {
    u32 ip = 0, op = 0;
    const int l = g.l;
    for (;;) {
        u32 token = ld8(bsrc + ip++);
        u32 lit = token >> 4;
        if (lit == 15) {
            u32 b;
            do {
                if (ip >= bsz) return -ST_DECODE_FAILED;
                b = ld8(bsrc + ip++);
                lit += b;
            } while (b == 255 && lit < 0x40000000u);
            if (b == 255) return -ST_DECODE_FAILED;
        }
        if (lit > bsz - ip) return -ST_DECODE_FAILED;
        if (lit > cap - op) return cap < hard_cap ? -ST_TOO_SMALL : -ST_DECODE_FAILED;
        bool last = (ip + lit == bsz);
        if (!last && ip + lit + 8 > bsz) return -ST_DECODE_FAILED;  // lz4.c:2055-2077
        group_copy<G>(g, bdst + op, bsrc + ip, lit);
        ip += lit; op += lit;
        if (last) return (int)op;

        u32 off = ld16u(bsrc + ip);
        ip += 2;
        u32 ml = token & 15;
        if (ml == 15) {
            u32 b;
            do {
                if (ip >= bsz) return -ST_DECODE_FAILED;
                b = ld8(bsrc + ip++);
                ml += b;
            } while (b == 255 && ml < 0x40000000u);
            if (b == 255) return -ST_DECODE_FAILED;
        }
        ml += 4;
        if (off == 0 || (u64)off > (u64)op + prefix) return -ST_DECODE_FAILED;  // lz4.c:2093
        if ((u64)op + ml + 5 > hard_cap) return -ST_DECODE_FAILED;              // lz4.c:2139
        if (ml > cap - op) return -ST_TOO_SMALL;
        if (ip >= bsz) return -ST_DECODE_FAILED;

        g.sync();  // literals (and everything earlier) visible before the match reads them
        u8 *d = bdst + op;
        if (off >= ml) {
            group_copy<G>(g, d, d - off, ml);
        } else if (off == 1) {
            u32 v = ld8(d - 1);
            for (u32 i = l; i < ml; i += G) d[i] = (u8)v;
        } else {
            // periodic source: byte i comes from (i mod off) inside the `off` bytes before op
            u32 r = (u32)l % off, step = (u32)G % off;
            const u8 *s = d - off;
            for (u32 i = l; i < ml; i += G) {
                d[i] = s[r];
                r += step;
                if (r >= off) r -= off;
            }
        }
        op += ml;
        hs.advance(out_before + op, g);
    }
}

// Decode every frame of one entry (the loop at lib/zpack_read.c:414-450).  Returns a status;
// *produced gets the number of bytes written to dst.
template <int G>
ZPB_DEVINL int lz4f_decode_entry(const Group<G> &g, const u8 *src, u64 n, u8 *dst, u64 cap,
                                 Xxh3Stream<G> &hs, u64 *produced) {
    u64 ip = 0, op = 0;
    int rc = ST_OK;
#define NEED_MORE() (op < cap ? ST_INCOMPLETE : ST_TOO_SMALL)
    while (ip < n && op < cap) {
        if (n - ip < 7) { rc = NEED_MORE(); break; }
        u32 magic = ld32u(src + ip);
        if ((magic & 0xFFFFFFF0u) == LZ4F_MAGIC_SKIP) {  // lz4frame.c:1125-1136
            if (n - ip < 8) { rc = NEED_MORE(); break; }
            u64 skip = ld32u(src + ip + 4);
            if (skip > n - ip - 8) { rc = NEED_MORE(); break; }
            ip += 8 + skip;
            continue;
        }
        if (magic != LZ4F_MAGIC) { rc = ST_DECODE_FAILED; break; }
        u32 flg = ld8(src + ip + 4);
        if (((flg >> 1) & 1) || ((flg >> 6) & 3) != 1) { rc = ST_DECODE_FAILED; break; }
        bool indep = (flg >> 5) & 1, bsum = (flg >> 4) & 1, has_size = (flg >> 3) & 1,
             csum = (flg >> 2) & 1, has_dict = flg & 1;
        u32 hsize = 7 + (has_size ? 8 : 0) + (has_dict ? 4 : 0);
        if (n - ip < hsize) { rc = NEED_MORE(); break; }
        u32 bd = ld8(src + ip + 5);
        if ((bd >> 7) || ((bd >> 4) & 7) < 4 || (bd & 15)) { rc = ST_DECODE_FAILED; break; }
        u32 max_block = 1u << (8 + 2 * ((bd >> 4) & 7));
        if (((xxh32_dev(src + ip + 4, hsize - 5, 0) >> 8) & 0xFF) != ld8(src + ip + hsize - 1)) {
            rc = ST_DECODE_FAILED; break;  // lz4frame.c:294-298,1184-1186
        }
        u64 remaining = has_size ? ld64u(src + ip + 6) : 0;
        ip += hsize;
        u64 frame_start = op;
        bool done = false;
        while (!done) {
            if (n - ip < 4) { rc = NEED_MORE(); break; }
            u32 bh = ld32u(src + ip);
            ip += 4;
            if (bh == 0) {  // EndMark
                if (has_size && remaining != 0) { rc = ST_DECODE_FAILED; break; }
                if (csum) {
                    if (n - ip < 4) { rc = NEED_MORE(); break; }
                    g.sync();
                    if (ld32u(src + ip) != xxh32_dev(dst + frame_start, op - frame_start, 0)) {
                        rc = ST_DECODE_FAILED; break;
                    }
                    ip += 4;
                }
                done = true;
                break;
            }
            u32 bsz = bh & 0x7FFFFFFFu;
            if (bsz > max_block) { rc = ST_DECODE_FAILED; break; }
            if (bh & 0x80000000u) {  // stored block: lz4frame.c:1534-1572
                u64 take = bsz;
                if (take > n - ip) take = n - ip;
                if (take > cap - op) take = cap - op;
                group_copy<G>(g, dst + op, src + ip, (u32)take);
                if (bsum && take == bsz) {
                    if (n - ip - bsz < 4) { op += take; rc = NEED_MORE(); break; }
                    if (ld32u(src + ip + bsz) != xxh32_dev(src + ip, bsz, 0)) { rc = ST_DECODE_FAILED; break; }
                }
                op += take; remaining -= take;
                hs.advance(op, g);
                if (take < bsz) { rc = NEED_MORE(); break; }
                ip += (u64)bsz + (bsum ? 4 : 0);
                continue;
            }
            if (op == cap) { rc = ST_TOO_SMALL; break; }  // lz4frame.c:1527-1530
            if (n - ip < (u64)bsz + (bsum ? 4 : 0)) { rc = NEED_MORE(); break; }
            if (bsz == 0) { rc = ST_DECODE_FAILED; break; }
            if (bsum && ld32u(src + ip + bsz) != xxh32_dev(src + ip, bsz, 0)) { rc = ST_DECODE_FAILED; break; }
            u64 room = cap - op;
            u32 bcap = room >= max_block ? max_block : (u32)room;
            int got = lz4_block_decode<G>(g, src + ip, bsz, dst + op, bcap, max_block,
                                          indep ? 0 : op - frame_start, hs, op);
            if (got < 0) { rc = -got; break; }
            op += (u32)got; remaining -= (u32)got;
            hs.advance(op, g);
            ip += (u64)bsz + (bsum ? 4 : 0);
        }
        if (rc != ST_OK) break;
    }
#undef NEED_MORE
    *produced = op;
    return rc;
}
