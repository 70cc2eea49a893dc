// pack_kernel.cuh — LZ4 frame encoder + fused XXH3-64 of the input: one warp per file.
//
// Replaces zpack_compress_file's LZ4 arm (/root/reference/lib/zpack_write.c:192-214:
// LZ4F_compressBegin / Update / End, externals/lz4/lib/lz4frame.c:598-1022) and the second pass
// zpack_add_written_file_entry makes for the digest (lib/zpack_write.c:256).  The block compressor
// restates the greedy single-probe hash-table matcher of LZ4_compress_generic_validated
// (externals/lz4/lib/lz4.c:851-1240) for a warp:
//
//   * 32 consecutive positions are probed per step, one per lane: 4-byte word, multiplicative hash
//     into a 4096-entry table in shared memory (entry = position | 16-bit tag of the word, so the
//     verifying gather from the input is only issued by lanes whose tag already agrees);
//   * candidates closer than 32 bytes (runs, short periods) cannot be in the table yet — they come
//     from __match_any_sync over the 32 words of the step;
//   * the warp then walks its hits in position order (greedy, like the reference): the match is
//     extended cooperatively, 128 bytes per compare round, and emitted; hits covered by an emitted
//     match are dropped;
//   * after steps without a hit the stride grows (the reference's skip acceleration, lz4.c:634,957).
//
// Frames are written with B.Indep = 1 (FLG 0x60), 64 KB blocks, no checksums, no content size:
// valid for the reference reader (lz4frame.c:1151,1676) and decodable block-parallel.  A block that
// does not shrink is stored (bit 31), as LZ4F_makeBlock does (lz4frame.c:750-754).  Output bytes are
// not identical to the reference's (the format does not require it); the ratio is reported next to
// the reference's in the tests and the bench.
#pragma once
#include "common.cuh"
#include "xxh3.cuh"
#include "lz4_decode.cuh"
#include "../../include/zpack_b200.h"

#define PK_HASH_LOG 12
#define PK_TABLE (1u << PK_HASH_LOG)
#define PK_WARPS 2

ZPB_DEVINL u32 pk_load32(const u8 *p) {  // unaligned 4-byte read-only load
    const u32 *s = reinterpret_cast<const u32 *>((uintptr_t)p & ~(uintptr_t)3);
    u32 sh = ((u32)(uintptr_t)p & 3u) * 8u;
    u32 w0 = __ldg(s);
    if (sh == 0) return w0;
    return __funnelshift_r(w0, __ldg(s + 1), sh);
}

// number of equal bytes of a[0..max) and b[0..max), computed by the whole warp (uniform result)
ZPB_DEVINL u32 pk_extend(const u8 *a, const u8 *b, u32 max, int lane) {
    u32 base = 0;
    for (;;) {
        u32 k = base + 4u * lane;
        u32 eq = 4;  // equal bytes in this lane's 4-byte unit
        bool stop;
        if (k + 4 <= max) {
            u32 d = pk_load32(a + k) ^ pk_load32(b + k);
            if (d) eq = (u32)(__ffs(d) - 1) >> 3;
            stop = d != 0;
        } else {
            eq = 0;
            u32 rem = k < max ? max - k : 0u;
            while (eq < rem && a[k + eq] == b[k + eq]) ++eq;
            stop = true;
        }
        u32 bal = __ballot_sync(0xffffffffu, stop);
        if (bal) {
            int fl = __ffs(bal) - 1;
            u32 e = __shfl_sync(0xffffffffu, eq, fl);
            return base + 4u * fl + e;
        }
        base += 128;
    }
}

// one LZ4 sequence; all lanes call with uniform arguments.  Returns the new output position.
ZPB_DEVINL u32 pk_emit(u8 *dst, u32 op, const u8 *lit_src, u32 lit, u32 off, u32 ml, int lane) {
    // token + literal-length extension
    u32 mlc = ml ? ml - 4 : 0;
    if (lane == 0) dst[op] = (u8)(((lit < 15 ? lit : 15u) << 4) | (mlc < 15 ? mlc : 15u));
    ++op;
    if (lit >= 15) {
        u32 r = lit - 15, n255 = r / 255;
        for (u32 i = lane; i < n255; i += 32) dst[op + i] = 255;
        if (lane == 0) dst[op + n255] = (u8)(r - n255 * 255);
        op += n255 + 1;
    }
    Group<32> g;
    group_copy<32>(g, dst + op, lit_src, lit);
    op += lit;
    if (ml) {
        if (lane == 0) { dst[op] = (u8)off; dst[op + 1] = (u8)(off >> 8); }
        op += 2;
        if (mlc >= 15) {
            u32 r = mlc - 15, n255 = r / 255;
            for (u32 i = lane; i < n255; i += 32) dst[op + i] = 255;
            if (lane == 0) dst[op + n255] = (u8)(r - n255 * 255);
            op += n255 + 1;
        }
    }
    return op;
}
ZPB_DEVINL u32 pk_seq_bound(u32 lit, u32 ml) {
    return 1 + lit + (lit >= 15 ? (lit - 15) / 255 + 1 : 0) + (ml ? 2 + (ml - 4 >= 15 ? (ml - 19) / 255 + 1 : 0) : 0);
}

// Compress src[0..n) into dst[0..cap).  Returns the compressed size, or 0 when it does not fit
// (the caller then stores the block).  `table` is this warp's PK_TABLE-entry table; stale entries
// from earlier blocks are harmless because every candidate is verified against the input.
__device__ __noinline__ u32 pk_compress_block(const u8 *__restrict__ src, u32 n, u8 *dst, u32 cap, u32 *table,
                                              int accel, int lane) {
    u32 anchor = 0, op = 0;
    if (n >= 13) {  // lz4.c:883: shorter inputs are all literals
        const u32 mflimit = n - 12, matchlimit = n - 5;
        u32 cur = 0, miss = 0;
        while (cur <= mflimit) {
            const u32 p = cur + lane;
            const bool valid = p <= mflimit;
            const u32 w = valid ? pk_load32(src + p) : 0u;
            const u32 hv = w * 2654435761u;
            const u32 h = hv >> (32 - PK_HASH_LOG), tag = (hv >> 4) & 0xFFFFu;
            const u32 vmask = __ballot_sync(0xffffffffu, valid);
            u32 ent = 0, peers = 0;
            if (valid) {
                ent = table[h];
                peers = __match_any_sync(vmask, w) & ((1u << lane) - 1u);
            }
            __syncwarp();
            if (valid) table[h] = p | (tag << 16);
            u32 cand = 0;
            bool hit = false;
            if (peers) {  // same word at a lower lane of this step: exact, no gather needed
                cand = cur + (31 - __clz(peers));
                hit = true;
            } else if (valid && (ent >> 16) == tag) {
                cand = ent & 0xFFFFu;
                hit = cand < p && pk_load32(src + cand) == w;
            }
            u32 mm = __ballot_sync(0xffffffffu, hit);
            u32 next_cur = cur + 32;
            if (!mm) {
                ++miss;
                next_cur += 32u * ((miss * (u32)accel) >> 3);
            } else {
                miss = 0;
                while (mm) {
                    const int fl = __ffs(mm) - 1;
                    const u32 P = cur + fl, Cd = __shfl_sync(0xffffffffu, cand, fl);
                    const u32 len = 4 + pk_extend(src + P + 4, src + Cd + 4, matchlimit - (P + 4), lane);
                    const u32 lit = P - anchor;
                    // room for this sequence and for the worst-case tail (last literals are at least 5)
                    if (op + pk_seq_bound(lit, len) + 8 > cap) return 0;
                    op = pk_emit(dst, op, src + anchor, lit, P - Cd, len, lane);
                    anchor = P + len;
                    if (anchor >= cur + 32) { if (anchor > next_cur) next_cur = anchor; mm = 0; }
                    else mm &= ~((1u << (anchor - cur)) - 1u);
                }
            }
            cur = next_cur;
        }
    }
    const u32 lit = n - anchor;
    if (op + pk_seq_bound(lit, 0) > cap) return 0;
    op = pk_emit(dst, op, src + anchor, lit, 0, 0, lane);
    return op;
}

__global__ void __launch_bounds__(32 * PK_WARPS)
lz4_pack_kernel(const u8 *__restrict__ in, u64 in_size, u8 *out, u64 out_size, const zpb_file *__restrict__ files,
                const u32 *__restrict__ order, u32 n, u32 *counter, u64 *comp_size, u64 *digest, int *status) {
    __shared__ u32 tables[PK_WARPS][PK_TABLE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 *table = tables[warp];
    Group<32> g;
    for (;;) {
        u32 slot = 0;
        if (lane == 0) slot = atomicAdd(counter, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= n) break;
        const u32 idx = order ? order[slot] : slot;
        const zpb_file f = files[idx];
        int st = ST_OK;
        u64 csz = 0, dg = 0;
        if (f.src_off > in_size || f.size > in_size - f.src_off || f.dst_off > out_size ||
            f.dst_cap > out_size - f.dst_off) {
            st = ZPB_ST_COMPRESS_FAILED;
        } else if (f.method == ZPB_METHOD_NONE) {              // zpack_write.c:216-219
            if (f.dst_cap < f.size) st = ZPB_ST_COMPRESS_FAILED;
            else {
                const u8 *src = in + f.src_off;
                u8 *dst = out + f.dst_off;
                Xxh3Stream<32> hs;
                hs.init(src, f.size, g);
                for (u64 done = 0; done < f.size;) {
                    u64 chunk = f.size - done;
                    if (chunk > (1u << 20)) chunk = 1u << 20;
                    group_copy<32>(g, dst + done, src + done, (u32)chunk);
                    done += chunk;
                    hs.advance(done, g);
                }
                dg = hs.finish(g);
                csz = f.size;
            }
        } else if (f.method == ZPB_METHOD_LZ4) {
            const u64 nblocks = (f.size + 65535) >> 16;
            if (f.dst_cap < 7 + 4 * nblocks + f.size + 4) st = ZPB_ST_COMPRESS_FAILED;  // LZ4F_compressBound role
            else if (f.level >= 3) st = ST_NOT_AVAILABLE;   // lz4frame.c:799-807 selects LZ4 HC from level 3 on: not built; never a silent downgrade
            else {
                const u8 *src = in + f.src_off;
                u8 *dst = out + f.dst_off;
                // frame header: magic, FLG (v01, B.Indep), BD (64 KB), HC (lz4frame.c:669-700)
                if (lane == 0) {
                    dst[0] = 0x04; dst[1] = 0x22; dst[2] = 0x4D; dst[3] = 0x18; dst[4] = 0x60; dst[5] = 0x40;
                }
                __syncwarp();
                u32 hc = (xxh32_dev(dst + 4, 2, 0) >> 8) & 0xFF;
                if (lane == 0) dst[6] = (u8)hc;
                u64 op = 7;
                int accel = f.level < 0 ? -f.level + 1 : 1;             // lz4frame.c:768,779
                Xxh3Stream<32> hs;
                hs.init(src, f.size, g);
                for (u64 b = 0; b < nblocks; ++b) {
                    const u32 blen = (u32)(f.size - (b << 16) < 65536 ? f.size - (b << 16) : 65536);
                    const u8 *bsrc = src + (b << 16);
                    u32 c = pk_compress_block(bsrc, blen, dst + op + 4, blen - 1, table, accel, lane);
                    u32 hdr = c;
                    if (c == 0) {                                       // stored (lz4frame.c:750-754)
                        __syncwarp();                                   // the abandoned attempt's bytes sit where the copy writes
                        group_copy<32>(g, dst + op + 4, bsrc, blen);
                        c = blen;
                        hdr = blen | 0x80000000u;
                    }
                    if (lane == 0) {
                        dst[op] = (u8)hdr; dst[op + 1] = (u8)(hdr >> 8); dst[op + 2] = (u8)(hdr >> 16); dst[op + 3] = (u8)(hdr >> 24);
                    }
                    op += 4 + c;
                    hs.advance((b << 16) + blen, g);
                }
                if (lane < 4) dst[op + lane] = 0;                      // EndMark
                op += 4;
                dg = hs.finish(g);
                csz = op;
            }
        } else if (f.method == ZPB_METHOD_ZSTD) {
            // A VALID zstd frame without entropy coding: Raw_Blocks of up to 128 KB (zstd_compression_format.md:320-404;
            // what ZSTD_compress itself emits for incompressible input, zstd_compress.c ZSTD_noCompressBlock).  It keeps the
            // writer usable for ZPACK_COMPRESSION_ZSTD (the reference reader and CLI accept the archives) until the dfast +
            // FSE / Huffman encoder of SURVEY §8(f) row 3 exists; the ratio is 1.0 and is reported as such.
            const u64 nblocks = f.size ? (f.size + 131071) >> 17 : 1;
            if (f.dst_cap < 14 + 3 * nblocks + f.size) st = ZPB_ST_COMPRESS_FAILED;
            else {
                const u8 *src = in + f.src_off;
                u8 *dst = out + f.dst_off;
                if (lane == 0) {
                    // magic; FHD: 8-byte frame content size, no single-segment, no checksum, no dictionary; window 128 KB
                    dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD; dst[4] = 0xC0; dst[5] = 0x38;
                    for (int k = 0; k < 8; ++k) dst[6 + k] = (u8)(f.size >> (8 * k));
                }
                u64 op = 14;
                Xxh3Stream<32> hs;
                hs.init(src, f.size, g);
                for (u64 b = 0; b < nblocks; ++b) {
                    const u32 blen = (u32)(f.size - (b << 17) < 131072 ? f.size - (b << 17) : 131072);
                    const u32 hdr = (b + 1 == nblocks ? 1u : 0u) | (blen << 3);      // last | Raw_Block | size
                    if (lane == 0) { dst[op] = (u8)hdr; dst[op + 1] = (u8)(hdr >> 8); dst[op + 2] = (u8)(hdr >> 16); }
                    group_copy<32>(g, dst + op + 3, src + (b << 17), blen);
                    op += 3 + blen;
                    hs.advance((b << 17) + blen, g);
                }
                dg = hs.finish(g);
                csz = op;
            }
        } else {
            st = ST_METHOD_INVALID;
        }
        __syncwarp();
        if (lane == 0) {
            comp_size[idx] = csz;
            digest[idx] = dg;
            status[idx] = st;
        }
    }
}
