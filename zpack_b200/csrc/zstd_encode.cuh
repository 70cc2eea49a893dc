// zstd_encode.cuh — zstd Compressed_Blocks from the LZ4 block compressor's matches: one thread per 64 KB block.
//
// The role of ZSTD_compressBlock_internal's back end (/root/reference/externals/zstd/lib/compress/zstd_compress.c,
// zstd_compress_sequences.c:ZSTD_encodeSequences, fse_compress.c) for the writer's ZPACK_COMPRESSION_ZSTD arm
// (lib/zpack_write.c:179).  Match finding is NOT zstd's: the sequences are the ones lz4_pack_blocks_kernel found for the
// block (pack_blocks.cuh; greedy, 64 KB window, 4-byte minimum match), read back from its LZ4 payload.  They are written
// as a valid zstd block (zstd_compression_format.md): Raw_Literals_Block + a sequences section in Predefined_Mode — the
// three default FSE distributions, encoded backwards exactly as ZSTD_encodeSequences does (states initialised from the
// last sequence, extra bits LL / ML / OF, per earlier sequence the OF, ML, LL state transitions, final states ML, OF,
// LL, end mark).  No repeat-offset codes are emitted (offset value = offset + 3 always), no Huffman literals: the ratio
// is the LZ4 compressor's plus what FSE saves on the length codes, and is reported next to ZSTD_compress level 3 by the
// tests and the bench.  A block that does not shrink stays a Raw_Block.
#pragma once
#include "common.cuh"

#define ZE_SEQ_MAX 16384u      // sequences of one 64 KB block: each covers at least 4 input bytes
#ifdef ZPB_SIM
#define ZE_CONST static const
#else
#define ZE_CONST __device__ __constant__
#endif

ZE_CONST short ZE_LL_NORM[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                                 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
ZE_CONST short ZE_ML_NORM[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
ZE_CONST short ZE_OF_NORM[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
ZE_CONST u32 ZE_LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
                               48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
ZE_CONST u8 ZE_LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,
                              1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
ZE_CONST u32 ZE_ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20,
                               21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41,
                               43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
ZE_CONST u8 ZE_ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                              0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

// FSE compression tables of the three predefined distributions (FSE_buildCTable_wksp, fse_compress.c:68-180):
// next-state table + per symbol (deltaNbBits, deltaFindState).  Built once per CTA into shared memory.
struct ZeTables {
    u16 st_ll[64], st_of[32], st_ml[64];
    u32 dnb_ll[36], dnb_of[29], dnb_ml[53];
    int dfs_ll[36], dfs_of[29], dfs_ml[53];
};

#ifdef ZPB_SIM
ZPB_DEVINL int ze_highbit(u32 v) { int r = 0; while (v >>= 1) ++r; return r; }
#else
ZPB_DEVINL int ze_highbit(u32 v) { return 31 - __clz((int)v); }
#endif

ZPB_DEVINL void ze_build(u16 *st, u32 *dnb, int *dfs, const short *norm, int nsym, int log) {
    const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u8 sym_of[64];
    int cumul[54];
    int high = size - 1;
    cumul[0] = 0;
    for (int s = 0; s < nsym; ++s) {
        if (norm[s] == -1) { cumul[s + 1] = cumul[s] + 1; sym_of[high--] = (u8)s; }
        else cumul[s + 1] = cumul[s] + norm[s];
    }
    int pos = 0;
    for (int s = 0; s < nsym; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            sym_of[pos] = (u8)s;
            do pos = (pos + step) & mask; while (pos > high);
        }
    for (int u = 0; u < size; ++u) { const int s = sym_of[u]; st[cumul[s]++] = (u16)(size + u); }
    int total = 0;
    for (int s = 0; s < nsym; ++s) {
        const int n = norm[s];
        if (n == 0) { dnb[s] = (u32)(((log + 1) << 16) - size); dfs[s] = 0; }
        else if (n == -1 || n == 1) { dnb[s] = (u32)((log << 16) - size); dfs[s] = total - 1; ++total; }
        else {
            const int mb = log - ze_highbit((u32)(n - 1));
            dnb[s] = (u32)((mb << 16) - (n << mb));
            dfs[s] = total - n;
            total += n;
        }
    }
}
ZPB_DEVINL void ze_build_tables(ZeTables &T) {
    ze_build(T.st_ll, T.dnb_ll, T.dfs_ll, ZE_LL_NORM, 36, 6);
    ze_build(T.st_of, T.dnb_of, T.dfs_of, ZE_OF_NORM, 29, 5);
    ze_build(T.st_ml, T.dnb_ml, T.dfs_ml, ZE_ML_NORM, 53, 6);
}

// codes (zstd_compress_internal.h: ZSTD_LLcode / ZSTD_MLcode, written as searches over the base tables)
ZPB_DEVINL u32 ze_ll_code(u32 ll) {
    if (ll < 16) return ll;
    if (ll >= 64) return (u32)ze_highbit(ll) + 19u;
    u32 c = 16;
    while (c < 24 && ZE_LL_BASE[c + 1] <= ll) ++c;
    return c;
}
ZPB_DEVINL u32 ze_ml_code(u32 ml) {            // ml = match length (>= 3)
    if (ml < 35) return ml - 3;
    if (ml >= 131) return (u32)ze_highbit(ml - 3) + 36u;
    u32 c = 32;
    while (c < 42 && ZE_ML_BASE[c + 1] <= ml) ++c;
    return c;
}

struct ZeBits {                 // forward little-endian bit writer (bitstream.h: BIT_addBits / BIT_flushBits)
    u64 acc;
    u32 n;
    u8 *p, *end;
    bool ovf;
};
ZPB_DEVINL void ze_add(ZeBits &b, u32 v, u32 nb) { b.acc |= (u64)v << b.n; b.n += nb; }
// The writer only ever stores aligned 32-bit words (one predicated store per flush: lanes of a warp stay converged).
// A stream that starts at an unaligned byte starts at the word below it, the bytes in front taken as `fill` (what is
// already there, or zeros where nothing is yet).
ZPB_DEVINL void ze_open(ZeBits &b, u8 *at, u8 *end, bool keep) {
    const u32 k = (u32)(uintptr_t)at & 3u;
    b.p = at - k;
    b.end = end;
    b.ovf = false;
    b.n = 8u * k;
    b.acc = keep && k ? (u64)(*reinterpret_cast<const u32 *>(b.p) & ((1u << (8u * k)) - 1u)) : 0ull;
}
ZPB_DEVINL void ze_flush(ZeBits &b) {                 // leaves at most 31 bits pending
    if (b.n >= 32) {
        if (b.p + 4 <= b.end) { *reinterpret_cast<u32 *>(b.p) = (u32)b.acc; b.p += 4; } else b.ovf = true;
        b.acc >>= 32;
        b.n -= 32;
    }
}
ZPB_DEVINL void ze_flush_all(ZeBits &b) {             // every whole byte, then the partial one
    while (b.n >= 8) {
        if (b.p < b.end) *b.p++ = (u8)b.acc; else b.ovf = true;
        b.acc >>= 8;
        b.n -= 8;
    }
    if (b.n) { if (b.p < b.end) *b.p++ = (u8)b.acc; else b.ovf = true; b.n = 0; }
}
// sequential reader of the LZ4 payload: eight aligned bytes per load
struct ZeIn {
    const u8 *base;
    u64 w;
    u32 pos;            // next byte to hand out
};
ZPB_DEVINL void ze_in_seek(ZeIn &r, u32 pos) {
    r.pos = pos;
    const u8 *a = r.base + pos;
    r.w = *reinterpret_cast<const u64 *>((uintptr_t)a & ~(uintptr_t)7) >> (8u * ((u32)(uintptr_t)a & 7u));
}
ZPB_DEVINL u32 ze_in_byte(ZeIn &r) {
    const u32 v = (u32)r.w & 0xFFu;
    ++r.pos;
    if ((((uintptr_t)r.base + r.pos) & 7u) == 0) r.w = *reinterpret_cast<const u64 *>(r.base + r.pos);
    else r.w >>= 8;
    return v;
}
ZPB_DEVINL u32 ze_init_state(const u16 *st, const u32 *dnb, const int *dfs, u32 sym) {      // FSE_initCState2
    const u32 nbo = (dnb[sym] + (1u << 15)) >> 16;
    const u32 value = (nbo << 16) - dnb[sym];
    return st[(int)(value >> nbo) + dfs[sym]];
}
ZPB_DEVINL u32 ze_encode(ZeBits &b, const u16 *st, const u32 *dnb, const int *dfs, u32 state, u32 sym) {   // FSE_encodeSymbol
    const u32 nbo = (state + dnb[sym]) >> 16;
    ze_add(b, state & ((1u << nbo) - 1u), nbo);
    return st[(int)(state >> nbo) + dfs[sym]];
}

#define ZE_FAIL 0xFFFFFFFFu
#define ZE_WIN_SEQ 1040u       // records per window: at most 1024 matches start inside 4096 positions
// a block's bodies: window w's at 1.25 x the payload offset of its first sequence + 24 * w (a zstd body can be somewhat
// larger than the LZ4 bytes it restates: 16-bit offsets cost 2 bytes there, code + extra bits here)
#define ZE_SLOT (65536u + 16384u + 512u)
#define ZE_OFF(begin, w) ((begin) + ((begin) >> 2) + 24u * (w))
#define ZE_ZBODY 16u           // per block: the body size of every window's sub-block (0: nothing to emit, ZE_FAIL: store the block raw)

// One sub-block: the LZ4 sequences in lz[begin, end) — whole sequences, each with a match — and, for the last
// sub-block of a block, the closing literals-only sequence lz[end, tail_end) -> a zstd Compressed_Block body at `out`.
// `seq`: scratch for `seq_cap` records.  Returns the body size; 0 when the range is empty; ZE_FAIL when the body does
// not fit `cap` or the payload is not what pack_blocks.cuh writes (the caller then stores the whole block raw).
ZPB_DEVINL u32 ze_encode_range(const u8 *lz, u32 begin, u32 end, u32 tail_end, u8 *out, u32 cap, u64 *seq, u32 seq_cap,
                               const ZeTables &T) {
    if (begin == end && tail_end == end) return 0;
    if (cap < 16) return ZE_FAIL;
    // ---- forward: literals to the literals section (3-byte Raw_Literals_Block header, filled in below), sequences to
    // `seq`.  The payload is read eight bytes at a time; literal bytes are gathered in a bit writer and leave as words.
    u32 lit = 0, nseq = 0;
    ZeIn in;
    in.base = lz;
    ze_in_seek(in, begin);
    ZeBits lw;
    ze_open(lw, out + 3, out + cap, false);          // the bytes in front: the header (written below), up to 3 bytes of slack of the window before
    bool bad = false;                                 // single-exit loops: the lanes of a warp (one window each) reconverge every iteration
    while (in.pos < tail_end && !bad) {
        const u32 token = ze_in_byte(in);
        u32 ll = token >> 4;
        if (ll == 15) {
            u32 x;
            do { x = in.pos < tail_end ? ze_in_byte(in) : 0u; ll += x; } while (x == 255);
        }
        if (in.pos + ll > tail_end || 3 + lit + ll + 8 >= cap) { bad = true; ll = 0; }
        for (u32 i = 0; i < ll; ++i) { ze_add(lw, ze_in_byte(in), 8); ze_flush(lw); }
        lit += ll;
        if (in.pos > end) {                           // the closing sequence of the block: literals only
            bad |= in.pos != tail_end;
        } else if (in.pos + 2 > end) {
            bad = true;
        } else {
            u32 off = ze_in_byte(in);
            off |= ze_in_byte(in) << 8;
            u32 ml = token & 15u;
            if (ml == 15) {
                u32 x;
                do { x = in.pos < end ? ze_in_byte(in) : 0u; ml += x; } while (x == 255);
            }
            ml += 4;
            if (off == 0 || nseq >= seq_cap || ll > 65535u || ml > 65535u) bad = true;
            else seq[nseq++] = (u64)ll | ((u64)ml << 16) | ((u64)off << 32);
        }
    }
    if (bad) return ZE_FAIL;
    ze_flush_all(lw);
    if (lw.ovf) return ZE_FAIL;
    out[0] = (u8)(0x0Cu | ((lit & 0xFu) << 4));       // Raw_Literals_Block, size format 11: 20-bit size
    out[1] = (u8)(lit >> 4);
    out[2] = (u8)(lit >> 12);
    u32 op = 3 + lit;
    // ---- sequences section header: count, then symbol compression modes = 0 (all predefined)
    if (nseq == 0) { out[op++] = 0; return op <= cap ? op : ZE_FAIL; }
    if (nseq < 128) out[op++] = (u8)nseq;
    else if (nseq < 0x7F00) { out[op++] = (u8)((nseq >> 8) + 0x80); out[op++] = (u8)nseq; }
    else { out[op++] = 0xFF; out[op++] = (u8)(nseq - 0x7F00); out[op++] = (u8)((nseq - 0x7F00) >> 8); }
    out[op++] = 0;
    // ---- the bitstream, from the last sequence to the first (zstd_compress_sequences.c:ZSTD_encodeSequences_body)
    ZeBits b;
    ze_open(b, out + op, out + cap, true);
    u32 st_ll, st_of, st_ml;
    {
        const u64 r = seq[nseq - 1];
        const u32 ll = (u32)(r & 0xFFFF), ml = (u32)((r >> 16) & 0xFFFF), ob = (u32)(r >> 32) + 3u;
        const u32 cl = ze_ll_code(ll), cm = ze_ml_code(ml), co = (u32)ze_highbit(ob);
        st_ml = ze_init_state(T.st_ml, T.dnb_ml, T.dfs_ml, cm);
        st_of = ze_init_state(T.st_of, T.dnb_of, T.dfs_of, co);
        st_ll = ze_init_state(T.st_ll, T.dnb_ll, T.dfs_ll, cl);
        ze_add(b, ll - ZE_LL_BASE[cl], ZE_LL_BITS[cl]);
        ze_add(b, ml - ZE_ML_BASE[cm], ZE_ML_BITS[cm]);
        ze_flush(b);
        ze_add(b, ob - (1u << co), co);
        ze_flush(b);
    }
    u64 r_next = nseq > 1 ? seq[nseq - 2] : 0ull;
    for (u32 k = nseq - 1; k-- > 0;) {
        const u64 r = r_next;
        if (k) r_next = seq[k - 1];                  // the record after this one is on its way while this one is encoded
        const u32 ll = (u32)(r & 0xFFFF), ml = (u32)((r >> 16) & 0xFFFF), ob = (u32)(r >> 32) + 3u;
        const u32 cl = ze_ll_code(ll), cm = ze_ml_code(ml), co = (u32)ze_highbit(ob);
        st_of = ze_encode(b, T.st_of, T.dnb_of, T.dfs_of, st_of, co);
        st_ml = ze_encode(b, T.st_ml, T.dnb_ml, T.dfs_ml, st_ml, cm);
        st_ll = ze_encode(b, T.st_ll, T.dnb_ll, T.dfs_ll, st_ll, cl);
        ze_flush(b);
        ze_add(b, ll - ZE_LL_BASE[cl], ZE_LL_BITS[cl]);
        ze_add(b, ml - ZE_ML_BASE[cm], ZE_ML_BITS[cm]);
        ze_flush(b);
        ze_add(b, ob - (1u << co), co);
        ze_flush(b);
    }
    ze_add(b, st_ml & 63u, 6);                       // FSE_flushCState: ML, OF, LL
    ze_add(b, st_of & 31u, 5);
    ze_add(b, st_ll & 63u, 6);
    ze_add(b, 1u, 1);                                // BIT_closeCStream: the end mark
    ze_flush_all(b);
    if (b.ovf) return ZE_FAIL;
    return (u32)(b.p - out);
}

#ifndef ZPB_SIM
// One THREAD per 4 KB window of the block compressor: the window's sequences become one zstd sub-block (the windows'
// payload offsets come from lz4_pack_blocks_kernel: winop).  Each walk is serial and latency-bound, so the kernel relies
// on the number of windows in flight (16 per block, every block of the round at once), not on staging.
// zbody[b * 16 + w] = body size of window w (0: none, ZE_FAIL: the block is stored raw).  Only the blocks the host marked
// (PackBlock::pad != 0: they belong to a ZPACK_COMPRESSION_ZSTD file) are encoded.
__global__ void __launch_bounds__(128)
zstd_encode_blocks_kernel(const u8 *__restrict__ lz_slots, const u32 *__restrict__ csize, const PackBlock *__restrict__ blocks,
                          const u32 *__restrict__ winop, u32 nblocks, u8 *zslot, u64 *zseq, u32 *zbody) {
    __shared__ ZeTables T;
    if (threadIdx.x == 0) ze_build_tables(T);
    __syncthreads();
    const u64 nwork = (u64)nblocks * 16u;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nwork; t += (u64)gridDim.x * blockDim.x) {
        const u32 b = (u32)(t >> 4), w = (u32)(t & 15u);
        const u32 cs = csize[b], len = blocks[b].len;
        u32 z = 0;
        if (blocks[b].pad && cs) {
            const u32 nwin = (len - 12u) / 4096u + 1u;           // as in pack_blocks.cuh (cs != 0 implies len >= 13)
            if (w < nwin) {
                const u32 begin = winop[(u64)b * 17u + w], end = winop[(u64)b * 17u + w + 1u];
                const u32 tail_end = w + 1u == nwin ? cs : end;
                z = begin <= end && tail_end <= cs && end <= tail_end
                        ? ze_encode_range(lz_slots + ((u64)b << 16), begin, end, tail_end, zslot + (u64)b * ZE_SLOT + ZE_OFF(begin, w),
                                          (tail_end - begin) + ((tail_end - begin) >> 2) + 20u, zseq + t * ZE_WIN_SEQ, ZE_WIN_SEQ, T)
                        : ZE_FAIL;
            }
        } else if (w == 0) z = ZE_FAIL;
        zbody[t] = z;
    }
}
#endif
