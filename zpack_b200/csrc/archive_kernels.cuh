// archive_kernels.cuh — the container level on the device: an archive image that stays in HBM
// is opened (central directory -> entry table), assembled (payload slots -> header + data + CDR + EOCDR) and copied
// entry by entry into another archive without a host pass over its bytes.
//
//   arc_layout_kernel   exclusive prefix sums over the entries (CTA per tile of 4096, running totals handed from tile to tile): entry.offset (the offset table of
//                       zpack_write_files / ZPACK_ADD_OFFSET_AND_SIZE, /root/reference/lib/zpack_write.c:280-343), the
//                       position of every CDR record (zpack_write_cdr_ex's block size loop, zpack_write.c:720-736) and
//                       the first 64 KB copy chunk of every entry
//   arc_chunks_kernel   thread per 64 KB copy chunk: its entry (binary search over the chunk table) and byte ranges
//   arc_copy_kernel     the byte moves of zpack_write_files_from_archive (zpack_write.c:345-428: one memcpy per entry)
//                       as 64 KB chunks that the CTAs draw from a counter as they become free; 16-byte stores, every aligned source vector loaded
//                       once and realigned in registers (neighbouring lane's vector by shuffle, word select, funnel shift)
//   arc_cdr_kernel      zpack_write_cdr_memory + zpack_write_eocdr (+ the archive header) (zpack_write.c:687-711, 778-785)
//   cdr_*_kernel        zpack_read_file_entries_memory (/root/reference/lib/zpack_read.c:109-166).  The records are a
//                       linked list (each starts where the previous one's name ends), which the reference walks serially;
//                       here every byte of the directory is taken as a possible record start and walked to the end of
//                       its 1 KB tile (shared memory), the tile exits are composed over 64 KB super-tiles, one thread
//                       hops over the super-tiles, and the true starts come back down the same two levels — 64 K entries
//                       cost ~60 dependent global loads on the serial thread instead of 64 K.
//
// HBM-bound (copy) or latency-bound at microsecond scale (the rest); nothing here is a contraction.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

struct ArcEntry {          // = zpb_arc_entry (include/zpack_b200.h), 64 bytes
    u64 src_off;           // where the entry's compressed bytes are in the source buffer
    u64 comp_size, uncomp_size, hash;
    u64 name_off;          // of the name in the names blob (build) / in the CDR body (open)
    u32 name_len, method;
    u64 offset;            // entry.offset in the archive (build: out, or in when the layout is the caller's)
    u64 reserved;
};
static_assert(sizeof(ArcEntry) == 64, "zpb_arc_entry layout");

#define ARC_CHUNK_LOG 16u
#define ARC_CHUNK (1u << ARC_CHUNK_LOG)   // bytes of one copy work item
#define ARC_COPY_THREADS 256
#define ARC_SCAN_THREADS 1024
#define ARC_FIXED 35u                     // ZPACK_FILE_ENTRY_FIXED_SIZE (lib/zpack.h:44)
#define ARC_CDR_HDR 20u
#define ARC_DATA_START 10u

ZPB_DEVINL void arc_sts64(u32 a, u64 v) { sts32(a, (u32)v); sts32(a + 4, (u32)(v >> 32)); }
ZPB_DEVINL u64 arc_lds64(u32 a) { return (u64)lds32(a) | ((u64)lds32(a + 4) << 32); }
ZPB_DEVINL void arc_st16(u8 *p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); }
ZPB_DEVINL void arc_st32(u8 *p, u32 v) { arc_st16(p, v); arc_st16(p + 2, v >> 16); }
ZPB_DEVINL void arc_st64(u8 *p, u64 v) { arc_st32(p, (u32)v); arc_st32(p + 4, (u32)(v >> 32)); }

// ---- layout: three exclusive prefix sums over the table.  totals: [0] Σ comp_size, [1] Σ (35 + name_len), [2] Σ chunks,
// [4] the copy kernel's chunk counter, [5] the tile ticket (both zeroed by the host before the launch).
// One CTA per tile of 4 x blockDim entries, tiles taken by ticket so that a tile's predecessor has always started: every
// CTA scans its tile on its own (loads of all tiles overlap — one SM alone pulls the 64-byte records at ~35 GB/s), then
// thread 0 waits for the predecessor's running totals in `chain` (4 u64 per tile: three sums and a ready flag), adds its
// own and publishes them; only that hand-over is serial.
ZPB_DEVINL void arc_layout_body(ArcEntry *e, u64 n, u64 base, u32 assign, u64 *rec_off, u64 *chunk_first, u64 *totals,
                                volatile u64 *chain) {
    ZPB_DYN_SMEM(smem);
    const u32 sm = smem_window(smem);            // (32 warps + the CTA) x 3 sums x 8 bytes, + 3 x 8 bytes for the tile's base, + the ticket
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) {
        const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long *>(totals + 5), 1ull);
        sts32(sm + 36 * 24, (u32)t);
    }
    __syncthreads();
    const u64 tile = lds32(sm + 36 * 24), ntiles = (n + (u64)blockDim.x * 4 - 1) / ((u64)blockDim.x * 4);
    const u64 i = tile * blockDim.x * 4 + (u64)tid * 4;      // four consecutive entries per thread: their loads overlap
    u64 a[4], b[4], c[4];
#pragma unroll
    for (u32 j = 0; j < 4; ++j) {
        a[j] = b[j] = c[j] = 0;
        if (i + j < n) { a[j] = e[i + j].comp_size; b[j] = ARC_FIXED + (u64)e[i + j].name_len; c[j] = (a[j] + ARC_CHUNK - 1) >> ARC_CHUNK_LOG; }
    }
    const u64 ma = a[0] + a[1] + a[2] + a[3], mb = b[0] + b[1] + b[2] + b[3], mc = c[0] + c[1] + c[2] + c[3];
    u64 sa = ma, sb = mb, sc = mc;
    for (u32 d = 1; d < 32; d <<= 1) {
        const u64 ta = __shfl_up_sync(0xffffffffu, sa, d), tb = __shfl_up_sync(0xffffffffu, sb, d),
                  tc = __shfl_up_sync(0xffffffffu, sc, d);
        if (lane >= d) { sa += ta; sb += tb; sc += tc; }
    }
    if (lane == 31) { arc_sts64(sm + warp * 24, sa); arc_sts64(sm + warp * 24 + 8, sb); arc_sts64(sm + warp * 24 + 16, sc); }
    __syncthreads();
    if (warp == 0) {                             // the warps' totals: scanned by one warp, exclusive prefixes back in place
        u64 x = 0, y = 0, z = 0;
        if (lane < nw) { x = arc_lds64(sm + lane * 24); y = arc_lds64(sm + lane * 24 + 8); z = arc_lds64(sm + lane * 24 + 16); }
        u64 sx = x, sy = y, sz = z;
        for (u32 d = 1; d < 32; d <<= 1) {
            const u64 tx = __shfl_up_sync(0xffffffffu, sx, d), ty = __shfl_up_sync(0xffffffffu, sy, d),
                      tz = __shfl_up_sync(0xffffffffu, sz, d);
            if (lane >= d) { sx += tx; sy += ty; sz += tz; }
        }
        if (lane < nw) { arc_sts64(sm + lane * 24, sx - x); arc_sts64(sm + lane * 24 + 8, sy - y); arc_sts64(sm + lane * 24 + 16, sz - z); }
        if (lane == 31) {                        // the tile's totals: wait for the tiles before, publish the running totals
            u64 b0 = 0, b1 = 0, b2 = 0;
            if (tile > 0) {
                while (chain[(tile - 1) * 4 + 3] == 0) spin_pause();
                __threadfence();
                b0 = chain[(tile - 1) * 4]; b1 = chain[(tile - 1) * 4 + 1]; b2 = chain[(tile - 1) * 4 + 2];
            }
            chain[tile * 4] = b0 + sx; chain[tile * 4 + 1] = b1 + sy; chain[tile * 4 + 2] = b2 + sz;
            __threadfence();
            chain[tile * 4 + 3] = 1;
            arc_sts64(sm + 33 * 24, b0); arc_sts64(sm + 33 * 24 + 8, b1); arc_sts64(sm + 33 * 24 + 16, b2);
            if (tile + 1 == ntiles) { totals[0] = b0 + sx; totals[1] = b1 + sy; totals[2] = b2 + sz; chunk_first[n] = b2 + sz; }
        }
    }
    __syncthreads();
    u64 ra = arc_lds64(sm + 33 * 24) + arc_lds64(sm + warp * 24) + sa - ma, rb = arc_lds64(sm + 33 * 24 + 8) + arc_lds64(sm + warp * 24 + 8) + sb - mb,
        rc = arc_lds64(sm + 33 * 24 + 16) + arc_lds64(sm + warp * 24 + 16) + sc - mc;
#pragma unroll
    for (u32 j = 0; j < 4; ++j) {
        if (i + j < n) {
            if (assign) e[i + j].offset = base + ra;
            rec_off[i + j] = rb;
            chunk_first[i + j] = rc;
        }
        ra += a[j]; rb += b[j]; rc += c[j];
    }
}

// ---- copy: dst and src never overlap (different buffers, or the caller's disjoint ranges)
ZPB_DEVINL uint4 arc_realign(uint4 A, uint4 B, u32 w, u32 b) {
    u32 x0, x1, x2, x3, x4;
    switch (w) {
        case 0: x0 = A.x; x1 = A.y; x2 = A.z; x3 = A.w; x4 = B.x; break;
        case 1: x0 = A.y; x1 = A.z; x2 = A.w; x3 = B.x; x4 = B.y; break;
        case 2: x0 = A.z; x1 = A.w; x2 = B.x; x3 = B.y; x4 = B.z; break;
        default: x0 = A.w; x1 = B.x; x2 = B.y; x3 = B.z; x4 = B.w; break;
    }
    return make_uint4(__funnelshift_r(x0, x1, b), __funnelshift_r(x1, x2, b), __funnelshift_r(x2, x3, b), __funnelshift_r(x3, x4, b));
}

template <u32 ARC_U>   // 16-byte vectors in flight per thread in the realigning path
ZPB_DEVINL void arc_copy_span(u8 *dst, const u8 *src, u32 len, u32 head_mask) {
    const u32 tid = threadIdx.x, nt = blockDim.x;
    u32 head = (u32)(-(intptr_t)dst) & head_mask;    // 15: vectors on the 16-byte grid; 127: warp stores on whole lines
    if (head > len) head = len;
    if (tid < head) dst[tid] = src[tid];
    dst += head; src += head; len -= head;
    const u32 rel = (u32)(uintptr_t)src & 15u;
    u32 nvec;
    if (rel == 0) {
        nvec = len >> 4;
        for (u32 v = tid; v < nvec; v += 4 * nt) {
            uint4 r[4];
#pragma unroll
            for (u32 k = 0; k < 4; ++k) if (v + k * nt < nvec) r[k] = ldg128_stream(src + 16 * (size_t)(v + k * nt));
#pragma unroll
            for (u32 k = 0; k < 4; ++k) if (v + k * nt < nvec) stg128(dst + 16 * (size_t)(v + k * nt), r[k]);
        }
    } else {
        // Source and destination are not congruent mod 16: destination vector v is made of aligned source vectors v and
        // v + 1.  A warp takes 32 * ARC_U consecutive vectors per round, lane L the vectors base + 32 k + L: every aligned
        // source vector is loaded once, v + 1 comes from the next lane by shuffle (lane 31: from lane 0's next slot; after
        // the last slot, from one extra vector that lane 0 loads).  The last source vector read is vector nvec, which has
        // to lie inside [src, src + len): 16 bytes are kept for the byte loop.
        nvec = len >= 32 ? (len - 16) >> 4 : 0;
        const u8 *sa = src - rel;
        const u32 w = rel >> 2, b = (rel & 3u) * 8u, lane = tid & 31u;
        for (u32 base = (tid >> 5) * (32u * ARC_U); base < nvec; base += nt * ARC_U) {
            const u8 *p = sa + 16 * (size_t)(base + lane);
            u8 *q = dst + 16 * (size_t)(base + lane);
            const bool full = base + 32u * ARC_U <= nvec;      // warp-uniform: every load and store of the round is in range
            uint4 A[ARC_U], X = make_uint4(0, 0, 0, 0);
            if (full) {
#pragma unroll
                for (u32 k = 0; k < ARC_U; ++k) A[k] = ldg128_stream(p + 512u * k);
                if (lane == 0) X = ldg128_stream(p + 512u * ARC_U);
            } else {
#pragma unroll
                for (u32 k = 0; k < ARC_U; ++k) {
                    A[k] = make_uint4(0, 0, 0, 0);
                    if (base + 32u * k + lane <= nvec) A[k] = ldg128_stream(p + 512u * k);
                }
            }
#pragma unroll
            for (u32 k = 0; k < ARC_U; ++k) {
                const uint4 N = k + 1 < ARC_U ? A[k + 1 < ARC_U ? k + 1 : k] : X;
                uint4 B, W;
                B.x = __shfl_down_sync(0xffffffffu, A[k].x, 1); B.y = __shfl_down_sync(0xffffffffu, A[k].y, 1);
                B.z = __shfl_down_sync(0xffffffffu, A[k].z, 1); B.w = __shfl_down_sync(0xffffffffu, A[k].w, 1);
                W.x = __shfl_sync(0xffffffffu, N.x, 0); W.y = __shfl_sync(0xffffffffu, N.y, 0);
                W.z = __shfl_sync(0xffffffffu, N.z, 0); W.w = __shfl_sync(0xffffffffu, N.w, 0);
                if (lane == 31) B = W;
                if (full || base + 32u * k + lane < nvec) stg128(q + 512u * k, arc_realign(A[k], B, w, b));
            }
        }
    }
    for (u32 i = (nvec << 4) + tid; i < len; i += nt) dst[i] = src[i];
}

struct ArcChunk { u64 src_off, dst_off; u32 len, entry; u64 pad; };   // one copy work item, 32 bytes
static_assert(sizeof(ArcChunk) == 32, "chunk descriptor");

// thread per chunk: which entry it belongs to (binary search over the chunk table — 16 dependent loads that the copy CTAs
// would otherwise each pay in front of a 2.5 us copy) and the byte ranges it moves
ZPB_DEVINL void arc_chunks_body(const ArcEntry *e, u64 n, const u64 *chunk_first, u64 nchunks, ArcChunk *out) {
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    u64 lo = 0, hi = n;                           // chunk_first[lo] <= c < chunk_first[hi]; empty entries are skipped
    while (hi - lo > 1) {
        const u64 mid = (lo + hi) >> 1;
        if (chunk_first[mid] <= c) lo = mid; else hi = mid;
    }
    const u64 off = (c - chunk_first[lo]) << ARC_CHUNK_LOG;
    const u64 left = e[lo].comp_size - off;
    uint4 *o = reinterpret_cast<uint4 *>(out + c);
    const u64 so = e[lo].src_off + off, d = e[lo].offset + off;
    stg128(o, make_uint4((u32)so, (u32)(so >> 32), (u32)d, (u32)(d >> 32)));
    stg128(o + 1, make_uint4(left < ARC_CHUNK ? (u32)left : ARC_CHUNK, (u32)lo, 0u, 0u));
}

// `next` == NULL: chunks dealt round-robin (c, c + grid, ...), the next work item fetched while this one moves;
// otherwise the CTAs draw chunk numbers from the counter (zeroed by arc_layout_kernel) as they become free.
template <u32 ARC_U>
ZPB_DEVINL void arc_copy_body(const u8 *src, u8 *dst, const ArcChunk *chunks, u64 nchunks, unsigned long long *next, u32 head_mask) {
    if (next) {
        ZPB_DYN_SMEM(smem);
        const u32 sm = smem_window(smem);
        for (;;) {
            if (threadIdx.x == 0) {
                const unsigned long long c = atomicAdd(next, 1ull);
                sts32(sm, (u32)c); sts32(sm + 4, (u32)(c >> 32));
            }
            __syncthreads();
            const u64 c = (u64)lds32(sm) | ((u64)lds32(sm + 4) << 32);
            __syncthreads();
            if (c >= nchunks) return;
            const uint4 a = ldg128(chunks + c);
            arc_copy_span<ARC_U>(dst + ((u64)a.z | ((u64)a.w << 32)), src + ((u64)a.x | ((u64)a.y << 32)), chunks[c].len, head_mask);
        }
    }
    u64 c = blockIdx.x;
    if (c >= nchunks) return;
    uint4 a = ldg128(chunks + c);
    u32 len = chunks[c].len;
    for (;;) {
        const u64 cn = c + gridDim.x;
        const bool more = cn < nchunks;
        uint4 an = a;
        u32 lenn = 0;
        if (more) { an = ldg128(chunks + cn); lenn = chunks[cn].len; }
        arc_copy_span<ARC_U>(dst + ((u64)a.z | ((u64)a.w << 32)), src + ((u64)a.x | ((u64)a.y << 32)), len, head_mask);
        if (!more) break;
        a = an; len = lenn; c = cn;
    }
}

// ---- central directory + end record (+ the 10 header bytes)
ZPB_DEVINL void arc_cdr_body(u8 *arch, const ArcEntry *e, u64 n, const u8 *names, const u64 *rec_off, u64 cdr_off,
                             u64 block_size, u32 write_header) {
    const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x, step = (u64)gridDim.x * blockDim.x;
    if (gid == 0) {
        if (write_header) { arc_st32(arch, 0x154B505Au); arc_st16(arch + 4, 1u); arc_st32(arch + 6, 0x144B505Au); }
        u8 *p = arch + cdr_off;
        arc_st32(p, 0x134B505Au); arc_st64(p + 4, n); arc_st64(p + 12, block_size);
        p += ARC_CDR_HDR + block_size;
        arc_st32(p, 0x124B505Au); arc_st64(p + 4, cdr_off);
    }
    for (u64 i = gid; i < n; i += step) {
        u8 *p = arch + cdr_off + ARC_CDR_HDR + rec_off[i];
        const u32 len = e[i].name_len;
        const u8 *nm = names + e[i].name_off;
        arc_st16(p, len);
        for (u32 k = 0; k < len; ++k) p[2 + k] = nm[k];
        p += 2 + len;
        arc_st64(p, e[i].offset); arc_st64(p + 8, e[i].comp_size); arc_st64(p + 16, e[i].uncomp_size); arc_st64(p + 24, e[i].hash);
        p[32] = (u8)e[i].method;
    }
}

// ---- open: the directory body (block_size = B bytes after the 20-byte CDR header) -> entry table
#define CDR_T 1024u                    // tile
#define CDR_S 64u                      // tiles per super-tile
#define CDR_ST (CDR_T * CDR_S)
#define CDR_END 0xFFFFFFFFu            // "the record here does not fit the block" / "no anchor"
#define CDR_MAX_BODY 0xFFF00000ull     // positions are 32-bit

// every byte of tile t as a record start: where the walk leaves the tile (absolute) and how many records it saw
ZPB_DEVINL void cdr_tile_body(const u8 *body, u64 B, u32 *jump, u8 *cnt) {
    ZPB_DYN_SMEM(smem);
    const u32 sm = smem_window(smem);
    const u64 g0 = (u64)blockIdx.x * CDR_T;
    const u32 have = (u32)(B - g0 < CDR_T + 2 ? B - g0 : CDR_T + 2);
    for (u32 i = threadIdx.x; i < have; i += blockDim.x) sts8(sm + i, body[g0 + i]);
    __syncthreads();
    for (u32 s = threadIdx.x; s < CDR_T && g0 + s < B; s += blockDim.x) {
        u32 pos = s, c = 0, out;
        for (;;) {
            if (pos >= CDR_T) { out = (u32)(g0 + pos); break; }
            const u64 gp = g0 + pos;
            if (gp + ARC_FIXED > B) { out = CDR_END; break; }
            const u32 len = lds8(sm + pos) | (lds8(sm + pos + 1) << 8);
            if (gp + ARC_FIXED + len > B) { out = CDR_END; break; }
            pos += ARC_FIXED + len; ++c;
        }
        jump[g0 + s] = out; cnt[g0 + s] = (u8)c;
    }
}

// the same over a super-tile, for every start inside its first tile
ZPB_DEVINL void cdr_super_body(u64 B, const u32 *jump, const u8 *cnt, u32 *jump2, u32 *cnt2, u64 nsuper) {
    const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nsuper * CDR_T) return;
    const u64 st = gid / CDR_T, limit = (st + 1) * CDR_ST;
    u64 p = st * CDR_ST + gid % CDR_T;
    u32 c = 0;
    while (p != CDR_END && p < limit && p < B) { c += cnt[p]; p = jump[p]; }
    jump2[gid] = (u32)p; cnt2[gid] = c;
}

// one thread: from super-tile to super-tile; leaves the first record start of each and its index
ZPB_DEVINL void cdr_chain_body(u64 B, const u32 *jump, const u8 *cnt, const u32 *jump2, const u32 *cnt2, u32 *sup_pos,
                               u32 *sup_idx, u64 *found) {
    if (blockIdx.x || threadIdx.x) return;
    u64 p = 0, cur = ~0ull;
    u32 idx = 0;
    while (p != CDR_END && p < B) {
        const u64 st = p / CDR_ST, rel = p - st * CDR_ST;
        if (st != cur) { sup_pos[st] = (u32)p; sup_idx[st] = idx; cur = st; }
        if (rel < CDR_T) { idx += cnt2[st * CDR_T + rel]; p = jump2[st * CDR_T + rel]; }
        else { idx += cnt[p]; p = jump[p]; }
    }
    *found = idx;
}

// thread per super-tile: the first record start of each of its tiles
ZPB_DEVINL void cdr_anchor_body(u64 B, const u32 *jump, const u8 *cnt, const u32 *sup_pos, const u32 *sup_idx, u32 *tile_pos,
                                u32 *tile_idx, u64 nsuper) {
    const u64 st = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (st >= nsuper) return;
    u64 p = sup_pos[st];
    if (p == CDR_END) return;
    u32 idx = sup_idx[st];
    const u64 limit = (st + 1) * CDR_ST;
    u64 cur = ~0ull;
    while (p != CDR_END && p < limit && p < B) {
        const u64 t = p / CDR_T;
        if (t != cur) { tile_pos[t] = (u32)p; tile_idx[t] = idx; cur = t; }
        idx += cnt[p]; p = jump[p];
    }
}

// thread per tile: the records that start in it
ZPB_DEVINL void cdr_emit_body(const u8 *body, u64 B, const u32 *tile_pos, const u32 *tile_idx, ArcEntry *out, u64 count,
                              u64 ntiles) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    u64 p = tile_pos[t];
    if (p == CDR_END) return;
    u64 idx = tile_idx[t];
    const u64 limit = (t + 1) * CDR_T;
    while (p < limit && idx < count && p + ARC_FIXED <= B) {
        const u32 len = ld16u(body + p);
        if (p + ARC_FIXED + len > B) break;
        const u8 *f = body + p + 2 + len;
        const u64 off = ld64u(f), cs = ld64u(f + 8), us = ld64u(f + 16), h = ld64u(f + 24);
        uint4 *o = reinterpret_cast<uint4 *>(out + idx);
        stg128(o, make_uint4((u32)off, (u32)(off >> 32), (u32)cs, (u32)(cs >> 32)));
        stg128(o + 1, make_uint4((u32)us, (u32)(us >> 32), (u32)h, (u32)(h >> 32)));
        stg128(o + 2, make_uint4((u32)(p + 2), (u32)((p + 2) >> 32), len, (u32)f[32]));
        stg128(o + 3, make_uint4((u32)off, (u32)(off >> 32), 0u, 0u));
        p += ARC_FIXED + len; ++idx;
    }
}

#ifndef ZPB_SIM
__global__ void __launch_bounds__(ARC_SCAN_THREADS)
arc_layout_kernel(ArcEntry *e, u64 n, u64 base, u32 assign, u64 *rec_off, u64 *chunk_first, u64 *totals, u64 *chain) {
    arc_layout_body(e, n, base, assign, rec_off, chunk_first, totals, chain);
}
__global__ void __launch_bounds__(ARC_COPY_THREADS, 4)
arc_copy_kernel(const u8 *__restrict__ src, u8 *__restrict__ dst, const ArcChunk *__restrict__ chunks, u64 nchunks,
                unsigned long long *next, u32 head_mask) {
    arc_copy_body<4>(src, dst, chunks, nchunks, next, head_mask);
}

__global__ void arc_chunks_kernel(const ArcEntry *e, u64 n, const u64 *chunk_first, u64 nchunks, ArcChunk *out) {
    arc_chunks_body(e, n, chunk_first, nchunks, out);
}
__global__ void arc_cdr_kernel(u8 *arch, const ArcEntry *e, u64 n, const u8 *names, const u64 *rec_off, u64 cdr_off,
                               u64 block_size, u32 write_header) {
    arc_cdr_body(arch, e, n, names, rec_off, cdr_off, block_size, write_header);
}
__global__ void __launch_bounds__(256) cdr_tile_kernel(const u8 *body, u64 B, u32 *jump, u8 *cnt) { cdr_tile_body(body, B, jump, cnt); }
__global__ void cdr_super_kernel(u64 B, const u32 *jump, const u8 *cnt, u32 *jump2, u32 *cnt2, u64 nsuper) {
    cdr_super_body(B, jump, cnt, jump2, cnt2, nsuper);
}
__global__ void cdr_chain_kernel(u64 B, const u32 *jump, const u8 *cnt, const u32 *jump2, const u32 *cnt2, u32 *sup_pos,
                                 u32 *sup_idx, u64 *found) {
    cdr_chain_body(B, jump, cnt, jump2, cnt2, sup_pos, sup_idx, found);
}
__global__ void cdr_anchor_kernel(u64 B, const u32 *jump, const u8 *cnt, const u32 *sup_pos, const u32 *sup_idx, u32 *tile_pos,
                                  u32 *tile_idx, u64 nsuper) {
    cdr_anchor_body(B, jump, cnt, sup_pos, sup_idx, tile_pos, tile_idx, nsuper);
}
__global__ void cdr_emit_kernel(const u8 *body, u64 B, const u32 *tile_pos, const u32 *tile_idx, ArcEntry *out, u64 count,
                                u64 ntiles) {
    cdr_emit_body(body, B, tile_pos, tile_idx, out, count, ntiles);
}
#endif
