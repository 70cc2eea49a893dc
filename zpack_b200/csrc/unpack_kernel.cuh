// unpack_kernel.cuh — the batched "read + verify" kernel: one lane group per entry.
//
// Device-side restatement of zpack_read_file (/root/reference/lib/zpack_read.c:326-471):
// guards -> dispatch on comp_method -> decode -> XXH3-64 over uncomp_size bytes -> compare.
// Persistent CTAs pull entry indices from a global counter in the order the host chose
// (most expensive first) so that a long text entry never starts in the last wave.
#pragma once
#include "common.cuh"
#include "xxh3.cuh"
#include "lz4_decode.cuh"
#include "../../include/zpack_b200.h"

template <int G>
ZPB_DEVINL void unpack_one(const Group<G> &g, const u8 *archive, u64 archive_size, u8 *out,
                           const zpb_entry &e, int *status, u64 *digest, u64 *produced_out,
                           u32 idx, u32 *zstd_list, u32 *zstd_count) {
    int st = ST_OK;
    u64 dg = 0, produced = 0;
    if (e.comp_size == 0) {
        st = ST_OK;  // zpack_read.c:328 — nothing decoded, nothing hashed
    } else if (e.dst_cap < e.uncomp_size) {
        st = ST_TOO_SMALL;  // :329
    } else if (e.src_off > archive_size || e.comp_size > archive_size - e.src_off) {
        st = ST_OFFSET_INVALID;  // memory-safety form of :331 (the strict file_size test is the host's)
    } else {
        const u8 *src = archive + e.src_off;
        u8 *dst = out + e.dst_off;
        Xxh3Stream<G> hs;
        hs.init(dst, e.uncomp_size, g);
        switch (e.method) {
        case ZPB_METHOD_NONE:  // :352-368
            if (e.uncomp_size > e.comp_size) { st = ST_SIZE_INVALID; break; }
            for (u64 done = 0; done < e.uncomp_size;) {
                u64 chunk = e.uncomp_size - done;
                if (chunk > (1u << 20)) chunk = 1u << 20;
                group_copy<G>(g, dst + done, src + done, (u32)chunk);
                done += chunk;
                hs.advance(done, g);
            }
            produced = e.uncomp_size;
            break;
        case ZPB_METHOD_LZ4:  // :396-453
            st = lz4f_decode_entry<G>(g, src, e.comp_size, dst, e.dst_cap, hs, &produced);
            break;
        case ZPB_METHOD_ZSTD:  // :370-390 — guards passed: queue it for zstd_unpack_kernel, which reports
            if (g.l == 0) zstd_list[atomicAdd(zstd_count, 1u)] = idx;
            return;
        default:
            st = ST_METHOD_INVALID;  // :459-461
        }
        if (st == ST_OK) {
            dg = hs.finish(g);  // :466
            if (!(e.flags & ZPB_F_NO_VERIFY) && dg != e.hash) st = ST_HASH_MISMATCH;
        }
    }
    if (g.l == 0) {
        status[idx] = st;
        digest[idx] = dg;
        if (produced_out) produced_out[idx] = produced;
    }
}

template <int G>
__global__ void __launch_bounds__(256)
unpack_kernel(const u8 *__restrict__ archive, u64 archive_size, u8 *__restrict__ out,
              const zpb_entry *__restrict__ entries, const u32 *order, u32 n, const u32 *n_ptr,
              u32 *counter, int *status, u64 *digest, u64 *produced, u32 *zstd_list, u32 *zstd_count) {
    Group<G> g;
    if (n_ptr) n = *n_ptr;  // device-side count: the entries the fast path handed over
    for (;;) {
        u32 slot = 0;
        if (g.l == 0) slot = atomicAdd(counter, 1u);
        slot = g.bcast(slot, 0);
        if (slot >= n) break;
        u32 idx = order ? order[slot] : slot;
        zpb_entry e = entries[idx];
        unpack_one<G>(g, archive, archive_size, out, e, status, digest, produced, idx, zstd_list, zstd_count);
    }
}

// Standalone digest kernel (zpb_xxh3_*): one group per range.
template <int G>
__global__ void __launch_bounds__(256)
xxh3_kernel(const u8 *__restrict__ data, const u64 *__restrict__ offsets,
            const u64 *__restrict__ lengths, u32 n, u32 *counter, u64 *digest) {
    Group<G> g;
    for (;;) {
        u32 slot = 0;
        if (g.l == 0) slot = atomicAdd(counter, 1u);
        slot = g.bcast(slot, 0);
        if (slot >= n) break;
        Xxh3Stream<G> hs;
        hs.init(data + offsets[slot], lengths[slot], g);
        u64 dg = hs.finish(g);
        if (g.l == 0) digest[slot] = dg;
    }
}
