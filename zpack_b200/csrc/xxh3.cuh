// xxh3.cuh — XXH3-64 (seed 0, default secret) for a lane group: ZPack's entry digest.
//
// Replaces the reference's XXH3_64bits call sites (/root/reference/lib/zpack_read.c:466,
// lib/zpack_write.c:256) — algorithm: xxHash 0.8.0, externals/xxHash/xxhash.h:2763-2969
// (<= 240 bytes) and :3682-3768 (long input).
//
// Layout of the long path on a group of G lanes: the 8 accumulators are owned pairwise —
// lane l works for pair j = l & 3 (accumulators 2j and 2j+1), because a 16-byte load at
// stripe offset 16j carries exactly input words 2j and 2j+1 and the "swap adjacent lanes"
// add (xxhash.h:3515) never leaves the pair.  One cooperative step therefore covers G/4
// stripes with one LDG.128 per lane and no cross-lane traffic; the G/4 partial sums per pair
// are folded with xor-shuffles once per 1024-byte block, right before the scramble
// (xxhash.h:3527-3534).  The running accumulators are kept replicated in every lane.
//
// The digest is "fused" into the decoders by calling Xxh3Stream::advance() as the output front
// moves: each finished 1 KiB block is re-read while it is still in L2, so the verify pass
// costs no extra HBM traffic (algorithmic bytes = comp + uncomp).
#pragma once
#include "common.cuh"

__constant__ u64 c_xxh3_key[24];      // le64(secret + 8*i), i = 0..23
__constant__ u64 c_xxh3_key_last[8];  // le64(secret + 121 + 8*i): the final stripe's keys
__constant__ u8 c_xxh3_secret[192];

#define XXH_P32_1 0x9E3779B1u
#define XXH_P32_2 0x85EBCA77u
#define XXH_P32_3 0xC2B2AE3Du
#define XXH_P64_1 0x9E3779B185EBCA87ull
#define XXH_P64_2 0xC2B2AE3D27D4EB4Full
#define XXH_P64_3 0x165667B19E3779F9ull
#define XXH_P64_4 0x85EBCA77C2B2AE63ull
#define XXH_P64_5 0x27D4EB2F165667C5ull

static const u8 h_xxh3_secret[192] = {
    0xb8,0xfe,0x6c,0x39,0x23,0xa4,0x4b,0xbe,0x7c,0x01,0x81,0x2c,0xf7,0x21,0xad,0x1c,
    0xde,0xd4,0x6d,0xe9,0x83,0x90,0x97,0xdb,0x72,0x40,0xa4,0xa4,0xb7,0xb3,0x67,0x1f,
    0xcb,0x79,0xe6,0x4e,0xcc,0xc0,0xe5,0x78,0x82,0x5a,0xd0,0x7d,0xcc,0xff,0x72,0x21,
    0xb8,0x08,0x46,0x74,0xf7,0x43,0x24,0x8e,0xe0,0x35,0x90,0xe6,0x81,0x3a,0x26,0x4c,
    0x3c,0x28,0x52,0xbb,0x91,0xc3,0x00,0xcb,0x88,0xd0,0x65,0x8b,0x1b,0x53,0x2e,0xa3,
    0x71,0x64,0x48,0x97,0xa2,0x0d,0xf9,0x4e,0x38,0x19,0xef,0x46,0xa9,0xde,0xac,0xd8,
    0xa8,0xfa,0x76,0x3f,0xe3,0x9c,0x34,0x3f,0xf9,0xdc,0xbb,0xc7,0xc7,0x0b,0x4f,0x1d,
    0x8a,0x51,0xe0,0x4b,0xcd,0xb4,0x59,0x31,0xc8,0x9f,0x7e,0xc9,0xd9,0x78,0x73,0x64,
    0xea,0xc5,0xac,0x83,0x34,0xd3,0xeb,0xc3,0xc5,0x81,0xa0,0xff,0xfa,0x13,0x63,0xeb,
    0x17,0x0d,0xdd,0x51,0xb7,0xf0,0xda,0x49,0xd3,0x16,0x55,0x26,0x29,0xd4,0x68,0x9e,
    0x2b,0x16,0xbe,0x58,0x7d,0x47,0xa1,0xfc,0x8f,0xf8,0xb8,0xd1,0x7a,0xd0,0x31,0xce,
    0x45,0xcb,0x3a,0x8f,0x95,0x16,0x04,0x28,0xaf,0xd7,0xfb,0xca,0xbb,0x4b,0x40,0x7e,
};

// Host: fill the __constant__ tables once per device.
static inline cudaError_t xxh3_upload_tables() {
    u64 key[24], last[8];
    for (int i = 0; i < 24; ++i) {
        u64 v = 0;
        for (int b = 7; b >= 0; --b) v = (v << 8) | h_xxh3_secret[8 * i + b];
        key[i] = v;
    }
    for (int i = 0; i < 8; ++i) {
        u64 v = 0;
        for (int b = 7; b >= 0; --b) v = (v << 8) | h_xxh3_secret[121 + 8 * i + b];
        last[i] = v;
    }
    cudaError_t e = cudaMemcpyToSymbol(c_xxh3_key, key, sizeof key);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_xxh3_key_last, last, sizeof last);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_xxh3_secret, h_xxh3_secret, sizeof h_xxh3_secret);
}

ZPB_DEVINL u64 xxh_sec64(int off) {  // unaligned le64 from the secret
    u64 v = 0;
#pragma unroll
    for (int b = 7; b >= 0; --b) v = (v << 8) | c_xxh3_secret[off + b];
    return v;
}
ZPB_DEVINL u64 xxh_fold128(u64 a, u64 b) { return (a * b) ^ __umul64hi(a, b); }
ZPB_DEVINL u64 xxh3_avalanche(u64 h) {
    h ^= h >> 37; h *= 0x165667919E3779F9ull; h ^= h >> 32; return h;
}
ZPB_DEVINL u64 xxh64_avalanche(u64 h) {
    h ^= h >> 33; h *= XXH_P64_2; h ^= h >> 29; h *= XXH_P64_3; h ^= h >> 32; return h;
}
ZPB_DEVINL u64 xxh_rotl64(u64 v, int r) { return (v << r) | (v >> (64 - r)); }
ZPB_DEVINL u64 xxh_mix16(const u8 *p, int soff) {
    return xxh_fold128(ld64u(p) ^ xxh_sec64(soff), ld64u(p + 8) ^ xxh_sec64(soff + 8));
}

// Inputs of 0..240 bytes: closed forms, evaluated by whoever calls (uniformly by a whole group:
// every lane computes the same value; these entries are tiny so the redundancy is free).
__device__ __noinline__ u64 xxh3_small(const u8 *p, u32 n) {
    if (n <= 16) {
        if (n > 8) {
            u64 lo = ld64u(p) ^ (xxh_sec64(24) ^ xxh_sec64(32));
            u64 hi = ld64u(p + n - 8) ^ (xxh_sec64(40) ^ xxh_sec64(48));
            u64 sw = ((u64)__byte_perm((u32)lo, 0, 0x0123) << 32) | __byte_perm((u32)(lo >> 32), 0, 0x0123);
            return xxh3_avalanche((u64)n + sw + hi + xxh_fold128(lo, hi));
        }
        if (n >= 4) {
            u64 v = (u64)ld32u(p + n - 4) + ((u64)ld32u(p) << 32);
            u64 h = v ^ (xxh_sec64(8) ^ xxh_sec64(16));
            h ^= xxh_rotl64(h, 49) ^ xxh_rotl64(h, 24);
            h *= 0x9FB21C651E98DF25ull;
            h ^= (h >> 35) + n;
            h *= 0x9FB21C651E98DF25ull;
            return h ^ (h >> 28);
        }
        if (n) {
            u32 comb = ((u32)p[0] << 16) | ((u32)p[n >> 1] << 24) | p[n - 1] | (n << 8);
            u64 flip = (u64)((u32)xxh_sec64(0) ^ (u32)(xxh_sec64(0) >> 32));
            return xxh64_avalanche((u64)comb ^ flip);
        }
        return xxh64_avalanche(xxh_sec64(56) ^ xxh_sec64(64));
    }
    u64 acc = (u64)n * XXH_P64_1;
    if (n <= 128) {
        int pairs = n > 96 ? 4 : n > 64 ? 3 : n > 32 ? 2 : 1;
        for (int k = pairs - 1; k >= 0; --k) {
            acc += xxh_mix16(p + 16 * k, 32 * k);
            acc += xxh_mix16(p + n - 16 * (k + 1), 32 * k + 16);
        }
        return xxh3_avalanche(acc);
    }
    int rounds = (int)n / 16;
    for (int i = 0; i < 8; ++i) acc += xxh_mix16(p + 16 * i, 16 * i);
    acc = xxh3_avalanche(acc);
    for (int i = 8; i < rounds; ++i) acc += xxh_mix16(p + 16 * i, 16 * (i - 8) + 3);
    acc += xxh_mix16(p + n - 16, 136 - 17);
    return xxh3_avalanche(acc);
}

// Incremental long-input hasher over a buffer that is being produced front-to-back.
// `base` must be 16-byte aligned (the batch API guarantees it for every entry's dst).
template <int G>
struct Xxh3Stream {
    u64 acc0, acc1;   // running accumulators of this lane's pair (replicated across the group)
    u64 full_blocks;  // (len-1)/1024: blocks that end in a scramble
    u64 next_block;   // first block not yet folded in
    u64 len;
    const u8 *base;
    bool aligned;     // base % 16 == 0 -> LDG.128, else byte-assembled loads

    ZPB_DEVINL void init(const u8 *b, u64 n, const Group<G> &g) {
        base = b; len = n; next_block = 0;
        aligned = (((uintptr_t)b) & 15) == 0;
        full_blocks = n > 240 ? (n - 1) >> 10 : 0;
        int j = g.l & 3;
        // XXH3_INIT_ACC (xxhash.h:3749-3750)
        const u64 init[8] = {XXH_P32_3, XXH_P64_1, XXH_P64_2, XXH_P64_3,
                             XXH_P64_4, XXH_P32_2, XXH_P64_5, XXH_P32_1};
        acc0 = j == 0 ? init[0] : j == 1 ? init[2] : j == 2 ? init[4] : init[6];
        acc1 = j == 0 ? init[1] : j == 1 ? init[3] : j == 2 ? init[5] : init[7];
    }

    // one 16-byte piece = words 2j, 2j+1 of a stripe; k0/k1 are that stripe's keys for them
    static ZPB_DEVINL void piece(u64 &s0, u64 &s1, uint4 v, u64 k0, u64 k1) {
        u64 d0 = ((u64)v.y << 32) | v.x, d1 = ((u64)v.w << 32) | v.z;
        u64 x0 = d0 ^ k0, x1 = d1 ^ k1;
        s0 += d1 + (u64)(u32)x0 * (u64)(u32)(x0 >> 32);
        s1 += d0 + (u64)(u32)x1 * (u64)(u32)(x1 >> 32);
    }

    ZPB_DEVINL void fold(u64 &s0, u64 &s1, const Group<G> &g) const {
#pragma unroll
        for (int m = 4; m < G; m <<= 1) {
            s0 += g.xor_(s0, m);
            s1 += g.xor_(s1, m);
        }
    }

    // accumulate `stripes` (<= 16) stripes starting at p, which is stripe 0 of its block
    ZPB_DEVINL void stripes_sum(u64 &s0, u64 &s1, const u8 *p, u32 stripes, const Group<G> &g) const {
        int j = g.l & 3;
        constexpr int SPS = G / 4;  // stripes per step
#pragma unroll 1
        for (u32 s = g.l >> 2; s < stripes; s += SPS) {
            const u8 *q = p + 64 * s + 16 * j;
            uint4 v = aligned ? ldg128(q)
                              : make_uint4(ld32u(q), ld32u(q + 4), ld32u(q + 8), ld32u(q + 12));
            piece(s0, s1, v, c_xxh3_key[s + 2 * j], c_xxh3_key[s + 2 * j + 1]);
        }
    }

    // Fold in every full block that lies entirely below `front` (bytes produced so far).
    // Callers need not synchronise first: the group barrier here orders the other lanes' stores.
    ZPB_DEVINL void advance(u64 front, const Group<G> &g) {
        if (!(next_block < full_blocks && ((next_block + 1) << 10) <= front)) return;
        g.sync();
        while (next_block < full_blocks && ((next_block + 1) << 10) <= front) {
            u64 s0 = 0, s1 = 0;
            stripes_sum(s0, s1, base + (next_block << 10), 16, g);
            fold(s0, s1, g);
            int j = g.l & 3;
            u64 a0 = acc0 + s0, a1 = acc1 + s1;
            a0 ^= a0 >> 47; a0 ^= c_xxh3_key[16 + 2 * j]; a0 *= XXH_P32_1;
            a1 ^= a1 >> 47; a1 ^= c_xxh3_key[16 + 2 * j + 1]; a1 *= XXH_P32_1;
            acc0 = a0; acc1 = a1;
            ++next_block;
        }
    }

    // All `len` bytes are in place: finish and return the digest (same value in every lane).
    ZPB_DEVINL u64 finish(const Group<G> &g) {
        g.sync();
        if (len <= 240) return xxh3_small(base, (u32)len);
        advance(len, g);
        int j = g.l & 3;
        const u8 *tail = base + (full_blocks << 10);
        u32 tail_stripes = (u32)(((len - 1) - (full_blocks << 10)) >> 6);
        u64 s0 = 0, s1 = 0;
        stripes_sum(s0, s1, tail, tail_stripes, g);
        if (g.l < 4) {  // last stripe: the final 64 bytes, arbitrary alignment, keys at secret+121
            const u8 *q = base + len - 64 + 16 * j;
            uint4 v = make_uint4(ld32u(q), ld32u(q + 4), ld32u(q + 8), ld32u(q + 12));
            piece(s0, s1, v, c_xxh3_key_last[2 * j], c_xxh3_key_last[2 * j + 1]);
        }
        fold(s0, s1, g);
        u64 a0 = acc0 + s0, a1 = acc1 + s1;
        // mergeAccs (xxhash.h:3714-3747): pair j uses secret + 11 + 16j
        u64 m = xxh_fold128(a0 ^ xxh_sec64(11 + 16 * j), a1 ^ xxh_sec64(19 + 16 * j));
        m += g.xor_(m, 1);
        m += g.xor_(m, 2);
        return xxh3_avalanche(len * XXH_P64_1 + m);
    }
};
