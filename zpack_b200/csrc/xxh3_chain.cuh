// xxh3_chain.cuh — the serial half of XXH3-64 for ONE entry whose blocks were decoded in parallel.
//
// XXH3's long-input loop (externals/xxHash/xxhash.h:3682-3711) is, per 1 KiB block n,
//     acc = scramble(acc + S_n)          S_n = the 16 stripe sums of block n (xxhash.h:3502-3518)
// where S_n depends on the data only (pure adds) and scramble is the non-linear per-lane step
// (xxhash.h:3527-3534).  The block-sharded decode (lz4_fast_exec_kernel in partial mode) leaves
// S_n for every complete KiB in HBM (64 bytes each); this kernel runs the chain.  Nothing about
// it is parallel beyond the 8 independent accumulator lanes: one warp, lanes 0-3
// each carrying accumulator pair (2j, 2j+1), the other lanes only help to stage S_n.
//
// Multi-GPU (BASELINE config C5): shard k's chain starts from the 64-byte accumulator state shard
// k-1 ended with (acc_in / acc_out); the last shard also folds the tail (xxhash.h:3701-3711) from
// the decoded bytes and merges (xxhash.h:3714-3747).
#pragma once
#include "common.cuh"
#include "xxh3.cuh"

#define XC_BATCH 64u   // KiB steps staged per round: 64 x 64 B = 4 KiB of shared memory, double buffered

// acc[8] in xxHash order.  nscr = how many leading S_n are followed by a scramble (all of the shard's
// complete KiBs, except that the entry's very last KiB is never scrambled: it belongs to the tail).
// When `final`: tail_ptr = decoded bytes starting at entry position tail_pos (<= both the tail start
// and total - 64), total = the entry's uncomp_size.
__global__ void __launch_bounds__(32)
xxh3_chain_kernel(const u64 *__restrict__ partials, u64 nscr, const u64 *__restrict__ acc_in, u64 *acc_out,
                  int final, const u8 *tail_ptr, u64 tail_pos, u64 total, u64 *digest_out) {
    __shared__ ulonglong2 stage[2][XC_BATCH * 4];
    const int lane = threadIdx.x;
    const int j = lane & 3;
    u64 a0, a1;
    if (acc_in) { a0 = acc_in[2 * j]; a1 = acc_in[2 * j + 1]; }
    else {
        a0 = j == 0 ? (u64)XXH_P32_3 : j == 1 ? XXH_P64_2 : j == 2 ? XXH_P64_4 : XXH_P64_5;
        a1 = j == 0 ? XXH_P64_1 : j == 1 ? XXH_P64_3 : j == 2 ? (u64)XXH_P32_2 : (u64)XXH_P32_1;
    }
    const u64 k0 = c_xxh3_key[16 + 2 * j], k1 = c_xxh3_key[16 + 2 * j + 1];
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(partials);   // 4 x 16 B per KiB
    const u64 nbatch = (nscr + XC_BATCH - 1) / XC_BATCH;
    // software pipeline: batch b+1 travels HBM -> registers while lanes 0-3 walk batch b in shared memory
    ulonglong2 r[8];
    auto fetch = [&](u64 b) {
        const u64 lo = b * XC_BATCH * 4, hi = nscr * 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            u64 i = lo + (u64)k * 32 + lane;
            r[k] = i < hi ? src[i] : make_ulonglong2(0, 0);
        }
    };
    if (nbatch) fetch(0);
    for (u64 b = 0; b < nbatch; ++b) {
        ulonglong2 *st = stage[b & 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) st[k * 32 + lane] = r[k];
        __syncwarp();
        if (b + 1 < nbatch) fetch(b + 1);
        if (lane < 4) {
            const u32 steps = (u32)(nscr - b * XC_BATCH < XC_BATCH ? nscr - b * XC_BATCH : XC_BATCH);
            u32 s = 0;
            for (; s + 8 <= steps; s += 8) {
                ulonglong2 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = st[(s + k) * 4 + j];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a0 += v[k].x; a1 += v[k].y;
                    a0 ^= a0 >> 47; a0 ^= k0; a0 *= XXH_P32_1;
                    a1 ^= a1 >> 47; a1 ^= k1; a1 *= XXH_P32_1;
                }
            }
            for (; s < steps; ++s) {
                ulonglong2 v = st[s * 4 + j];
                a0 += v.x; a1 += v.y;
                a0 ^= a0 >> 47; a0 ^= k0; a0 *= XXH_P32_1;
                a1 ^= a1 >> 47; a1 ^= k1; a1 *= XXH_P32_1;
            }
        }
        __syncwarp();
    }
    if (acc_out && lane < 4) { acc_out[2 * j] = a0; acc_out[2 * j + 1] = a1; }
    if (!final) return;

    // ---- tail (xxhash.h:3701-3711) + merge (:3714-3747); every lane needs its pair's accumulators
    a0 = __shfl_sync(0xffffffffu, a0, j);
    a1 = __shfl_sync(0xffffffffu, a1, j);
    u64 dg;
    if (total <= 240) {
        dg = xxh3_small(tail_ptr, (u32)total);   // tail_pos == 0
    } else {
        const u64 full_blocks = (total - 1) >> 10;
        const u64 tail_start = full_blocks << 10;
        const u32 tail_stripes = (u32)(((total - 1) - tail_start) >> 6);
        u64 s0 = 0, s1 = 0;
        for (u32 s = lane >> 2; s < tail_stripes; s += 8) {
            const u8 *q = tail_ptr + (tail_start - tail_pos) + 64 * s + 16 * j;
            uint4 v = make_uint4(ld32u(q), ld32u(q + 4), ld32u(q + 8), ld32u(q + 12));
            Xxh3Stream<32>::piece(s0, s1, v, c_xxh3_key[s + 2 * j], c_xxh3_key[s + 2 * j + 1]);
        }
        if (lane < 4) {
            const u8 *q = tail_ptr + (total - 64 - tail_pos) + 16 * j;
            uint4 v = make_uint4(ld32u(q), ld32u(q + 4), ld32u(q + 8), ld32u(q + 12));
            Xxh3Stream<32>::piece(s0, s1, v, c_xxh3_key_last[2 * j], c_xxh3_key_last[2 * j + 1]);
        }
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, m);
            s1 += __shfl_xor_sync(0xffffffffu, s1, m);
        }
        u64 b0 = a0 + s0, b1 = a1 + s1;
        u64 m = xxh_fold128(b0 ^ xxh_sec64(11 + 16 * j), b1 ^ xxh_sec64(19 + 16 * j));
        m += __shfl_xor_sync(0xffffffffu, m, 1);
        m += __shfl_xor_sync(0xffffffffu, m, 2);
        dg = xxh3_avalanche(total * XXH_P64_1 + m);
    }
    if (lane == 0 && digest_out) *digest_out = dg;
}
