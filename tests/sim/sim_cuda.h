// sim_cuda.h — TEST INFRASTRUCTURE: the CUDA names the kernels of zpack_b200/csrc use, mapped onto the fiber emulation in
// sim_rt.h.  Included (instead of <cuda_runtime.h>) only when a kernel header is compiled by g++ with -DZPB_SIM.
#pragma once
#include <cstddef>
#include <cstdint>

#include "sim_rt.h"

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __constant__

#define threadIdx (sim::tidx())
#define blockIdx (sim::bidx())
#define blockDim (sim::bdim())
#define gridDim (sim::gdim())

typedef int cudaError_t;
#define cudaSuccess 0

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ---- warp collectives
template <typename T> static inline uint64_t sim_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T) < 8 ? sizeof(T) : 8); return b; }
template <typename T> static inline T sim_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <typename T> static inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
    const uint64_t *rx = sim::collective(mask, sim_bits(v), false);
    const int lane = sim::cur_lane();
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    return ((mask >> s) & 1u) ? sim_unbits<T>(rx[s]) : v;
}
template <typename T> static inline T __shfl_xor_sync(uint32_t mask, T v, int m, int width = 32) {
    const uint64_t *rx = sim::collective(mask, sim_bits(v), false);
    const int lane = sim::cur_lane();
    const int s = lane ^ m;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) return v;
    return ((mask >> s) & 1u) ? sim_unbits<T>(rx[s]) : v;
}
template <typename T> static inline T __shfl_down_sync(uint32_t mask, T v, unsigned d, int width = 32) {
    const uint64_t *rx = sim::collective(mask, sim_bits(v), false);
    const int lane = sim::cur_lane();
    const int s = lane + (int)d;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) return v;
    return ((mask >> s) & 1u) ? sim_unbits<T>(rx[s]) : v;
}
template <typename T> static inline T __shfl_up_sync(uint32_t mask, T v, unsigned d, int width = 32) {
    const uint64_t *rx = sim::collective(mask, sim_bits(v), false);
    const int lane = sim::cur_lane();
    const int s = lane - (int)d;
    if (s < (lane & ~(width - 1))) return v;
    return ((mask >> s) & 1u) ? sim_unbits<T>(rx[s]) : v;
}
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    const uint64_t *rx = sim::collective(mask, pred ? 1 : 0, false);
    uint32_t r = 0;
    for (int k = 0; k < 32; ++k) if (((mask >> k) & 1u) && rx[k]) r |= 1u << k;
    return r;
}
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { sim::collective(mask, 0, true); }
static inline void __syncthreads() { sim::cta_barrier(); }
static inline uint32_t __activemask() { return sim::activemask(); }
static inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) {
    const uint64_t *rx = sim::collective(mask, v, false);
    uint32_t r = 0;
    for (int k = 0; k < 32; ++k) if (((mask >> k) & 1u) && (uint32_t)rx[k] > r) r = (uint32_t)rx[k];
    return r;
}
static inline uint32_t __reduce_min_sync(uint32_t mask, uint32_t v) {
    const uint64_t *rx = sim::collective(mask, v, false);
    uint32_t r = 0xffffffffu;
    for (int k = 0; k < 32; ++k) if (((mask >> k) & 1u) && (uint32_t)rx[k] < r) r = (uint32_t)rx[k];
    return r;
}
static inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v) {
    const uint64_t *rx = sim::collective(mask, v, false);
    uint32_t r = 0;
    for (int k = 0; k < 32; ++k) if ((mask >> k) & 1u) r += (uint32_t)rx[k];
    return r;
}
static inline uint32_t __reduce_or_sync(uint32_t mask, uint32_t v) {
    const uint64_t *rx = sim::collective(mask, v, false);
    uint32_t r = 0;
    for (int k = 0; k < 32; ++k) if ((mask >> k) & 1u) r |= (uint32_t)rx[k];
    return r;
}
template <typename T> static inline uint32_t __match_any_sync(uint32_t mask, T v) {
    const uint64_t *rx = sim::collective(mask, sim_bits(v), false);
    const uint64_t mine = sim_bits(v);
    uint32_t r = 0;
    for (int k = 0; k < 32; ++k) if (((mask >> k) & 1u) && rx[k] == mine) r |= 1u << k;
    return r;
}

// ---- scalar intrinsics
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline long long clock64() { return 0; }

template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

static inline size_t __cvta_generic_to_shared(const void *p) {
    return (size_t)((const uint8_t *)p - sim::S().smem.data()) + sim::SMEM_BASE;
}
static inline cudaError_t cudaMemcpyToSymbol(void *sym, const void *src, size_t n) { memcpy(sym, src, n); return cudaSuccess; }
#define cudaMemcpyToSymbol(sym, src, n) cudaMemcpyToSymbol((void *)&(sym), (src), (n))

template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
