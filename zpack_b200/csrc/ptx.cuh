// ptx.cuh — every inline-PTX access the kernels make, in one place: shared-window loads / stores by 32-bit address,
// generic loads, cp.async (LDGSTS), 1-D bulk async copies (cp.async.bulk -> UBLKCP, the TMA unit) and the mbarrier
// operations that track them (SYNCS).  With -DZPB_SIM (g++, tests/sim) the same names run on the CPU emulation, so the
// kernels' control logic is testable without a GPU; the product is always built by nvcc without ZPB_SIM.
#pragma once
#include "common.cuh"

#ifndef ZPB_SIM
// ---- shared memory by window address
ZPB_DEVINL u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
ZPB_DEVINL void sts8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
ZPB_DEVINL u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
ZPB_DEVINL u32 lds32_loose(u32 a) { return lds32(a); }   // the caller discards the bytes it does not own (see sim_rt.h)
ZPB_DEVINL void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
ZPB_DEVINL uint4 lds128(u32 a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
    return r;
}
ZPB_DEVINL void sts128(u32 a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// predicated forms (one instruction each, no branch): the load keeps `v` when !p
ZPB_DEVINL void lds8_if(bool p, u32 a, u32 &v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q ld.shared.u8 %0, [%2];\n\t}" : "+r"(v) : "r"((u32)p), "r"(a) : "memory");
}
ZPB_DEVINL void sts8_if(bool p, u32 a, u32 v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.shared.u8 [%1], %2;\n\t}" ::"r"((u32)p), "r"(a), "r"(v) : "memory");
}
ZPB_DEVINL void sts32_if(bool p, u32 a, u32 v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.shared.u32 [%1], %2;\n\t}" ::"r"((u32)p), "r"(a), "r"(v) : "memory");
}
// ---- generic addresses: one load instruction for a source that is either in shared memory or in HBM
ZPB_DEVINL u64 shared_to_generic(u32 a) {
    u64 g;
    asm volatile("{\n\t.reg .u64 t;\n\tcvt.u64.u32 t, %1;\n\tcvta.shared.u64 %0, t;\n\t}" : "=l"(g) : "r"(a));
    return g;
}
ZPB_DEVINL u64 global_to_generic(const void *p) { return (u64)(uintptr_t)p; }
ZPB_DEVINL void ldgen8_if(bool p, u64 g, u32 &v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q ld.u8 %0, [%2];\n\t}" : "+r"(v) : "r"((u32)p), "l"(g) : "memory");
}
// five consecutive aligned words; words the caller does not own may hold bytes other lanes are writing: discarded
ZPB_DEVINL void ldgen32x5_if(bool p, u64 g, u32 &a0, u32 &a1, u32 &a2, u32 &a3, u32 &a4) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
                 "@q ld.u32 %0, [%6];\n\t@q ld.u32 %1, [%6+4];\n\t@q ld.u32 %2, [%6+8];\n\t@q ld.u32 %3, [%6+12];\n\t@q ld.u32 %4, [%6+16];\n\t}"
                 : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4) : "r"((u32)p), "l"(g) : "memory");
}
ZPB_DEVINL void lds32x5_if(bool p, u32 a, u32 &a0, u32 &a1, u32 &a2, u32 &a3, u32 &a4) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
                 "@q ld.shared.u32 %0, [%6];\n\t@q ld.shared.u32 %1, [%6+4];\n\t@q ld.shared.u32 %2, [%6+8];\n\t@q ld.shared.u32 %3, [%6+12];\n\t@q ld.shared.u32 %4, [%6+16];\n\t}"
                 : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4) : "r"((u32)p), "r"(a) : "memory");
}
// ---- global memory written by the same kernel: plain (coherent) loads, never .nc
ZPB_DEVINL uint4 ldg128_coherent(const void *p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
ZPB_DEVINL u32 ldg8_coherent(const u8 *p) { u32 r; asm volatile("ld.global.u8 %0, [%1];" : "=r"(r) : "l"(p) : "memory"); return r; }
ZPB_DEVINL u32 ldg32_coherent(const void *p) { u32 r; asm volatile("ld.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory"); return r; }
ZPB_DEVINL void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
ZPB_DEVINL void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// ---- cp.async (Ampere-style, 16 bytes per lane)
ZPB_DEVINL void cp_async16(u32 smem_addr, const void *gptr, u32 src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes) : "memory");
}
ZPB_DEVINL void cp_async64_if(bool p, u32 s, const void *g) {   // one whole 64-byte chunk
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t"
                 "@q cp.async.ca.shared.global [%1], [%2], 16;\n\t@q cp.async.ca.shared.global [%1+16], [%2+16], 16;\n\t"
                 "@q cp.async.ca.shared.global [%1+32], [%2+32], 16;\n\t@q cp.async.ca.shared.global [%1+48], [%2+48], 16;\n\t}"
                 ::"r"((u32)p), "r"(s), "l"(g) : "memory");
}
ZPB_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ZPB_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
ZPB_DEVINL void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
ZPB_DEVINL void cp_async_wait_2() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }
// ---- mbarrier + 1-D bulk async copy (TMA unit)
ZPB_DEVINL void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
ZPB_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ZPB_DEVINL void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
ZPB_DEVINL void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
ZPB_DEVINL void mbar_wait(u32 bar, u32 parity) {
    asm volatile("{\n\t.reg .pred p;\n\tZPB_MBW_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra ZPB_MBD_%=;\n\tbra ZPB_MBW_%=;\n\tZPB_MBD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
#define ZPB_DYN_SMEM(name) extern __shared__ uint4 name[]
ZPB_DEVINL void spin_pause() { __nanosleep(32); }     // inside a wait on a flag another warp of the CTA sets
ZPB_DEVINL u32 smem_window(const void *p) { return (u32)__cvta_generic_to_shared(p); }
// where the dynamic shared window of a kernel without static shared memory starts on sm_100 (1 KiB is reserved in
// front of it); kernels that build addresses from it check it at entry and trap on a mismatch
#define ZPB_DYN_SMEM_BASE 0x400u

#else  // ------------------------------------------------------------------------------------------ ZPB_SIM
ZPB_DEVINL u32 lds8(u32 a) { return sim::ld_shared(a, 1); }
ZPB_DEVINL void sts8(u32 a, u32 v) { sim::st_shared(a, 1, v & 0xFFu); }
ZPB_DEVINL u32 lds32(u32 a) { return sim::ld_shared(a, 4); }
ZPB_DEVINL u32 lds32_loose(u32 a) { return sim::ld_shared(a, 4, true); }
ZPB_DEVINL void sts32(u32 a, u32 v) { sim::st_shared(a, 4, v); }
ZPB_DEVINL uint4 lds128(u32 a) {
    if (a & 15u) { fprintf(stderr, "sim: misaligned lds128 0x%x\n", a); abort(); }
    return make_uint4(sim::ld_shared(a, 4), sim::ld_shared(a + 4, 4), sim::ld_shared(a + 8, 4), sim::ld_shared(a + 12, 4));
}
ZPB_DEVINL void sts128(u32 a, uint4 v) {
    if (a & 15u) { fprintf(stderr, "sim: misaligned sts128 0x%x\n", a); abort(); }
    sim::st_shared(a, 4, v.x); sim::st_shared(a + 4, 4, v.y); sim::st_shared(a + 8, 4, v.z); sim::st_shared(a + 12, 4, v.w);
}
ZPB_DEVINL void lds8_if(bool p, u32 a, u32 &v) { if (p) v = lds8(a); }
ZPB_DEVINL void sts8_if(bool p, u32 a, u32 v) { if (p) sts8(a, v); }
ZPB_DEVINL void sts32_if(bool p, u32 a, u32 v) { if (p) sts32(a, v); }
// generic addresses: bit 63 set = shared window address in the low word, else a host pointer
ZPB_DEVINL u64 shared_to_generic(u32 a) { return (1ull << 63) | a; }
ZPB_DEVINL u64 global_to_generic(const void *p) { return (u64)(uintptr_t)p; }
ZPB_DEVINL void ldgen8_if(bool p, u64 g, u32 &v) {
    if (!p) return;
    if (g >> 63) v = lds8((u32)g); else v = *(const u8 *)(uintptr_t)g;
}
ZPB_DEVINL void ldgen32x5_if(bool p, u64 g, u32 &a0, u32 &a1, u32 &a2, u32 &a3, u32 &a4) {
    if (!p) return;
    u32 *o[5] = {&a0, &a1, &a2, &a3, &a4};
    for (int k = 0; k < 5; ++k) {
        if (g >> 63) *o[k] = lds32_loose((u32)g + 4 * k);
        else {
            if (g & 3u) { fprintf(stderr, "sim: misaligned generic word load\n"); abort(); }
            memcpy(o[k], (const u8 *)(uintptr_t)g + 4 * k, 4);
        }
    }
}
ZPB_DEVINL void lds32x5_if(bool p, u32 a, u32 &a0, u32 &a1, u32 &a2, u32 &a3, u32 &a4) {
    if (!p) return;
    a0 = lds32_loose(a); a1 = lds32_loose(a + 4); a2 = lds32_loose(a + 8); a3 = lds32_loose(a + 12); a4 = lds32_loose(a + 16);
}
ZPB_DEVINL uint4 ldg128_coherent(const void *p) {
    if ((uintptr_t)p & 15u) { fprintf(stderr, "sim: misaligned ldg128\n"); abort(); }
    uint4 r; memcpy(&r, p, 16); return r;
}
ZPB_DEVINL u32 ldg8_coherent(const u8 *p) { return *p; }
ZPB_DEVINL u32 ldg32_coherent(const void *p) { u32 r; memcpy(&r, p, 4); return r; }
ZPB_DEVINL void prefetch_l2(const void *) {}
ZPB_DEVINL void prefetch_l1(const void *) {}
ZPB_DEVINL void cp_async16(u32 s, const void *g, u32 src_bytes) { sim::cp_async(s, g, 16, src_bytes); }
ZPB_DEVINL void cp_async64_if(bool p, u32 s, const void *g) { if (p) sim::cp_async(s, g, 64, 64); }
ZPB_DEVINL void cp_async_commit() { sim::cp_async_commit(); }
ZPB_DEVINL void cp_async_wait_all() { sim::cp_async_wait(0); }
ZPB_DEVINL void cp_async_wait_1() { sim::cp_async_wait(1); }
ZPB_DEVINL void cp_async_wait_2() { sim::cp_async_wait(2); }
ZPB_DEVINL void mbar_init(u32 bar, u32 count) { sim::mbar_init(bar, count); }
ZPB_DEVINL void mbar_fence_init() {}
ZPB_DEVINL void mbar_arrive_expect_tx(u32 bar, u32 bytes) { sim::mbar_arrive_expect_tx(bar, bytes); }
ZPB_DEVINL void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) { sim::bulk_g2s(dst, src, bytes, bar); }
ZPB_DEVINL void mbar_wait(u32 bar, u32 parity) { while (!sim::mbar_test_wait(bar, parity)) sim::park(); }
#define ZPB_DYN_SMEM(name) uint4 *name = reinterpret_cast<uint4 *>(sim::S().smem.data())
ZPB_DEVINL void spin_pause() { sim::park(); }
ZPB_DEVINL u32 smem_window(const void *p) { return (u32)__cvta_generic_to_shared(p); }
#define ZPB_DYN_SMEM_BASE (sim::SMEM_BASE)
static inline void __trap() { fprintf(stderr, "sim: __trap()\n"); abort(); }
#endif

// ------------------------------------------------------------------------------------------ lane-copy blocks
// The three instruction blocks of lane_copy (lz4_fast.cuh), written out so that each predicated access is exactly one
// compare and one memory instruction: a shared-memory to shared-memory copy of up to 3 head bytes + whole words + up
// to 3 tail bytes per lane.  The word block reads five aligned words around the source (bytes of neighbouring
// sequences included) and keeps only the caller's own bytes.
#ifndef ZPB_SIM
ZPB_DEVINL u32 lane_id() { u32 l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
template <typename T> ZPB_DEVINL T keep_in_register(T v) { asm volatile("" : "+r"(v)); return v; }   // never rematerialised
// h bytes s -> d and t bytes ts -> td (h, t <= 3)
ZPB_DEVINL void copy_edges_ss(u32 h, u32 t, u32 s, u32 d, u32 ts, u32 td) {
    asm volatile(
        "{\n\t"
        ".reg .pred p0, p1, p2, q0, q1, q2;\n\t"
        ".reg .b32 a0, a1, a2, b0, b1, b2;\n\t"
        "setp.gt.u32 p0, %0, 0;\n\t"
        "setp.gt.u32 p1, %0, 1;\n\t"
        "setp.gt.u32 p2, %0, 2;\n\t"
        "setp.gt.u32 q0, %1, 0;\n\t"
        "setp.gt.u32 q1, %1, 1;\n\t"
        "setp.gt.u32 q2, %1, 2;\n\t"
        "@p0 ld.shared.u8 a0, [%2];\n\t"
        "@p1 ld.shared.u8 a1, [%2+1];\n\t"
        "@p2 ld.shared.u8 a2, [%2+2];\n\t"
        "@q0 ld.shared.u8 b0, [%4];\n\t"
        "@q1 ld.shared.u8 b1, [%4+1];\n\t"
        "@q2 ld.shared.u8 b2, [%4+2];\n\t"
        "@p0 st.shared.u8 [%3], a0;\n\t"
        "@p1 st.shared.u8 [%3+1], a1;\n\t"
        "@p2 st.shared.u8 [%3+2], a2;\n\t"
        "@q0 st.shared.u8 [%5], b0;\n\t"
        "@q1 st.shared.u8 [%5+1], b1;\n\t"
        "@q2 st.shared.u8 [%5+2], b2;\n\t"
        "}" ::"r"(h), "r"(t), "r"(s), "r"(d), "r"(ts), "r"(td) : "memory");
}
// rem (0..) whole words: dst words da[0..min(rem,4)) = the bytes at source sa (word aligned) + sh / 8
ZPB_DEVINL void copy_words_ss(u32 rem, u32 sa, u32 sh, u32 da) {
    asm volatile(
        "{\n\t"
        ".reg .pred p0, p1, p2, p3;\n\t"
        ".reg .b32 a0, a1, a2, a3, a4, w0, w1, w2, w3;\n\t"
        "setp.gt.u32 p0, %0, 0;\n\t"
        "setp.gt.u32 p1, %0, 1;\n\t"
        "setp.gt.u32 p2, %0, 2;\n\t"
        "setp.gt.u32 p3, %0, 3;\n\t"
        "mov.b32 a2, 0;\n\tmov.b32 a3, 0;\n\tmov.b32 a4, 0;\n\tmov.b32 a0, 0;\n\tmov.b32 a1, 0;\n\t"
        "@p0 ld.shared.u32 a0, [%1];\n\t"
        "@p0 ld.shared.u32 a1, [%1+4];\n\t"
        "@p1 ld.shared.u32 a2, [%1+8];\n\t"
        "@p2 ld.shared.u32 a3, [%1+12];\n\t"
        "@p3 ld.shared.u32 a4, [%1+16];\n\t"
        "shf.r.wrap.b32 w0, a0, a1, %2;\n\t"
        "shf.r.wrap.b32 w1, a1, a2, %2;\n\t"
        "shf.r.wrap.b32 w2, a2, a3, %2;\n\t"
        "shf.r.wrap.b32 w3, a3, a4, %2;\n\t"
        "@p0 st.shared.u32 [%3], w0;\n\t"
        "@p1 st.shared.u32 [%3+4], w1;\n\t"
        "@p2 st.shared.u32 [%3+8], w2;\n\t"
        "@p3 st.shared.u32 [%3+12], w3;\n\t"
        "}" ::"r"(rem), "r"(sa), "r"(sh), "r"(da) : "memory");
}
ZPB_DEVINL void cp_async16_cg_if(bool p, u32 smem_addr, const void *gptr) {   // L2 only: coherent with this kernel's own stores
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q cp.async.cg.shared.global [%1], [%2], 16;\n\t}"
                 ::"r"((u32)p), "r"(smem_addr), "l"(gptr) : "memory");
}
// np (0..5) 16-byte pieces g[16k] -> s[16k], through L2
ZPB_DEVINL void cp_async_pieces_cg(u32 np, u32 s, const void *g) {
    asm volatile(
        "{\n\t"
        ".reg .pred p0, p1, p2, p3, p4;\n\t"
        "setp.gt.u32 p0, %0, 0;\n\t"
        "setp.gt.u32 p1, %0, 1;\n\t"
        "setp.gt.u32 p2, %0, 2;\n\t"
        "setp.gt.u32 p3, %0, 3;\n\t"
        "setp.gt.u32 p4, %0, 4;\n\t"
        "@p0 cp.async.cg.shared.global [%1], [%2], 16;\n\t"
        "@p1 cp.async.cg.shared.global [%1+16], [%2+16], 16;\n\t"
        "@p2 cp.async.cg.shared.global [%1+32], [%2+32], 16;\n\t"
        "@p3 cp.async.cg.shared.global [%1+48], [%2+48], 16;\n\t"
        "@p4 cp.async.cg.shared.global [%1+64], [%2+64], 16;\n\t"
        "}" ::"r"(np), "r"(s), "l"(g) : "memory");
}
#else
ZPB_DEVINL u32 lane_id() { return (u32)sim::cur_lane(); }
template <typename T> ZPB_DEVINL T keep_in_register(T v) { return v; }
ZPB_DEVINL void copy_edges_ss(u32 h, u32 t, u32 s, u32 d, u32 ts, u32 td) {
    u32 a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    for (u32 k = 0; k < 3; ++k) { if (h > k) a[k] = lds8(s + k); if (t > k) b[k] = lds8(ts + k); }
    for (u32 k = 0; k < 3; ++k) { if (h > k) sts8(d + k, a[k]); if (t > k) sts8(td + k, b[k]); }
}
ZPB_DEVINL void copy_words_ss(u32 rem, u32 sa, u32 sh, u32 da) {
    u32 a[5] = {0, 0, 0, 0, 0};
    if (rem > 0) { a[0] = lds32_loose(sa); a[1] = lds32_loose(sa + 4); }
    if (rem > 1) a[2] = lds32_loose(sa + 8);
    if (rem > 2) a[3] = lds32_loose(sa + 12);
    if (rem > 3) a[4] = lds32_loose(sa + 16);
    for (u32 k = 0; k < 4; ++k) if (rem > k) sts32(da + 4 * k, __funnelshift_r(a[k], a[k + 1], sh));
}
ZPB_DEVINL void cp_async_pieces_cg(u32 np, u32 s, const void *g) {
    for (u32 k = 0; k < np; ++k) sim::cp_async(s + 16 * k, (const u8 *)g + 16 * k, 16, 16);
}
#endif
