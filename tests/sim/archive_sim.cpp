// archive_sim.cpp — TEST INFRASTRUCTURE: runs the container kernels of zpack_b200/csrc/archive_kernels.cuh on the CPU
// emulation (sim_rt.h) so that tests/test_archive_sim.py can compare them with the host mirror (zpack_b200/container.py)
// and the unmodified reference without a GPU.
// Built by tests/test_archive_sim.py:  g++ -O1 -g -DZPB_SIM -shared -fPIC archive_sim.cpp -o libarchive_sim.so
#include <cstring>
#include <vector>

#include "../../include/zpack_b200.h"
#include "../../zpack_b200/csrc/archive_kernels.cuh"

// layout (assign != 0: offsets from 10, back to back) + directory + payload copy, in the order of zpb_archive_build_device;
// arch == NULL: layout + copy only (zpb_copy_entries_device, offsets are the caller's)
extern "C" int sim_archive_build(const uint8_t *src, ArcEntry *e, uint64_t n, const uint8_t *names, uint8_t *dst, int with_cdr,
                                 uint64_t cdr_off, uint64_t block, int scan_threads, int grid, uint64_t seed, uint64_t *totals) {
    std::vector<u64> rec(n + 1), chunk(n + 2);
    u64 *rp = rec.data(), *cp = chunk.data();
    const u64 ntiles = n ? (n + 4 * (u64)scan_threads - 1) / (4 * (u64)scan_threads) : 1;
    std::vector<u64> chain(ntiles * 4, 0);
    u64 *chp = chain.data();
    for (int k = 0; k < 8; ++k) totals[k] = 0;
    sim::launch(sim::Dim3((unsigned)ntiles), sim::Dim3((unsigned)scan_threads), 37 * 24, [&] {
        arc_layout_body(e, n, ARC_DATA_START, with_cdr ? 1u : 0u, rp, cp, totals, chp);
    }, seed);
    if (with_cdr)
        sim::launch(sim::Dim3((unsigned)grid), sim::Dim3(64), 0, [&] { arc_cdr_body(dst, e, n, names, rp, cdr_off, block, 1u); }, seed);
    const u64 nchunks = totals[2];
    if (nchunks) {
        std::vector<ArcChunk> work(nchunks);
        ArcChunk *wp = work.data();
        sim::launch(sim::Dim3((unsigned)((nchunks + 31) / 32)), sim::Dim3(32), 0, [&] { arc_chunks_body(e, n, cp, nchunks, wp); }, seed);
        unsigned long long *next = (seed & 2) ? (unsigned long long *)(totals + 4) : nullptr;
        const u32 mask = (seed & 4) ? 127u : 15u;
        sim::launch(sim::Dim3((unsigned)grid), sim::Dim3(seed & 4 ? 128 : 64), 16, [&] {
            if (seed & 1) arc_copy_body<4>(src, dst, wp, nchunks, next, mask); else arc_copy_body<8>(src, dst, wp, nchunks, next, mask);
        }, seed);
    }
    return 0;
}

// the five directory kernels in the order of zpb_archive_open_device; returns the number of records that fit the block
extern "C" uint64_t sim_archive_open(const uint8_t *body, uint64_t B, ArcEntry *out, uint64_t count, uint64_t seed) {
    const u64 ntiles = (B + CDR_T - 1) / CDR_T, nsuper = (B + CDR_ST - 1) / CDR_ST;
    std::vector<u32> jump(B + 4), jump2(nsuper * CDR_T), cnt2(nsuper * CDR_T), sup(2 * nsuper, CDR_END), tile(2 * ntiles, CDR_END);
    std::vector<u8> cnt(B + 4);
    u64 found = 0;
    u32 *jp = jump.data(), *j2 = jump2.data(), *c2 = cnt2.data(), *sp = sup.data(), *si = sp + nsuper, *tp = tile.data(), *ti = tp + ntiles;
    u8 *cn = cnt.data();
    u64 *fp = &found;
    sim::launch(sim::Dim3((unsigned)ntiles), sim::Dim3(64), CDR_T + 16, [&] { cdr_tile_body(body, B, jp, cn); }, seed);
    sim::launch(sim::Dim3((unsigned)((nsuper * CDR_T + 63) / 64)), sim::Dim3(64), 0, [&] { cdr_super_body(B, jp, cn, j2, c2, nsuper); }, seed);
    sim::launch(sim::Dim3(1), sim::Dim3(32), 0, [&] { cdr_chain_body(B, jp, cn, j2, c2, sp, si, fp); }, seed);
    sim::launch(sim::Dim3((unsigned)((nsuper + 31) / 32)), sim::Dim3(32), 0, [&] { cdr_anchor_body(B, jp, cn, sp, si, tp, ti, nsuper); }, seed);
    sim::launch(sim::Dim3((unsigned)((ntiles + 31) / 32)), sim::Dim3(32), 0, [&] { cdr_emit_body(body, B, tp, ti, out, count, ntiles); }, seed);
    return found;
}
