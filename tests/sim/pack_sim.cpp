// pack_sim.cpp — TEST INFRASTRUCTURE: runs the block compressor of zpack_b200/csrc/pack_blocks.cuh on the CPU emulation
// (sim_rt.h) so that tests/test_pack_sim.py can decode its output with the oracle without a GPU.
// Built by tests/test_pack_sim.py:  g++ -O1 -g -DZPB_SIM -shared -fPIC pack_sim.cpp -o libpack_sim.so
#include <cstring>
#include <vector>

#include "../../include/zpack_b200.h"
#include "../../zpack_b200/csrc/pack_blocks.cuh"

// blocks: (src_off, len) pairs; scratch: nblocks * 65536 bytes; csize: nblocks
extern "C" int sim_lz4_pack_blocks(const uint8_t *in, const uint64_t *src_off, const uint32_t *len, uint32_t nblocks,
                                   uint8_t *scratch, uint32_t *csize, int grid, uint64_t seed, uint32_t *winop) {
    std::vector<PackBlock> pb(nblocks);
    for (u32 i = 0; i < nblocks; ++i) { pb[i].src_off = src_off[i]; pb[i].len = len[i]; pb[i].pad = 0; }
    u32 counter = 0;
    const PackBlock *dpb = pb.data();
    u32 *cnt = &counter;
    sim::launch(sim::Dim3((unsigned)grid), sim::Dim3(32 * P2_WARPS), P2_SMEM, [&] {
        lz4_pack_blocks_body(in, dpb, nblocks, cnt, scratch, csize, winop);
    }, seed);
    return 0;
}

// ---- the zstd encoder (zpack_b200/csrc/zstd_encode.cuh) is plain serial code per window of a block: called directly
#include "../../zpack_b200/csrc/zstd_encode.cuh"
extern "C" uint32_t sim_zstd_encode_range(const uint8_t *lz, uint32_t begin, uint32_t end, uint32_t tail_end, uint8_t *out, uint32_t cap) {
    static ZeTables T;
    static bool built = false;
    if (!built) { ze_build_tables(T); built = true; }
    std::vector<u64> seq(ZE_WIN_SEQ);
    return ze_encode_range(lz, begin, end, tail_end, out, cap, seq.data(), ZE_WIN_SEQ, T);
}
