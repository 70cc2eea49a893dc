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

// ---- the zstd encoder (zpack_b200/csrc/zstd_encode.cuh) is plain serial code per window / per block: called directly,
// in the order of the three kernels.  lz: one block's LZ4 payload (csize bytes), winop: its window offsets, raw_len: the
// block's size.  bodies: ZE_SLOT bytes, window w's body at ZE_OFF(winop[w], w); zbody[16]: sizes (0: none).
// Returns 1 when the block is encoded, 0 when it has to be stored raw.
#include "../../zpack_b200/csrc/zstd_encode.cuh"
extern "C" int sim_zstd_encode_block(const uint8_t *lz, uint32_t csize, uint32_t raw_len, const uint32_t *winop, uint8_t *bodies,
                                     uint32_t *zbody, int use_huffman /* bit 0: Huffman literals, bit 1: fitted FSE tables */, uint32_t *modes) {
    static ZeTables T;
    static bool built = false;
    if (!built) { ze_build_tables(T); built = true; }
    const u32 nwin = (raw_len - 12u) / 4096u + 1u;
    std::vector<u8> zlit(ZE_LITSLOT + 64, 0);
    std::vector<u64> zseq(16 * ZE_WIN_SEQ);
    u32 lit_n[16] = {0}, seq_n[16] = {0};
    std::vector<u32> hist(256, 0), hll(36, 0), hof(32, 0), hml(53, 0);
    u32 total = 0, nseq_total = 0;
    for (u32 w = 0; w < nwin; ++w) {                                                            // stage A
        const u32 begin = winop[w], end = winop[w + 1], tail_end = w + 1 == nwin ? csize : end;
        if (!(begin <= end && tail_end <= csize && end <= tail_end)) return 0;
        if (!ze_parse_range(lz, begin, end, tail_end, zlit.data() + ZE_LOFF(begin, w), zseq.data() + w * ZE_WIN_SEQ, ZE_WIN_SEQ, &lit_n[w], &seq_n[w]))
            return 0;
        for (u32 i = 0; i < lit_n[w]; ++i) ++hist[zlit[ZE_LOFF(begin, w) + i]];
        total += lit_n[w];
        for (u32 i = 0; i < seq_n[w]; ++i) {
            const u64 r = zseq[w * ZE_WIN_SEQ + i];
            ++hll[ze_ll_code((u32)(r & 0xFFFF))]; ++hml[ze_ml_code((u32)((r >> 16) & 0xFFFF))]; ++hof[ze_highbit((u32)(r >> 32))];
        }
        nseq_total += seq_n[w];
    }
    ZeHuf H;                                                                                    // stage B
    H.desc_len = 0;
    if ((use_huffman & 1) && total >= 256) ze_huf_build(hist.data(), H);
    static ZeBlockTabs B;
    B.valid = 0;
    if ((use_huffman & 2) && nseq_total >= ZE_TABS_MIN) {
        ze_fit(B.ll, B.desc[0], &B.desc_len[0], hll.data(), 36, nseq_total, 9);
        ze_fit(B.of, B.desc[1], &B.desc_len[1], hof.data(), 32, nseq_total, 8);
        ze_fit(B.ml, B.desc[2], &B.desc_len[2], hml.data(), 53, nseq_total, 9);
        B.valid = 1;
    }
    u32 sum = 0;
    for (u32 w = 0; w < 16; ++w) zbody[w] = 0;
    for (u32 w = 0; w < nwin; ++w) {                                                            // stage C
        const u32 begin = winop[w], end = winop[w + 1], tail_end = w + 1 == nwin ? csize : end;
        const u32 mode = ze_lit_mode(w, lit_n, nwin, H.desc_len != 0);
        if (modes) modes[w] = mode;
        const u32 seqmode = ze_seq_mode(w, seq_n, nwin, B.valid != 0);
        const u32 z = ze_emit_range(zlit.data() + ZE_LOFF(begin, w), lit_n[w], zseq.data() + w * ZE_WIN_SEQ, seq_n[w], mode, H,
                                    bodies + ZE_OFF(begin, w), (tail_end - begin) + ((tail_end - begin) >> 2) + 20u, T, seqmode, B);
        if (z == ZE_FAIL) return 0;
        zbody[w] = z;
        if (z) sum += z + 3;
    }
    return sum && sum < raw_len;
}
