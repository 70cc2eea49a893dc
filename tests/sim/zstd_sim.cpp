// tests/sim/zstd_sim.cpp — TEST HARNESS ONLY.
// Compiles the product's zstd decoder control logic (zpack_b200/csrc/zstd_decode.cuh) for the host
// with a 1-lane warp (-DZPB_HOST_SIM) so that the CPU-only test suite can check every table build,
// bit read and sequence rule against the oracle before the kernel ever sees a GPU.  Nothing under
// zpack_b200/ links or loads this.
#define ZPB_HOST_SIM 1
#include "../../zpack_b200/csrc/zstd_decode.cuh"
#include <cstdlib>

struct NoHash { void advance(u64, const ZWarp &) {} };

extern "C" int zs_sim_decode(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *out_len) {
    static thread_local ZstdShared S;
    static thread_local u8 *scratch = nullptr;
    if (!scratch) scratch = (u8 *)malloc(ZS_LIT_SCRATCH);
    // keep the aligned-word reader inside one allocation: copy the input with slack on both sides
    u8 *buf = (u8 *)malloc(n + 16);
    memset(buf, 0xA5, n + 16);
    if (n) memcpy(buf + 5, src, n);   // odd alignment on purpose
    ZWarp w;
    NoHash h;
    u64 produced = 0;
    int rc = zstd_decode_entry(w, S, scratch, buf + 5, n, dst, cap, h, &produced);
    free(buf);
    *out_len = produced;
    return rc ? 13 : 0;
}
