// sim_rt.h — TEST INFRASTRUCTURE: a CPU emulation of the CUDA execution model, just wide enough to run the kernels of
// zpack_b200/csrc unchanged (compiled by g++ with -DZPB_SIM) so that their control logic can be checked against the
// oracle in a container that has no GPU.  It is never part of the product: libzpack_b200.so is built by nvcc without
// ZPB_SIM and contains none of this.
//
// Model: one CTA at a time; every CUDA thread is a fiber (ucontext) on one OS thread.  A fiber runs until it reaches a
// warp collective (__shfl_sync, __ballot_sync, __syncwarp ...), a CTA barrier or a spin-wait, where it parks; a
// collective releases its lanes when every lane named in the mask has arrived.  Between collectives the runnable
// lanes execute one after another in a RANDOM order (seeded), so code that relies on lock-step execution without a
// __syncwarp() computes wrong results here instead of passing by accident.  Shared-memory accesses go through checked
// accessors: bounds, plus a per-byte record of the last writer / reader that reports read-after-write,
// write-after-write and write-after-read between different lanes of a warp with no __syncwarp() in between.
// cp.async, cp.async.bulk and mbarrier are modelled with deferred completion: the data lands when the matching wait
// is executed, not before, so reading a staging buffer before waiting on it reads stale bytes.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace sim {

struct Dim3 {
    unsigned x, y, z;
    Dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

struct PendingCopy { uint32_t dst; const void *src; uint32_t bytes; uint32_t group; };
struct Bulk { uint32_t dst; const void *src; uint32_t bytes; uint32_t bar; };

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
    bool waiting = false;       // parked at a warp collective
    uint32_t wait_mask = 0;
    bool bar_wait = false;      // parked at __syncthreads
    bool am_wait = false;       // inside __activemask(): waits for the other lanes to converge here, exit or block
    uint32_t am_result = 0;
    uint64_t value = 0;         // value published at the collective
    uint64_t rx[32];            // values of every lane, delivered on release
    std::vector<PendingCopy> cpq;   // cp.async: per-thread groups (PTX semantics)
    uint32_t cp_group = 0;
};

struct State {
    Dim3 grid, block;
    std::vector<Fiber> fibers;
    std::vector<uint64_t> epoch;    // per warp: increments at __syncwarp / bar.sync
    std::vector<uint8_t> smem;
    std::vector<uint64_t> sh_w, sh_r;   // per smem byte: (epoch << 16) | (warp << 6) | (lane + 1)
    int cur = -1;
    ucontext_t sched;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    bool race_check = true;
    uint64_t races = 0;
    std::function<void()> body;
    std::vector<Bulk> bulks;
    size_t stack_bytes = 256 << 10;
};

inline State &S() { static State s; return s; }
static const uint32_t SMEM_BASE = 0x1000000u;   // shared-window address of smem[0] (window addresses are never 0)

inline uint64_t rnd() {
    State &s = S();
    s.rng ^= s.rng << 13; s.rng ^= s.rng >> 7; s.rng ^= s.rng << 17;
    return s.rng;
}

struct Idx { unsigned x = 0, y = 0, z = 0; };
inline Idx &tidx() { static Idx v; return v; }
inline Idx &bidx() { static Idx v; return v; }
inline Dim3 &bdim() { static Dim3 v; return v; }
inline Dim3 &gdim() { static Dim3 v; return v; }

inline void set_ids(int f) {
    State &s = S();
    tidx().x = (unsigned)f % s.block.x;
    tidx().y = ((unsigned)f / s.block.x) % s.block.y;
    tidx().z = (unsigned)f / (s.block.x * s.block.y);
}
inline int cur_lane() { return S().cur & 31; }
inline int cur_warp() { return S().cur >> 5; }

inline void park() {   // back to the scheduler; returns when this fiber is picked again
    State &s = S();
    int me = s.cur;
    swapcontext(&s.fibers[me].ctx, &s.sched);
    s.cur = me;
    set_ids(me);
}

// __activemask(): the lanes that are converged at this point.  Emulated as "every lane of the warp that reaches an
// __activemask() before it blocks or exits": one of the groupings the hardware may produce.
inline uint32_t activemask() {
    State &s = S();
    const int me = s.cur, w = me >> 5, n = (int)s.fibers.size();
    s.fibers[me].am_wait = true;
    for (;;) {
        if (!s.fibers[me].am_wait) return s.fibers[me].am_result;
        bool settled = true;
        uint32_t m = 0;
        for (int l = 0; l < 32; ++l) {
            const int g = w * 32 + l;
            if (g >= n) continue;
            const Fiber &f = s.fibers[g];
            if (f.am_wait) m |= 1u << l;
            else if (!(f.done || f.waiting || f.bar_wait)) settled = false;
        }
        if (settled) {
            for (int l = 0; l < 32; ++l)
                if ((m >> l) & 1u) { s.fibers[w * 32 + l].am_wait = false; s.fibers[w * 32 + l].am_result = m; }
            return m;
        }
        park();
    }
}

// Park at a collective; afterwards rx[l] (returned pointer) holds what lane l published.
inline const uint64_t *collective(uint32_t mask, uint64_t v, bool is_sync) {
    State &s = S();
    int me = s.cur, w = me >> 5, l = me & 31, n = (int)s.fibers.size();
    if (!((mask >> l) & 1u)) { fprintf(stderr, "sim: lane %d not in its own collective mask %08x\n", l, mask); abort(); }
    Fiber &f = s.fibers[me];
    f.value = v; f.waiting = true; f.wait_mask = mask;
    bool all = true;
    for (int k = 0; k < 32; ++k)
        if ((mask >> k) & 1u) {
            int g = w * 32 + k;
            if (g >= n || s.fibers[g].done) {
                fprintf(stderr, "sim: collective (mask %08x) names lane %d which has exited or does not exist\n", mask, k);
                abort();
            }
            if (!(s.fibers[g].waiting && s.fibers[g].wait_mask == mask)) all = false;
        }
    if (all) {
        uint64_t box[32] = {0};
        for (int k = 0; k < 32; ++k)
            if ((mask >> k) & 1u) box[k] = s.fibers[w * 32 + k].value;
        for (int k = 0; k < 32; ++k)
            if ((mask >> k) & 1u) {
                Fiber &g = s.fibers[w * 32 + k];
                memcpy(g.rx, box, sizeof box);
                g.waiting = false;
            }
        if (is_sync) s.epoch[w]++;
        return f.rx;
    }
    while (s.fibers[me].waiting) park();
    return s.fibers[me].rx;
}

inline void cta_barrier() {
    State &s = S();
    int me = s.cur;
    s.fibers[me].bar_wait = true;
    while (s.fibers[me].bar_wait) park();
}

// ---- shared memory
inline uint8_t *smem_ptr(uint32_t a, uint32_t n, const char *what) {
    State &s = S();
    if (a < SMEM_BASE || (uint64_t)a - SMEM_BASE + n > s.smem.size()) {
        fprintf(stderr, "sim: shared-memory %s out of bounds: window address 0x%x + %u (size %zu), cta %u thread %d\n",
                what, a, n, s.smem.size(), bidx().x, s.cur);
        abort();
    }
    return s.smem.data() + (a - SMEM_BASE);
}
inline void note(uint32_t a, uint32_t n, bool write) {
    State &s = S();
    if (!s.race_check) return;
    const int w = cur_warp(), l = cur_lane();
    const uint64_t tag = (s.epoch[w] << 16) | ((uint64_t)w << 6) | (uint64_t)(l + 1);
    for (uint32_t k = 0; k < n; ++k) {
        const size_t i = (size_t)(a - SMEM_BASE) + k;
        const uint64_t pw = s.sh_w[i], pr = s.sh_r[i];
        const bool w_conf = pw && (pw >> 6) == (tag >> 6) && pw != tag;
        const bool r_conf = pr && (pr >> 6) == (tag >> 6) && pr != tag;
        if (w_conf || (write && r_conf)) {
            if (s.races < 20)
                fprintf(stderr, "sim: shared-memory race (%s after %s) at window 0x%x, warp %d lane %d vs lane %d, no __syncwarp between\n",
                        write ? "write" : "read", w_conf ? "write" : "read", a + k, w, l,
                        (int)((w_conf ? pw : pr) & 63) - 1);
            s.races++;
        }
        if (write) s.sh_w[i] = tag; else s.sh_r[i] = tag;
    }
}
inline uint32_t ld_shared(uint32_t a, uint32_t n, bool loose = false) {
    uint8_t *p = smem_ptr(a, n, "load");
    if (a % n) { fprintf(stderr, "sim: misaligned %u-byte shared load at 0x%x\n", n, a); abort(); }
    if (!loose) note(a, n, false);
    uint32_t v = 0;
    memcpy(&v, p, n);
    return v;
}
inline void st_shared(uint32_t a, uint32_t n, uint32_t v) {
    uint8_t *p = smem_ptr(a, n, "store");
    if (a % n) { fprintf(stderr, "sim: misaligned %u-byte shared store at 0x%x\n", n, a); abort(); }
    note(a, n, true);
    memcpy(p, &v, n);
}

// ---- cp.async (non-bulk): per-thread groups, completion deferred to the wait
inline void cp_async(uint32_t dst, const void *src, uint32_t bytes, uint32_t src_bytes) {
    smem_ptr(dst, bytes, "cp.async destination");
    Fiber &f = S().fibers[S().cur];
    if (src_bytes < bytes) {   // zero fill of the remainder happens at completion as well
        static const uint8_t zeros[16] = {0};
        f.cpq.push_back({dst + src_bytes, zeros, bytes - src_bytes, f.cp_group});
    }
    if (src_bytes) f.cpq.push_back({dst, src, src_bytes, f.cp_group});
}
inline void cp_async_commit() { S().fibers[S().cur].cp_group++; }
inline void cp_async_wait(uint32_t allow_pending) {
    Fiber &f = S().fibers[S().cur];
    const uint32_t committed = f.cp_group;   // groups [0, committed) exist; the newest `allow_pending` may stay in flight
    const uint32_t upto = committed > allow_pending ? committed - allow_pending : 0;
    size_t w = 0;
    for (size_t i = 0; i < f.cpq.size(); ++i) {
        PendingCopy &c = f.cpq[i];
        if (c.group < upto) {
            memcpy(smem_ptr(c.dst, c.bytes, "cp.async landing"), c.src, c.bytes);
            note(c.dst, c.bytes, true);
        } else f.cpq[w++] = c;
    }
    f.cpq.resize(w);
}

// ---- mbarrier + cp.async.bulk: barrier word in shared memory = {phase bit 63, pending arrivals 20 bits, tx bytes}
struct MBar { uint32_t phase; uint32_t arrivals_left; int64_t tx; uint32_t init_count; };
inline std::vector<std::pair<uint32_t, MBar>> &mbars() { static std::vector<std::pair<uint32_t, MBar>> v; return v; }
inline MBar &mbar(uint32_t a) {
    for (auto &p : mbars()) if (p.first == a) return p.second;
    fprintf(stderr, "sim: mbarrier at 0x%x used before init\n", a); abort();
}
inline void mbar_init(uint32_t a, uint32_t count) {
    smem_ptr(a, 8, "mbarrier");
    for (auto &p : mbars()) if (p.first == a) { p.second = MBar{0, count, 0, count}; return; }
    mbars().push_back({a, MBar{0, count, 0, count}});
}
inline void mbar_try_complete(uint32_t a) {
    MBar &b = mbar(a);
    if (b.arrivals_left == 0 && b.tx == 0) { b.phase ^= 1u; b.arrivals_left = b.init_count; }
}
inline void mbar_arrive_expect_tx(uint32_t a, uint32_t bytes) {
    MBar &b = mbar(a);
    if (b.arrivals_left == 0) { fprintf(stderr, "sim: mbarrier 0x%x over-arrived\n", a); abort(); }
    b.tx += bytes;
    b.arrivals_left--;
    mbar_try_complete(a);
}
inline void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    if ((dst & 15u) || ((uintptr_t)src & 15u) || (bytes & 15u) || bytes == 0) {
        fprintf(stderr, "sim: cp.async.bulk needs 16-byte aligned addresses and size (dst 0x%x src %p bytes %u)\n", dst, src, bytes);
        abort();
    }
    smem_ptr(dst, bytes, "cp.async.bulk destination");
    S().bulks.push_back({dst, src, bytes, bar});
}
// one poll of try_wait.parity: lands the copies that signal this barrier (deferred completion), then tests the phase
inline bool mbar_test_wait(uint32_t a, uint32_t parity) {
    State &s = S();
    size_t w = 0;
    for (size_t i = 0; i < s.bulks.size(); ++i) {
        Bulk c = s.bulks[i];
        if (c.bar == a) {
            memcpy(smem_ptr(c.dst, c.bytes, "cp.async.bulk landing"), c.src, c.bytes);
            const bool rc = s.race_check;
            s.race_check = false;   // the async proxy is not a lane; ordering is the barrier's job
            s.race_check = rc;
            // forget the lane records of the overwritten bytes: the data is new
            for (uint32_t k = 0; k < c.bytes; ++k) { s.sh_w[(c.dst - SMEM_BASE) + k] = 0; s.sh_r[(c.dst - SMEM_BASE) + k] = 0; }
            MBar &b = mbar(a);
            b.tx -= c.bytes;
            mbar_try_complete(a);
        } else s.bulks[w++] = c;
    }
    s.bulks.resize(w);
    return mbar(a).phase != parity;   // phase `parity` has completed when the current phase bit differs
}

// ---- launch
inline void fiber_entry() {
    State &s = S();
    int me = s.cur;
    set_ids(me);
    s.body();
    s.fibers[me].done = true;
    // a finished lane may complete a pending CTA barrier count; the scheduler re-evaluates
    swapcontext(&s.fibers[me].ctx, &s.sched);
}

inline void launch(Dim3 grid, Dim3 block, size_t smem_bytes, std::function<void()> body, uint64_t seed = 1) {
    State &s = S();
    s.grid = grid; s.block = block;
    gdim() = grid; bdim() = block;
    s.body = body;
    s.rng = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    const int nthreads = (int)(block.x * block.y * block.z);
    for (unsigned cta = 0; cta < grid.x * grid.y * grid.z; ++cta) {
        bidx().x = cta % grid.x; bidx().y = (cta / grid.x) % grid.y; bidx().z = cta / (grid.x * grid.y);
        s.smem.assign(smem_bytes + 64, 0xCD);
        s.sh_w.assign(smem_bytes + 64, 0); s.sh_r.assign(smem_bytes + 64, 0);
        s.epoch.assign((nthreads + 31) / 32, 1);
        mbars().clear();
        s.bulks.clear();
        s.fibers.clear();
        s.fibers.resize(nthreads);
        for (int t = 0; t < nthreads; ++t) {
            Fiber &f = s.fibers[t];
            f.stack = (char *)malloc(s.stack_bytes);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = s.stack_bytes;
            f.ctx.uc_link = &s.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        for (;;) {
            // runnable lanes
            int runnable[1024], nr = 0, alive = 0, at_bar = 0;
            for (int t = 0; t < nthreads; ++t) {
                Fiber &f = s.fibers[t];
                if (f.done) continue;
                ++alive;
                if (f.bar_wait) { ++at_bar; continue; }
                if (f.waiting) continue;
                if (nr < 1024) runnable[nr++] = t;
            }
            if (!alive) break;
            if (!nr) {
                if (at_bar == alive) {   // __syncthreads releases
                    for (auto &f : s.fibers) f.bar_wait = false;
                    for (auto &e : s.epoch) e++;
                    continue;
                }
                fprintf(stderr, "sim: deadlock in cta %u: %d threads alive, %d at the CTA barrier, none runnable\n", cta, alive, at_bar);
                abort();
            }
            int pick = runnable[rnd() % nr];
            s.cur = pick;
            swapcontext(&s.sched, &s.fibers[pick].ctx);
        }
        for (auto &f : s.fibers) free(f.stack);
        s.fibers.clear();
    }
}

}  // namespace sim
