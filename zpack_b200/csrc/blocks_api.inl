// blocks_api.inl — intra-entry parallelism for ONE large LZ4 entry written with independent blocks
// (BASELINE config C5).  Included by zpb_api.cu.
//
//   zpb_lz4_frame_index       host: frame header + the block-header chain (lz4frame.c:1113-1205,
//                             1503-1531).  Pure framing — each 4-byte header sits where the previous
//                             block ended, a serial pointer chase that costs the host ~20 ms for the
//                             262 144 blocks of a 16 GiB entry and would cost one GPU thread ~0.2 s.
//   zpb_unpack_blocks_device  a contiguous run of those blocks (one GPU's shard) through the same
//                             scan -> parse -> exec kernels as whole entries, one warp per block, every
//                             finished KiB leaving its XXH3 stripe sums instead of entering a digest.
//   zpb_blocks_digest         the XXH3 scramble chain over the shard's KiBs (xxh3_chain_kernel), starting
//                             from the previous shard's 64-byte accumulator state; the last shard also
//                             folds the tail and produces the entry digest.

extern "C" int zpb_lz4_frame_index(const uint8_t *h_frame, uint64_t comp_size, uint64_t archive_off,
                                   zpb_block *blocks, uint64_t cap, uint64_t *n_blocks, uint32_t *block_size,
                                   uint64_t *content_size) {
    if (!h_frame || !n_blocks || !block_size) return fail(nullptr, ZPB_E_ARG, "null argument");
    *n_blocks = 0; *block_size = 0;
    if (content_size) *content_size = ~0ull;
    auto rd32 = [&](u64 p) { return (u32)h_frame[p] | ((u32)h_frame[p + 1] << 8) | ((u32)h_frame[p + 2] << 16) | ((u32)h_frame[p + 3] << 24); };
    if (comp_size < 11 || rd32(0) != 0x184D2204u) return ZPB_INDEX_UNSUPPORTED;
    const u32 flg = h_frame[4], bd = h_frame[5];
    // version 01, reserved bits clear, B.Indep = 1, no block / content checksums (their XXH32 chains are not
    // part of this path), 64 KB blocks (the fast kernels' block size)
    if (((flg >> 6) & 3) != 1 || (flg & 2) || !((flg >> 5) & 1) || ((flg >> 4) & 1) || ((flg >> 2) & 1))
        return ZPB_INDEX_UNSUPPORTED;
    if ((bd >> 7) || ((bd >> 4) & 7) != 4 || (bd & 15)) return ZPB_INDEX_UNSUPPORTED;
    const bool has_size = (flg >> 3) & 1, has_dict = flg & 1;
    const u64 hsize = 7 + (has_size ? 8 : 0) + (has_dict ? 4 : 0);
    if (comp_size < hsize + 4) return ZPB_INDEX_UNSUPPORTED;
    // header checksum byte (lz4frame.c:294-298, 1184-1186) — XXH32 of the descriptor, host arithmetic on <= 14 bytes
    {
        const u8 *p = h_frame + 4;
        const u32 len = (u32)hsize - 5;
        const u32 P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, P4 = 0x27D4EB2Fu, P5 = 0x165667B1u;
        auto rotl = [](u32 v, int r) { return (v << r) | (v >> (32 - r)); };
        u32 h = P5 + len, i = 0;   // len < 16: no stripe loop
        for (; i + 4 <= len; i += 4) {
            u32 w = (u32)p[i] | ((u32)p[i + 1] << 8) | ((u32)p[i + 2] << 16) | ((u32)p[i + 3] << 24);
            h = rotl(h + w * P3, 17) * P4;
        }
        for (; i < len; ++i) h = rotl(h + p[i] * P5, 11) * P1;
        h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
        if (((h >> 8) & 0xFF) != h_frame[hsize - 1]) return ZPB_INDEX_UNSUPPORTED;
    }
    if (has_size && content_size) {
        u64 v = 0;
        for (int b = 7; b >= 0; --b) v = (v << 8) | h_frame[6 + b];
        *content_size = v;
    }
    u64 ip = hsize, nb = 0;
    for (;;) {
        if (comp_size - ip < 4) return ZPB_INDEX_UNSUPPORTED;
        const u32 bh = rd32(ip);
        ip += 4;
        if (bh == 0) break;
        const u32 bsz = bh & 0x7FFFFFFFu;
        if (bsz == 0 || bsz > 65536u || bsz > comp_size - ip) return ZPB_INDEX_UNSUPPORTED;
        if (blocks && nb < cap) {
            blocks[nb].src_off = archive_off + ip;
            blocks[nb].comp_size = bsz;
            blocks[nb].flags = (bh >> 31) ? ZPB_BLK_STORED : 0u;
        }
        ip += bsz;
        ++nb;
    }
    if (ip != comp_size) return ZPB_INDEX_UNSUPPORTED;   // another frame / trailing bytes: the general path's business
    *n_blocks = nb;
    *block_size = 65536u;
    return ZPB_OK;
}

extern "C" int zpb_unpack_blocks_device(zpb_ctx *ctx, const uint8_t *d_archive, uint64_t archive_size,
                                        uint8_t *d_out, uint64_t out_size, const zpb_block *blocks, uint64_t n,
                                        uint32_t block_size, uint64_t shard_uncomp_size, int32_t *status,
                                        void *stream) {
    if (!ctx || !d_archive || (!d_out && out_size) || (!blocks && n) || !status)
        return fail(ctx, ZPB_E_ARG, "null argument");
    if (block_size != 65536u) return fail(ctx, ZPB_E_ARG, "block_size must be 65536");
    if (n == 0 || n > 0x7fffffffull) return fail(ctx, ZPB_E_ARG, "block count out of range");
    if (shard_uncomp_size > n * (u64)block_size || shard_uncomp_size <= (n - 1) * (u64)block_size ||
        shard_uncomp_size > out_size)
        return fail(ctx, ZPB_E_ARG, "shard size does not match its block count (every block but the last is full)");
    if ((uintptr_t)d_out & 15) return fail(ctx, ZPB_E_ARG, "d_out must be 16-byte aligned");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->blk_uncomp = 0;
    const u64 kibs = (shard_uncomp_size + 1023) >> 10;
    if (!ctx->d_partials.ensure((kibs + 1) * 64)) return fail(ctx, ZPB_E_NOMEM, "stripe-sum scratch allocation failed");
    std::vector<zpb_entry> pe(n);
    for (u64 k = 0; k < n; ++k) {
        zpb_entry &e = pe[k];
        const u64 usz = k + 1 < n ? (u64)block_size : shard_uncomp_size - (n - 1) * (u64)block_size;
        e.src_off = blocks[k].src_off;
        e.comp_size = blocks[k].comp_size;
        e.dst_off = k * (u64)block_size;
        e.dst_cap = usz;
        e.uncomp_size = usz;
        e.hash = 0;
        e.method = ZPB_M_LZ4_BLOCK;
        e.flags = (blocks[k].flags & ZPB_BLK_STORED) ? ZPB_F_STORED_BLK : 0u;
        e.reserved = k * (u64)(block_size >> 10);
    }
    std::vector<int32_t> st(n);
    const int fast_was = ctx->fast;
    ctx->fast = 1;
    ctx->cur_partials = (u64 *)ctx->d_partials.p;
    int rc = unpack_device_impl(ctx, d_archive, archive_size, d_out, out_size, pe.data(), n, st.data(), nullptr, s);
    ctx->cur_partials = nullptr;
    ctx->fast = fast_was;
    if (rc != ZPB_OK) return rc;
    int32_t verdict = ZPB_ST_OK;
    for (u64 k = 0; k < n; ++k)
        if (st[k] != ZPB_ST_OK) { verdict = st[k] == ZPB_ST_OFFSET_INVALID ? ZPB_ST_OFFSET_INVALID : ZPB_ST_NOT_AVAILABLE; break; }
    *status = verdict;
    if (verdict == ZPB_ST_OK) ctx->blk_uncomp = shard_uncomp_size;
    return ZPB_OK;
}

extern "C" int zpb_blocks_digest(zpb_ctx *ctx, const uint64_t *acc_in, uint64_t *acc_out, uint64_t shard_pos,
                                 uint64_t total_size, const uint8_t *d_out, uint64_t *digest, void *stream) {
    if (!ctx) return fail(ctx, ZPB_E_ARG, "null argument");
    if (ctx->blk_uncomp == 0) return fail(ctx, ZPB_E_ARG, "no decoded shard: call zpb_unpack_blocks_device first (status must be 0)");
    const u64 shard = ctx->blk_uncomp;
    if ((shard_pos & 1023) || shard_pos > total_size || shard > total_size - shard_pos)
        return fail(ctx, ZPB_E_ARG, "shard position / size outside the entry");
    const bool final = shard_pos + shard == total_size;
    if (!final && (shard & 1023)) return fail(ctx, ZPB_E_ARG, "only the entry's last shard may end off a KiB boundary");
    if (final && !digest) return fail(ctx, ZPB_E_ARG, "the last shard needs a digest pointer");
    u64 nscr = shard >> 10, tail_pos = 0;
    if (final) {
        if (!d_out) return fail(ctx, ZPB_E_ARG, "the last shard needs its decoded bytes (d_out)");
        if (total_size <= 240) {
            if (shard_pos) return fail(ctx, ZPB_E_ARG, "an entry of <= 240 bytes cannot be split");
            nscr = 0;
        } else {
            const u64 full_blocks = (total_size - 1) >> 10;
            nscr = std::min<u64>(nscr, full_blocks - (shard_pos >> 10));
            const u64 need = std::min<u64>(full_blocks << 10, total_size - 64);
            if (need < shard_pos)
                return fail(ctx, ZPB_E_ARG, "the last shard must hold the entry's final 64 bytes and its last partial KiB");
        }
        tail_pos = shard_pos;
    }
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (!ctx->d_acc.ensure(256) || !ctx->h_stage.ensure(256)) return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    u64 *d_acc = (u64 *)ctx->d_acc.p;   // [0..7] in, [8..15] out, [16] digest
    u64 *h = (u64 *)ctx->h_stage.p;
    if (acc_in) {
        memcpy(h, acc_in, 64);
        CK(ctx, cudaMemcpyAsync(d_acc, h, 64, cudaMemcpyHostToDevice, s));
    }
    CK(ctx, cudaEventRecord(ctx->ev0, s));
    xxh3_chain_kernel<<<1, 32, 0, s>>>((const u64 *)ctx->d_partials.p, nscr, acc_in ? d_acc : nullptr, d_acc + 8,
                                       final ? 1 : 0, d_out, tail_pos, total_size, d_acc + 16);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaEventRecord(ctx->ev1, s));
    ctx->launches += 1;
    CK(ctx, cudaMemcpyAsync(h + 8, d_acc + 8, 72, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    CK(ctx, cudaEventElapsedTime(&ctx->chain_ms, ctx->ev0, ctx->ev1));
    if (acc_out) memcpy(acc_out, h + 8, 64);
    if (final) *digest = h[16];
    return ZPB_OK;
}

extern "C" int zpb_last_chain_ms(const zpb_ctx *ctx, float *ms) {
    if (!ctx || !ms) return ZPB_E_ARG;
    *ms = ctx->chain_ms;
    return ZPB_OK;
}

// ---- host buffers in, host buffers out: what zpack_read_file does for one large block-independent entry.
// The entry is cut into chunks of ~host_chunk_bytes decoded bytes; chunk c goes to worker c % W (a private
// sub-context: own stream and scratch), so that chunk c+1's H2D, chunk c's kernels and chunk c-1's D2H overlap.
// The workers are this call's "ranks": each decodes its run of blocks, waits for the previous chunk's 64-byte
// accumulator state, runs its part of the XXH3 chain and publishes the state — the same relay N GPUs do.
#include <condition_variable>
#include <mutex>

extern "C" int zpb_unpack_entry_blocks_host(zpb_ctx *ctx, const uint8_t *h_entry, uint64_t comp_size, uint8_t *h_out,
                                            uint64_t out_cap, uint64_t uncomp_size, uint64_t expect_hash,
                                            uint32_t flags, int32_t *status, uint64_t *digest) {
    if (!ctx || !h_entry || (!h_out && out_cap) || !status) return fail(ctx, ZPB_E_ARG, "null argument");
    u64 nb = 0, content = 0;
    u32 bs = 0;
    int rc = zpb_lz4_frame_index(h_entry, comp_size, 0, nullptr, 0, &nb, &bs, &content);
    if (rc != ZPB_OK) return rc;
    if (nb == 0 || uncomp_size > nb * (u64)bs || uncomp_size <= (nb - 1) * (u64)bs ||
        (content != ~0ull && content != uncomp_size))
        return ZPB_INDEX_UNSUPPORTED;                                  // sizes that do not add up: the general path decides
    if (out_cap < uncomp_size) { *status = ZPB_ST_BUFFER_TOO_SMALL; if (digest) *digest = 0; return ZPB_OK; }   // zpack_read.c:329
    std::vector<zpb_block> blocks(nb);
    rc = zpb_lz4_frame_index(h_entry, comp_size, 0, blocks.data(), nb, &nb, &bs, nullptr);
    if (rc != ZPB_OK) return rc;

    const u64 per_chunk = std::max<u64>(2, ctx->host_chunk_bytes / bs);
    std::vector<u64> cuts;
    for (u64 b = 0; b < nb; b += per_chunk) cuts.push_back(b);
    // the last chunk holds the XXH3 tail: never let it be a lone short block
    if (cuts.size() > 1 && nb - cuts.back() < 2) cuts.pop_back();
    cuts.push_back(nb);
    const size_t nchunks = cuts.size() - 1;
    const int W = (int)std::min<size_t>((size_t)std::max(1, ctx->host_workers), nchunks);
    while ((int)ctx->workers.size() < W) {
        zpb_ctx *w = zpb_create(ctx->device);
        if (!w) return fail(ctx, ZPB_E_CUDA, "pipeline sub-context creation failed");
        ctx->workers.push_back(w);
    }
    std::mutex mu;
    std::condition_variable cv;
    size_t chained = 0;             // chunks [0, chained) have been folded into `acc`
    u64 acc[8];
    bool have_acc = false, abort_all = false;
    int32_t verdict = ZPB_ST_OK;
    u64 dg = 0;
    std::vector<int> rcs(W, ZPB_OK);
    std::vector<std::string> errs(W);

    auto body = [&](int t) {
        zpb_ctx *w = ctx->workers[t];
        auto bail = [&](int code, int32_t st) {
            std::lock_guard<std::mutex> lk(mu);
            if (code != ZPB_OK) { rcs[t] = code; errs[t] = w->err; }
            if (st != ZPB_ST_OK && verdict == ZPB_ST_OK) verdict = st;
            abort_all = true;
            cv.notify_all();
        };
        if (cudaSetDevice(w->device) != cudaSuccess) { w->err = "cudaSetDevice failed"; bail(ZPB_E_CUDA, 0); return; }
        cudaStream_t s = w->stream;
        for (size_t c = t; c < nchunks; c += W) {
            { std::lock_guard<std::mutex> lk(mu); if (abort_all) return; }
            const u64 b0 = cuts[c], b1 = cuts[c + 1];
            const u64 pos = b0 * bs, size = std::min<u64>(uncomp_size, b1 * (u64)bs) - pos;
            const u64 lo = blocks[b0].src_off & ~15ull, hi = blocks[b1 - 1].src_off + blocks[b1 - 1].comp_size;
            if (!w->d_in.ensure(hi - lo + 64) || !w->d_out.ensure(size + 64)) { w->err = "device staging allocation failed"; bail(ZPB_E_NOMEM, 0); return; }
            if (cudaMemcpyAsync(w->d_in.p, h_entry + lo, hi - lo, cudaMemcpyHostToDevice, s) != cudaSuccess) { w->err = "H2D failed"; bail(ZPB_E_CUDA, 0); return; }
            std::vector<zpb_block> rel(blocks.begin() + b0, blocks.begin() + b1);
            for (auto &b : rel) b.src_off -= lo;
            int32_t st = 0;
            int r = zpb_unpack_blocks_device(w, (const u8 *)w->d_in.p, hi - lo, (u8 *)w->d_out.p, size, rel.data(), b1 - b0, bs, size, &st, s);
            if (r != ZPB_OK || st != ZPB_ST_OK) { bail(r, st); return; }
            if (!(flags & ZPB_F_DISCARD) &&
                cudaMemcpyAsync(h_out + pos, w->d_out.p, size, cudaMemcpyDeviceToHost, s) != cudaSuccess) { w->err = "D2H failed"; bail(ZPB_E_CUDA, 0); return; }
            u64 in[8], out[8], d = 0;
            bool first;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return chained == c || abort_all; });
                if (abort_all) return;
                first = !have_acc;
                memcpy(in, acc, sizeof in);
            }
            r = zpb_blocks_digest(w, first ? nullptr : in, out, pos, uncomp_size, (const u8 *)w->d_out.p, &d, s);   // syncs s: D2H done too
            if (r != ZPB_OK) { bail(r, 0); return; }
            {
                std::lock_guard<std::mutex> lk(mu);
                memcpy(acc, out, sizeof acc);
                have_acc = true;
                chained = c + 1;
                if (c + 1 == nchunks) dg = d;
                cv.notify_all();
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < W; ++t) th.emplace_back(body, t);
    body(0);
    for (auto &x : th) x.join();
    u64 launches = 0;
    for (int t = 0; t < W; ++t) { launches += ctx->workers[t]->launches; ctx->workers[t]->launches = 0; }
    ctx->launches += launches;
    for (int t = 0; t < W; ++t)
        if (rcs[t] != ZPB_OK) return fail(ctx, rcs[t], errs[t].c_str());
    if (verdict == ZPB_ST_NOT_AVAILABLE) return ZPB_INDEX_UNSUPPORTED;      // a block the fast kernels declined: general path
    if (verdict == ZPB_ST_OK && !(flags & ZPB_F_NO_VERIFY) && dg != expect_hash) verdict = ZPB_ST_HASH_MISMATCH;   // zpack_read.c:466-468
    *status = verdict;
    if (digest) *digest = verdict == ZPB_ST_OK || verdict == ZPB_ST_HASH_MISMATCH ? dg : 0;
    return ZPB_OK;
}
