// pack_api.inl — host side of zpb_pack_* (included by zpb_api.cu).
//
// zpack_write_files (/root/reference/lib/zpack_write.c:280-343) compresses file by file and appends;
// here every file of the batch is compressed by one kernel launch into its own bounded slot, and the
// caller (host) does the exclusive prefix sum of comp_size[] that yields entry.offset, exactly the
// division of labour the north-star asks for.
extern "C" uint64_t zpb_pack_bound(uint32_t method, uint64_t size) {
    // LZ4: frame header 7 + per 64 KB block (4-byte header + stored payload worst case) + EndMark 4
    // (same role as LZ4F_compressBound at lib/zpack_write.c:141; an empty file is 11 bytes, as in the reference)
    if (method == ZPB_METHOD_LZ4) return 7 + 4 * ((size + 65535) / 65536) + size + 4;
    if (method == ZPB_METHOD_NONE) return size;
    // zstd: frame header 14 + per 64 KB block a 3-byte header + at most the raw payload
    if (method == ZPB_METHOD_ZSTD) return 14 + 3 * (size ? (size + 65535) / 65536 : 1) + size;
    return 0;
}

// which files have blocks in the stage-1 list: LZ4 below the HC levels and zstd, in bounds, with a slot that holds the worst case
static bool pack_file_has_blocks(const zpb_file &f, u64 in_size, u64 out_size) {
    if (f.src_off > in_size || f.size > in_size - f.src_off || f.dst_off > out_size || f.dst_cap > out_size - f.dst_off) return false;
    const u64 fb = (f.size + 65535) >> 16;
    if (f.method == ZPB_METHOD_LZ4) return f.level < 3 && f.dst_cap >= 7 + 4 * fb + f.size + 4;
    if (f.method == ZPB_METHOD_ZSTD) return f.dst_cap >= 14 + 3 * (f.size ? fb : 1) + f.size;
    return false;
}

// Two launches per round: lz4_pack_blocks_kernel compresses every 64 KB block of the round's LZ4 files (one warp per
// block, payload into a scratch slot per block), lz4_pack_kernel lays the frames out and hashes the inputs (one warp per
// file).  A round takes as many files (largest first) as its blocks fit the scratch: pack_scratch_blocks slots of 64 KB.
static int pack_device_impl(zpb_ctx *ctx, const u8 *d_in, u64 in_size, u8 *d_out, u64 out_size,
                            const zpb_file *files, u64 n, u64 *comp_size, u64 *digest, int32_t *status,
                            cudaStream_t s) {
    if (n == 0) return ZPB_OK;
    if (n > 0x7fffffffull) return fail(ctx, ZPB_E_ARG, "too many files in one batch");
    // block list of the LZ4 files the stage-2 kernel will accept (same checks as there)
    u64 nblk = 0;
    bool any_zstd = false;
    for (u64 i = 0; i < n; ++i) {
        const zpb_file &f = files[i];
        if (!pack_file_has_blocks(f, in_size, out_size)) continue;
        nblk += (f.size + 65535) >> 16;
        any_zstd |= f.method == ZPB_METHOD_ZSTD;
    }
    if (nblk > 0x7fffffffull) return fail(ctx, ZPB_E_ARG, "too many blocks in one batch");
    const size_t desc_b = n * sizeof(zpb_file), ord_b = (n * sizeof(u32) + 15) & ~(size_t)15;
    const size_t res_b = n * (2 * sizeof(u64) + sizeof(int));
    const size_t blk_b = nblk * sizeof(PackBlock), base_b = ord_b;
    // zstd blocks need two more slots each (the block body and 128 KB of sequence records): smaller rounds
    const u64 cap_blocks = std::max<u64>(any_zstd ? std::min<u64>(ctx->pack_scratch_blocks, 4096) : ctx->pack_scratch_blocks, 1);
    if (!ctx->d_desc.ensure(desc_b) || !ctx->d_order.ensure(ord_b + base_b) || !ctx->d_res.ensure(res_b + 64) ||
        !ctx->d_counter.ensure(256) || !ctx->h_stage.ensure(desc_b + ord_b + base_b + blk_b + res_b + 64) ||
        !ctx->d_pblk.ensure(blk_b + 16) || !ctx->d_csize.ensure(std::min(nblk, cap_blocks) * sizeof(u32) + 16))
        return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    u8 *hs = (u8 *)ctx->h_stage.p;
    memcpy(hs, files, desc_b);
    u32 *h_order = (u32 *)(hs + desc_b);
    u32 *h_base = (u32 *)(hs + desc_b + ord_b);
    PackBlock *h_blk = (PackBlock *)(hs + desc_b + ord_b + base_b);
    {   // largest files first (counting sort on size / 4 KiB, descending)
        auto key = [&](u64 i) -> u32 { u64 c = files[i].size >> 12; return c > 1023 ? 0u : 1023u - (u32)c; };
        std::vector<u32> head(1025, 0);
        for (u64 i = 0; i < n; ++i) ++head[key(i) + 1];
        for (u32 k = 0; k < 1024; ++k) head[k + 1] += head[k];
        for (u64 i = 0; i < n; ++i) h_order[head[key(i)]++] = (u32)i;
    }
    // rounds: [first file in h_order, first block in h_blk); block indices in h_base are relative to the round's scratch
    struct Round { u64 f0, f1, b0, b1; };
    std::vector<Round> rounds;
    {
        u64 bi = 0, rb0 = 0, rf0 = 0;
        for (u64 k = 0; k < n; ++k) {
            const u32 i = h_order[k];
            const zpb_file &f = files[i];
            h_base[i] = 0;
            if (!pack_file_has_blocks(f, in_size, out_size)) continue;
            const u64 fb = (f.size + 65535) >> 16;
            if (bi - rb0 + fb > cap_blocks && bi > rb0) { rounds.push_back({rf0, k, rb0, bi}); rf0 = k; rb0 = bi; }
            h_base[i] = (u32)(bi - rb0);
            for (u64 b = 0; b < fb; ++b) {
                h_blk[bi].src_off = f.src_off + (b << 16);
                h_blk[bi].len = (u32)std::min<u64>(f.size - (b << 16), 65536);
                h_blk[bi].pad = f.method == ZPB_METHOD_ZSTD ? 1u : 0u;
                ++bi;
            }
        }
        rounds.push_back({rf0, n, rb0, bi});
    }
    u64 max_round = 0;
    for (const Round &r : rounds) max_round = std::max(max_round, r.b1 - r.b0);
    if (!ctx->d_pscratch.ensure((max_round << 16) + 64) || !ctx->d_csize.ensure(max_round * sizeof(u32) + 16))
        return fail(ctx, ZPB_E_NOMEM, "block scratch allocation failed");
    if (any_zstd && (!ctx->d_zslot.ensure(max_round * (u64)ZE_SLOT + 64) || !ctx->d_zseq.ensure(max_round * 16 * (u64)ZE_WIN_SEQ * sizeof(u64) + 64) ||
                     !ctx->d_zmeta.ensure(max_round * (17 + ZE_META) * sizeof(u32) + 64) || !ctx->d_zelit.ensure(max_round * (u64)ZE_LITSLOT + 64) ||
                     !ctx->d_zhuf.ensure(max_round * sizeof(ZeHuf) + 64) || !ctx->d_ztabs.ensure(max_round * sizeof(ZeBlockTabs) + 64)))
        return fail(ctx, ZPB_E_NOMEM, "zstd block scratch allocation failed");
    u32 *d_winop = any_zstd ? (u32 *)ctx->d_zmeta.p : nullptr;
    u32 *d_zbody = any_zstd ? d_winop + max_round * 17 : nullptr;      // ZE_META words per block, the body sizes first
    u64 *d_comp = (u64 *)ctx->d_res.p;
    u64 *d_dig = d_comp + n;
    int *d_st = (int *)(d_dig + n);
    const u32 *d_ord = (const u32 *)ctx->d_order.p;
    const u32 *d_base = (const u32 *)((const u8 *)ctx->d_order.p + ord_b);
    CK(ctx, cudaMemcpyAsync(ctx->d_desc.p, hs, desc_b, cudaMemcpyHostToDevice, s));
    CK(ctx, cudaMemcpyAsync(ctx->d_order.p, h_order, ord_b + base_b, cudaMemcpyHostToDevice, s));
    if (blk_b) CK(ctx, cudaMemcpyAsync(ctx->d_pblk.p, h_blk, blk_b, cudaMemcpyHostToDevice, s));
    CK(ctx, cudaEventRecord(ctx->ev0, s));
    while (ctx->pack_evs.size() < 3 * rounds.size()) {
        cudaEvent_t e = nullptr;
        CK(ctx, cudaEventCreate(&e));
        ctx->pack_evs.push_back(e);
    }
    size_t ri = 0;
    for (const Round &r : rounds) {
        CK(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, 256, s));
        CK(ctx, cudaEventRecord(ctx->pack_evs[3 * ri], s));
        const u64 rb = r.b1 - r.b0, rf = r.f1 - r.f0;
        if (rb) {
            const u32 grid = (u32)std::min<u64>((u64)ctx->sm_count * ctx->p2_per_sm, (rb + P2_WARPS - 1) / P2_WARPS);
            lz4_pack_blocks_kernel<<<grid, 32 * P2_WARPS, P2_SMEM, s>>>(d_in, (const PackBlock *)ctx->d_pblk.p + r.b0, (u32)rb,
                                                                        (u32 *)ctx->d_counter.p + 32, (u8 *)ctx->d_pscratch.p,
                                                                        (u32 *)ctx->d_csize.p, d_winop);
            CK(ctx, cudaGetLastError());
            ctx->launches += 1;
        }
        if (rb && any_zstd) {
            const u32 zgrid = (u32)std::min<u64>((rb * 16 + 127) / 128, (u64)ctx->sm_count * 16);
            const PackBlock *zb = (const PackBlock *)ctx->d_pblk.p + r.b0;
            zstd_parse_windows_kernel<<<zgrid, 128, 0, s>>>((const u8 *)ctx->d_pscratch.p, (const u32 *)ctx->d_csize.p, zb, d_winop, (u32)rb,
                                                            d_zbody, (u8 *)ctx->d_zelit.p, (u64 *)ctx->d_zseq.p);
            zstd_huf_tables_kernel<<<(u32)std::min<u64>((rb + 3) / 4, (u64)ctx->sm_count * 8), 128, 0, s>>>(
                zb, (const u32 *)ctx->d_csize.p, d_winop, (u32)rb, d_zbody, (const u8 *)ctx->d_zelit.p, (const u64 *)ctx->d_zseq.p, (ZeHuf *)ctx->d_zhuf.p,
                (ZeBlockTabs *)ctx->d_ztabs.p);
            zstd_encode_windows_kernel<<<zgrid, 128, 0, s>>>((const u32 *)ctx->d_csize.p, zb, d_winop, (u32)rb, d_zbody, (const u8 *)ctx->d_zelit.p,
                                                             (const u64 *)ctx->d_zseq.p, (const ZeHuf *)ctx->d_zhuf.p, (const ZeBlockTabs *)ctx->d_ztabs.p,
                                                             (u8 *)ctx->d_zslot.p);
            ctx->launches += 2;
            CK(ctx, cudaGetLastError());
            ctx->launches += 1;
        }
        CK(ctx, cudaEventRecord(ctx->pack_evs[3 * ri + 1], s));
        if (rf) {
            const u32 grid = (u32)std::min<u64>((u64)ctx->sm_count * ctx->pk_per_sm, (rf + PK_WARPS - 1) / PK_WARPS);
            lz4_pack_kernel<<<grid, 32 * PK_WARPS, 0, s>>>(d_in, in_size, d_out, out_size, (const zpb_file *)ctx->d_desc.p,
                                                           d_ord + r.f0, (u32)rf, (u32 *)ctx->d_counter.p, d_comp, d_dig, d_st,
                                                           d_base, (const u8 *)ctx->d_pscratch.p, (const u32 *)ctx->d_csize.p,
                                                           (const u8 *)ctx->d_zslot.p, d_zbody, d_winop);
            CK(ctx, cudaGetLastError());
            ctx->launches += 1;
        }
        CK(ctx, cudaEventRecord(ctx->pack_evs[3 * ri + 2], s));
        ++ri;
    }
    CK(ctx, cudaEventRecord(ctx->ev1, s));
    u8 *h_res = hs + desc_b + ord_b + base_b + blk_b;
    CK(ctx, cudaMemcpyAsync(h_res, ctx->d_res.p, res_b, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    CK(ctx, cudaEventElapsedTime(&ctx->pack_ms, ctx->ev0, ctx->ev1));
    ctx->pack_blocks_ms = ctx->pack_frames_ms = 0.f;
    for (size_t k = 0; k < rounds.size(); ++k) {
        float a = 0.f, b = 0.f;
        CK(ctx, cudaEventElapsedTime(&a, ctx->pack_evs[3 * k], ctx->pack_evs[3 * k + 1]));
        CK(ctx, cudaEventElapsedTime(&b, ctx->pack_evs[3 * k + 1], ctx->pack_evs[3 * k + 2]));
        ctx->pack_blocks_ms += a;
        ctx->pack_frames_ms += b;
    }
    if (comp_size) memcpy(comp_size, h_res, n * sizeof(u64));
    if (digest) memcpy(digest, h_res + n * sizeof(u64), n * sizeof(u64));
    if (status) memcpy(status, h_res + 2 * n * sizeof(u64), n * sizeof(int));
    return ZPB_OK;
}

extern "C" int zpb_last_pack_stage_ms(const zpb_ctx *ctx, float *blocks_ms, float *frames_ms) {
    if (!ctx) return ZPB_E_ARG;
    if (blocks_ms) *blocks_ms = ctx->pack_blocks_ms;
    if (frames_ms) *frames_ms = ctx->pack_frames_ms;
    return ZPB_OK;
}

extern "C" int zpb_pack_device(zpb_ctx *ctx, const uint8_t *d_in, uint64_t in_size, uint8_t *d_out,
                               uint64_t out_size, const zpb_file *files, uint64_t n, uint64_t *comp_size,
                               uint64_t *digest, int32_t *status, void *stream) {
    if (!ctx || (!files && n) || (!d_in && in_size) || (!d_out && out_size))
        return fail(ctx, ZPB_E_ARG, "null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    return pack_device_impl(ctx, d_in, in_size, d_out, out_size, files, n, comp_size, digest, status, s);
}

// Gathers every file's frame from its bounded slot into one contiguous staging buffer (16-byte aligned starts):
// a single D2H then carries exactly the compressed bytes instead of one small copy per file.
__global__ void __launch_bounds__(256)
pack_gather_kernel(const u8 *__restrict__ slots, u8 *__restrict__ compact, const zpb_file *__restrict__ files,
                   const u64 *__restrict__ comp, const u64 *__restrict__ off, u32 n) {
    const u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const u8 *src = slots + files[w].dst_off;
    u8 *dst = compact + off[w];
    const u64 len = comp[w];
    if ((((uintptr_t)src) & 15u) == 0) {
        const u64 n16 = len >> 4;
        for (u64 i = lane; i < n16; i += 32) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
        for (u64 i = (n16 << 4) + lane; i < len; i += 32) dst[i] = src[i];
    } else {
        for (u64 i = lane; i < len; i += 32) dst[i] = src[i];
    }
}

// One chunk on one (sub-)context: H2D of the chunk's input range -> pack kernel -> gather -> one D2H into a pinned
// bounce buffer -> host scatter into the caller's slots.
static int pack_host_chunk(zpb_ctx *ctx, const uint8_t *h_in, uint64_t in_size, uint8_t *h_out, uint64_t out_size,
                           const zpb_file *files, uint64_t n, uint64_t *comp_size, uint64_t *digest, int32_t *status) {
    if (n == 0) return ZPB_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    u64 lo = ~0ull, hi = 0, olo = ~0ull, ohi = 0;
    for (u64 i = 0; i < n; ++i) {
        const zpb_file &f = files[i];
        if (f.src_off > in_size || f.size > in_size - f.src_off || f.dst_off > out_size || f.dst_cap > out_size - f.dst_off)
            continue;                                             // flagged by the kernel's own bounds check
        if (f.size) { lo = std::min(lo, f.src_off); hi = std::max(hi, f.src_off + f.size); }
        olo = std::min(olo, f.dst_off); ohi = std::max(ohi, f.dst_off + f.dst_cap);
    }
    if (hi <= lo) lo = hi = 0;
    if (ohi <= olo) olo = ohi = 0;
    lo &= ~15ull; olo &= ~15ull;                                  // device layout congruent to the host layout mod 16
    std::vector<zpb_file> rel(files, files + n);
    for (auto &f : rel) {
        if (f.src_off > in_size || f.size > in_size - f.src_off || f.dst_off > out_size || f.dst_cap > out_size - f.dst_off) {
            f.src_off = ~0ull;                                    // stays invalid after rebasing
            continue;
        }
        f.src_off -= f.size ? lo : f.src_off;
        f.dst_off -= olo;
    }
    if (!ctx->d_in.ensure(hi - lo + 64) || !ctx->d_out.ensure(ohi - olo + 64))
        return fail(ctx, ZPB_E_NOMEM, "device staging allocation failed");
    if (hi > lo) CK(ctx, cudaMemcpyAsync(ctx->d_in.p, h_in + lo, hi - lo, cudaMemcpyHostToDevice, s));
    std::vector<int32_t> st_local;
    if (!status) { st_local.resize(n); status = st_local.data(); }
    int rc = pack_device_impl(ctx, (const u8 *)ctx->d_in.p, hi - lo, (u8 *)ctx->d_out.p, ohi - olo, rel.data(), n,
                              comp_size, digest, status, s);
    if (rc != ZPB_OK) return rc;
    // compact layout of what each file actually produced
    std::vector<u64> off(n);
    u64 total = 0;
    for (u64 i = 0; i < n; ++i) {
        const bool ok = status[i] == ZPB_ST_OK && comp_size[i] && rel[i].src_off != ~0ull && comp_size[i] <= files[i].dst_cap;
        off[i] = total;
        if (!ok) { off[i] = ~0ull; continue; }
        total += (comp_size[i] + 15) & ~15ull;
    }
    if (total == 0) return ZPB_OK;
    if (!ctx->d_gather.ensure(total + 64) || !ctx->d_goff.ensure(2 * n * sizeof(u64)) || !ctx->h_bounce.ensure(total + 64))
        return fail(ctx, ZPB_E_NOMEM, "gather staging allocation failed");
    std::vector<u64> up(2 * n);
    for (u64 i = 0; i < n; ++i) { up[i] = off[i] == ~0ull ? 0 : comp_size[i]; up[n + i] = off[i] == ~0ull ? 0 : off[i]; }
    CK(ctx, cudaMemcpyAsync(ctx->d_goff.p, up.data(), 2 * n * sizeof(u64), cudaMemcpyHostToDevice, s));
    pack_gather_kernel<<<(u32)((n * 32 + 255) / 256), 256, 0, s>>>((const u8 *)ctx->d_out.p, (u8 *)ctx->d_gather.p,
                                                                  (const zpb_file *)ctx->d_desc.p, (const u64 *)ctx->d_goff.p,
                                                                  (const u64 *)ctx->d_goff.p + n, (u32)n);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    CK(ctx, cudaMemcpyAsync(ctx->h_bounce.p, ctx->d_gather.p, total, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    const u8 *b = (const u8 *)ctx->h_bounce.p;
    for (u64 i = 0; i < n; ++i)
        if (off[i] != ~0ull) memcpy(h_out + files[i].dst_off, b + off[i], comp_size[i]);
    return ZPB_OK;
}

// Host buffers in, host buffers out.  Large batches are cut into chunks of ~host_chunk_bytes of input (files in
// source order, so a chunk's input is one contiguous H2D) and dealt to `host_workers` threads, each driving a private
// sub-context: while one chunk is being packed, the next one's input is on its way in and the previous one's frames
// on their way out.
extern "C" int zpb_pack_host(zpb_ctx *ctx, const uint8_t *h_in, uint64_t in_size, uint8_t *h_out,
                             uint64_t out_size, const zpb_file *files, uint64_t n, uint64_t *comp_size,
                             uint64_t *digest, int32_t *status) {
    if (!ctx || (!files && n) || (!h_in && in_size) || (!h_out && out_size) || !comp_size)
        return fail(ctx, ZPB_E_ARG, "null argument");
    if (n == 0) return ZPB_OK;
    u64 total = 0;
    for (u64 i = 0; i < n; ++i) total += files[i].size;
    const int W = ctx->host_workers;
    if (W <= 1 || total < 2 * ctx->host_chunk_bytes)
        return pack_host_chunk(ctx, h_in, in_size, h_out, out_size, files, n, comp_size, digest, status);

    std::vector<u32> by_src(n);
    std::iota(by_src.begin(), by_src.end(), 0u);
    std::sort(by_src.begin(), by_src.end(), [&](u32 a, u32 b) { return files[a].src_off < files[b].src_off; });
    std::vector<u64> cuts{0};
    u64 acc = 0;
    for (u64 k = 0; k < n; ++k) {
        acc += files[by_src[k]].size;
        if (acc >= ctx->host_chunk_bytes && k + 1 < n) { cuts.push_back(k + 1); acc = 0; }
    }
    cuts.push_back(n);
    const size_t nchunks = cuts.size() - 1;
    while ((int)ctx->workers.size() < W) {
        zpb_ctx *w = zpb_create(ctx->device);
        if (!w) return fail(ctx, ZPB_E_CUDA, "pipeline sub-context creation failed");
        ctx->workers.push_back(w);
    }
    std::vector<int> rcs(W, ZPB_OK);
    std::vector<std::string> errs(W);
    auto body = [&](int t) {
        zpb_ctx *w = ctx->workers[t];
        std::vector<zpb_file> sub;
        std::vector<int32_t> st;
        std::vector<u64> cs, dg;
        for (size_t c = t; c < nchunks; c += W) {
            const u64 a = cuts[c], b = cuts[c + 1], m = b - a;
            sub.resize(m); st.resize(m); cs.resize(m); dg.resize(m);
            for (u64 k = 0; k < m; ++k) sub[k] = files[by_src[a + k]];
            int rc = pack_host_chunk(w, h_in, in_size, h_out, out_size, sub.data(), m, cs.data(), dg.data(), st.data());
            if (rc != ZPB_OK) { rcs[t] = rc; errs[t] = w->err; return; }
            for (u64 k = 0; k < m; ++k) {
                comp_size[by_src[a + k]] = cs[k];
                if (digest) digest[by_src[a + k]] = dg[k];
                if (status) status[by_src[a + k]] = st[k];
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < W; ++t) th.emplace_back(body, t);
    body(0);
    for (auto &x : th) x.join();
    u64 launches = 0;
    float ms = 0.f;
    for (zpb_ctx *w : ctx->workers) { launches += w->launches; w->launches = 0; ms += w->pack_ms; }
    ctx->launches += launches;
    ctx->pack_ms = ms;
    for (int t = 0; t < W; ++t)
        if (rcs[t] != ZPB_OK) return fail(ctx, rcs[t], errs[t].c_str());
    return ZPB_OK;
}
