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
    return 0;
}

static int pack_device_impl(zpb_ctx *ctx, const u8 *d_in, u64 in_size, u8 *d_out, u64 out_size,
                            const zpb_file *files, u64 n, u64 *comp_size, u64 *digest, int32_t *status,
                            cudaStream_t s) {
    if (n == 0) return ZPB_OK;
    if (n > 0x7fffffffull) return fail(ctx, ZPB_E_ARG, "too many files in one batch");
    size_t desc_b = n * sizeof(zpb_file), ord_b = n * sizeof(u32);
    size_t res_b = n * (2 * sizeof(u64) + sizeof(int));
    if (!ctx->d_desc.ensure(desc_b) || !ctx->d_order.ensure(ord_b) || !ctx->d_res.ensure(res_b + 64) ||
        !ctx->d_counter.ensure(256) || !ctx->h_stage.ensure(desc_b + ord_b + res_b + 64))
        return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    u8 *hs = (u8 *)ctx->h_stage.p;
    memcpy(hs, files, desc_b);
    u32 *h_order = (u32 *)(hs + desc_b);
    {   // largest files first (counting sort on size / 4 KiB, descending)
        auto key = [&](u64 i) -> u32 { u64 c = files[i].size >> 12; return c > 1023 ? 0u : 1023u - (u32)c; };
        std::vector<u32> head(1025, 0);
        for (u64 i = 0; i < n; ++i) ++head[key(i) + 1];
        for (u32 k = 0; k < 1024; ++k) head[k + 1] += head[k];
        for (u64 i = 0; i < n; ++i) h_order[head[key(i)]++] = (u32)i;
    }
    u64 *d_comp = (u64 *)ctx->d_res.p;
    u64 *d_dig = d_comp + n;
    int *d_st = (int *)(d_dig + n);
    CK(ctx, cudaMemcpyAsync(ctx->d_desc.p, hs, desc_b, cudaMemcpyHostToDevice, s));
    CK(ctx, cudaMemcpyAsync(ctx->d_order.p, h_order, ord_b, cudaMemcpyHostToDevice, s));
    CK(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, 256, s));
    CK(ctx, cudaEventRecord(ctx->ev0, s));
    int per_sm = 0;
    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lz4_pack_kernel, 32 * PK_WARPS, 0));
    if (per_sm < 1) per_sm = 1;
    u64 want = (n + PK_WARPS - 1) / PK_WARPS;
    u32 grid = (u32)std::min<u64>((u64)ctx->sm_count * per_sm, std::max<u64>(want, 1));
    lz4_pack_kernel<<<grid, 32 * PK_WARPS, 0, s>>>(d_in, in_size, d_out, out_size, (const zpb_file *)ctx->d_desc.p,
                                                   (const u32 *)ctx->d_order.p, (u32)n, (u32 *)ctx->d_counter.p,
                                                   d_comp, d_dig, d_st);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    CK(ctx, cudaEventRecord(ctx->ev1, s));
    u8 *h_res = hs + desc_b + ord_b;
    CK(ctx, cudaMemcpyAsync(h_res, ctx->d_res.p, res_b, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    CK(ctx, cudaEventElapsedTime(&ctx->pack_ms, ctx->ev0, ctx->ev1));
    if (comp_size) memcpy(comp_size, h_res, n * sizeof(u64));
    if (digest) memcpy(digest, h_res + n * sizeof(u64), n * sizeof(u64));
    if (status) memcpy(status, h_res + 2 * n * sizeof(u64), n * sizeof(int));
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

extern "C" int zpb_pack_host(zpb_ctx *ctx, const uint8_t *h_in, uint64_t in_size, uint8_t *h_out,
                             uint64_t out_size, const zpb_file *files, uint64_t n, uint64_t *comp_size,
                             uint64_t *digest, int32_t *status) {
    if (!ctx || (!files && n) || (!h_in && in_size) || (!h_out && out_size) || !comp_size)
        return fail(ctx, ZPB_E_ARG, "null argument");
    if (n == 0) return ZPB_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (!ctx->d_in.ensure(in_size + 64) || !ctx->d_out.ensure(out_size + 64))
        return fail(ctx, ZPB_E_NOMEM, "device staging allocation failed");
    if (in_size) CK(ctx, cudaMemcpyAsync(ctx->d_in.p, h_in, in_size, cudaMemcpyHostToDevice, s));
    int rc = pack_device_impl(ctx, (const u8 *)ctx->d_in.p, in_size, (u8 *)ctx->d_out.p, out_size, files, n,
                              comp_size, digest, status, s);
    if (rc != ZPB_OK) return rc;
    // only the bytes each file actually produced travel back
    for (u64 i = 0; i < n; ++i) {
        if (!comp_size[i] || (status && status[i] != ZPB_ST_OK)) continue;
        if (files[i].dst_off > out_size || comp_size[i] > out_size - files[i].dst_off) continue;
        CK(ctx, cudaMemcpyAsync(h_out + files[i].dst_off, (u8 *)ctx->d_out.p + files[i].dst_off, comp_size[i],
                                cudaMemcpyDeviceToHost, s));
    }
    CK(ctx, cudaStreamSynchronize(s));
    return ZPB_OK;
}
