// archive_api.inl — host side of the device-resident container operations (part of zpb_api.cu): argument checks and the
// two or three scalars the host needs (totals, the 42 fixed bytes of an archive), launches, CUDA-event timing.  The
// per-entry work — offset table, record positions, byte moves, record parsing — runs in archive_kernels.cuh.

static int arc_upload_entries(zpb_ctx *ctx, const zpb_arc_entry *entries, u64 n, cudaStream_t s) {
    const size_t b = (size_t)n * sizeof(ArcEntry);
    if (!ctx->d_arc_e.ensure(b + 64) || !ctx->d_arc_rec.ensure((n + 1) * 8) || !ctx->d_arc_chunk.ensure((n + 2) * 8) ||
        !ctx->d_arc_tot.ensure(64) || !ctx->h_stage.ensure(b + 64))
        return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    if (n) {
        memcpy(ctx->h_stage.p, entries, b);
        CK(ctx, cudaMemcpyAsync(ctx->d_arc_e.p, ctx->h_stage.p, b, cudaMemcpyHostToDevice, s));
    }
    return ZPB_OK;
}

// the offset table: one CTA per tile of 4 x ARC_SCAN_THREADS entries (archive_kernels.cuh)
static int arc_launch_layout(zpb_ctx *ctx, u64 n, u64 base, u32 assign, cudaStream_t s) {
    const u64 ntiles = std::max<u64>(1, (n + 4 * ARC_SCAN_THREADS - 1) / (4 * ARC_SCAN_THREADS));
    if (!ctx->d_arc_scan.ensure(ntiles * 32 + 64)) return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    CK(ctx, cudaMemsetAsync(ctx->d_arc_scan.p, 0, ntiles * 32, s));
    CK(ctx, cudaMemsetAsync(ctx->d_arc_tot.p, 0, 64, s));
    arc_layout_kernel<<<(u32)ntiles, ARC_SCAN_THREADS, 37 * 24, s>>>((ArcEntry *)ctx->d_arc_e.p, n, base, assign, (u64 *)ctx->d_arc_rec.p,
                                                                    (u64 *)ctx->d_arc_chunk.p, (u64 *)ctx->d_arc_tot.p, (u64 *)ctx->d_arc_scan.p);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ZPB_OK;
}

// the copy's work items (before the event that starts the copy kernel's interval), then the copy itself
static int arc_launch_chunks(zpb_ctx *ctx, u64 n, u64 nchunks, cudaStream_t s) {
    if (!nchunks) return ZPB_OK;
    if (!ctx->d_arc_work.ensure(nchunks * sizeof(ArcChunk) + 64)) return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    arc_chunks_kernel<<<(u32)((nchunks + 255) / 256), 256, 0, s>>>((const ArcEntry *)ctx->d_arc_e.p, n, (const u64 *)ctx->d_arc_chunk.p,
                                                                   nchunks, (ArcChunk *)ctx->d_arc_work.p);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ZPB_OK;
}
static int arc_launch_copy(zpb_ctx *ctx, const u8 *d_src, u8 *d_dst, u64 nchunks, cudaStream_t s) {
    if (!nchunks) return ZPB_OK;
    const u32 grid = (u32)std::min<u64>(nchunks, (u64)ctx->sm_count * ctx->arc_ctas_per_sm);
    arc_copy_kernel<<<grid, ARC_COPY_THREADS, 16, s>>>(d_src, d_dst, (const ArcChunk *)ctx->d_arc_work.p, nchunks,
                                                       ctx->arc_dynamic ? (unsigned long long *)ctx->d_arc_tot.p + 4 : nullptr,
                                                       (u32)ctx->arc_head_mask);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ZPB_OK;
}

static bool arc_range_ok(u64 off, u64 len, u64 size) { return off <= size && len <= size - off; }

extern "C" int zpb_copy_entries_device(zpb_ctx *ctx, const uint8_t *d_src, uint64_t src_size, uint8_t *d_dst, uint64_t dst_size,
                                       const zpb_arc_entry *entries, uint64_t n, void *stream) {
    if (!ctx || (!entries && n)) return fail(ctx, ZPB_E_ARG, "null argument");
    if (n == 0) return ZPB_OK;
    if (!d_src || !d_dst) return fail(ctx, ZPB_E_ARG, "null device buffer");
    if (n > 0x7fffffffull) return fail(ctx, ZPB_E_ARG, "too many entries");
    u64 nchunks = 0;
    for (u64 i = 0; i < n; ++i) {
        if (!arc_range_ok(entries[i].src_off, entries[i].comp_size, src_size)) return fail(ctx, ZPB_E_ARG, "entry outside the source buffer");
        if (!arc_range_ok(entries[i].offset, entries[i].comp_size, dst_size)) return fail(ctx, ZPB_E_ARG, "entry outside the destination buffer");
        nchunks += (entries[i].comp_size + ARC_CHUNK - 1) >> ARC_CHUNK_LOG;
    }
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    int rc = arc_upload_entries(ctx, entries, n, s);
    if (rc) return rc;
    CK(ctx, cudaEventRecord(ctx->evs[0], s));
    if ((rc = arc_launch_layout(ctx, n, 0, 0u, s))) return rc;
    if ((rc = arc_launch_chunks(ctx, n, nchunks, s))) return rc;
    CK(ctx, cudaEventRecord(ctx->evs[1], s));
    if ((rc = arc_launch_copy(ctx, d_src, d_dst, nchunks, s))) return rc;
    CK(ctx, cudaEventRecord(ctx->evs[2], s));
    CK(ctx, cudaStreamSynchronize(s));
    CK(ctx, cudaEventElapsedTime(&ctx->arc_ms[0], ctx->evs[0], ctx->evs[1]));
    CK(ctx, cudaEventElapsedTime(&ctx->arc_ms[1], ctx->evs[1], ctx->evs[2]));
    return ZPB_OK;
}

extern "C" int zpb_archive_build_device(zpb_ctx *ctx, const uint8_t *d_src, uint64_t src_size, zpb_arc_entry *entries, uint64_t n,
                                        const uint8_t *names, uint64_t names_size, uint8_t *d_archive, uint64_t archive_cap,
                                        uint64_t *archive_size, void *stream) {
    if (!ctx || (!entries && n) || !d_archive || !archive_size) return fail(ctx, ZPB_E_ARG, "null argument");
    if (n > 0x7fffffffull) return fail(ctx, ZPB_E_ARG, "too many entries");
    u64 data = 0, block = 0, nchunks = 0;
    for (u64 i = 0; i < n; ++i) {
        const zpb_arc_entry &e = entries[i];
        if (e.name_len > 65535u) return fail(ctx, ZPB_E_ARG, "filename longer than 65535 bytes");     // ZPACK_ERROR_FILENAME_TOO_LONG
        if (!arc_range_ok(e.name_off, e.name_len, names_size) || (e.name_len && !names)) return fail(ctx, ZPB_E_ARG, "name outside the names blob");
        if (!arc_range_ok(e.src_off, e.comp_size, src_size) || (e.comp_size && !d_src)) return fail(ctx, ZPB_E_ARG, "entry outside the source buffer");
        if (data + e.comp_size < data) return fail(ctx, ZPB_E_ARG, "sizes overflow");
        data += e.comp_size;
        block += ARC_FIXED + e.name_len;
        nchunks += (e.comp_size + ARC_CHUNK - 1) >> ARC_CHUNK_LOG;
    }
    const u64 cdr_off = ARC_DATA_START + data, total = cdr_off + ARC_CDR_HDR + block + 12;
    if (total < data || total > archive_cap) return fail(ctx, ZPB_E_ARG, "archive buffer too small");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    int rc = arc_upload_entries(ctx, entries, n, s);
    if (rc) return rc;
    if (!ctx->d_arc_names.ensure(names_size + 16) || !ctx->h_bounce.ensure(names_size + 16)) return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    if (names_size) {
        memcpy(ctx->h_bounce.p, names, names_size);
        CK(ctx, cudaMemcpyAsync(ctx->d_arc_names.p, ctx->h_bounce.p, names_size, cudaMemcpyHostToDevice, s));
    }
    CK(ctx, cudaEventRecord(ctx->evs[0], s));
    if ((rc = arc_launch_layout(ctx, n, ARC_DATA_START, 1u, s))) return rc;
    const u32 cgrid = (u32)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)ctx->sm_count * 8));
    arc_cdr_kernel<<<cgrid, 256, 0, s>>>(d_archive, (const ArcEntry *)ctx->d_arc_e.p, n, (const u8 *)ctx->d_arc_names.p,
                                         (const u64 *)ctx->d_arc_rec.p, cdr_off, block, 1u);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    if ((rc = arc_launch_chunks(ctx, n, nchunks, s))) return rc;
    CK(ctx, cudaEventRecord(ctx->evs[1], s));
    if ((rc = arc_launch_copy(ctx, d_src, d_archive, nchunks, s))) return rc;
    CK(ctx, cudaEventRecord(ctx->evs[2], s));
    if (n) CK(ctx, cudaMemcpyAsync(ctx->h_stage.p, ctx->d_arc_e.p, (size_t)n * sizeof(ArcEntry), cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    if (n) memcpy(entries, ctx->h_stage.p, (size_t)n * sizeof(ArcEntry));
    CK(ctx, cudaEventElapsedTime(&ctx->arc_ms[0], ctx->evs[0], ctx->evs[1]));
    CK(ctx, cudaEventElapsedTime(&ctx->arc_ms[1], ctx->evs[1], ctx->evs[2]));
    *archive_size = total;
    return ZPB_OK;
}

extern "C" int zpb_archive_open_device(zpb_ctx *ctx, const uint8_t *d_archive, uint64_t archive_size, zpb_arc_entry *entries,
                                       uint64_t cap, uint64_t *n_out, uint8_t *names, uint64_t names_cap, uint64_t *names_size,
                                       int32_t *result, void *stream) {
    if (!ctx || !d_archive || !n_out || !result) return fail(ctx, ZPB_E_ARG, "null argument");
    *n_out = 0;
    if (names_size) *names_size = 0;
    // the reference's open, in its order (/root/reference/lib/zpack_read.c:225-260): size, header, data signature, EOCDR, CDR
    if (archive_size < 42) { *result = 5; return ZPB_OK; }                                    // ZPACK_ERROR_FILE_TOO_SMALL
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (!ctx->h_stage.ensure(4096)) return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    u8 *h = (u8 *)ctx->h_stage.p;
    CK(ctx, cudaMemcpyAsync(h, d_archive, 10, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaMemcpyAsync(h + 16, d_archive + archive_size - 12, 12, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    auto le32 = [](const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); };
    auto le64 = [&](const u8 *p) { return (u64)le32(p) | ((u64)le32(p + 4) << 32); };
    if (le32(h) != 0x154B505Au) { *result = 6; return ZPB_OK; }                               // SIGNATURE_INVALID
    if (((u32)h[4] | ((u32)h[5] << 8)) != 1u) { *result = 9; return ZPB_OK; }                 // VERSION_INCOMPATIBLE
    if (le32(h + 6) != 0x144B505Au) { *result = 6; return ZPB_OK; }
    if (le32(h + 16) != 0x124B505Au) { *result = 6; return ZPB_OK; }
    const u64 cdr_off = le64(h + 20);
    if (cdr_off >= archive_size) { *result = 7; return ZPB_OK; }                              // READ_FAILED
    const u64 size_left = archive_size - cdr_off;
    if (size_left < ARC_CDR_HDR) { *result = 8; return ZPB_OK; }       // the reference reads the header past the end here; BLOCK_SIZE_INVALID
    CK(ctx, cudaMemcpyAsync(h + 32, d_archive + cdr_off, ARC_CDR_HDR, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    if (le32(h + 32) != 0x134B505Au) { *result = 6; return ZPB_OK; }
    const u64 count = le64(h + 36), B = le64(h + 44);
    if (B > size_left - ARC_CDR_HDR) { *result = 8; return ZPB_OK; }                          // BLOCK_SIZE_INVALID
    *n_out = count;
    if (names_size) *names_size = B;
    if (count == 0) { *result = 0; return ZPB_OK; }
    if (count > B / ARC_FIXED) { *result = 8; return ZPB_OK; }
    if (B > CDR_MAX_BODY) return fail(ctx, ZPB_E_ARG, "central directory larger than the device parser's 32-bit positions");
    if (!entries || cap < count) return fail(ctx, ZPB_E_ARG, "entry table too small (n is set to the archive's file count)");
    if (names && names_cap < B) return fail(ctx, ZPB_E_ARG, "names buffer smaller than the directory block");

    const u8 *body = d_archive + cdr_off + ARC_CDR_HDR;
    const u64 ntiles = (B + CDR_T - 1) / CDR_T, nsuper = (B + CDR_ST - 1) / CDR_ST;
    if (!ctx->d_cdr_jump.ensure(B * 4 + 16) || !ctx->d_cdr_cnt.ensure(B + 16) || !ctx->d_cdr_j2.ensure(nsuper * CDR_T * 8 + 16) ||
        !ctx->d_cdr_anchor.ensure((nsuper + ntiles) * 8 + 64) || !ctx->d_arc_e.ensure(count * sizeof(ArcEntry) + 64) ||
        !ctx->d_arc_tot.ensure(64) || !ctx->h_stage.ensure(count * sizeof(ArcEntry) + 4096))
        return fail(ctx, ZPB_E_NOMEM, "scratch allocation failed");
    h = (u8 *)ctx->h_stage.p;
    u32 *jump = (u32 *)ctx->d_cdr_jump.p, *jump2 = (u32 *)ctx->d_cdr_j2.p, *cnt2 = jump2 + nsuper * CDR_T;
    u8 *cnt = (u8 *)ctx->d_cdr_cnt.p;
    u32 *sup_pos = (u32 *)ctx->d_cdr_anchor.p, *sup_idx = sup_pos + nsuper, *tile_pos = sup_idx + nsuper, *tile_idx = tile_pos + ntiles;
    u64 *found = (u64 *)ctx->d_arc_tot.p;
    CK(ctx, cudaEventRecord(ctx->evs[0], s));
    CK(ctx, cudaMemsetAsync(ctx->d_cdr_anchor.p, 0xFF, (nsuper + ntiles) * 8, s));
    cdr_tile_kernel<<<(u32)ntiles, 256, CDR_T + 16, s>>>(body, B, jump, cnt);
    CK(ctx, cudaGetLastError());
    cdr_super_kernel<<<(u32)((nsuper * CDR_T + 255) / 256), 256, 0, s>>>(B, jump, cnt, jump2, cnt2, nsuper);
    CK(ctx, cudaGetLastError());
    cdr_chain_kernel<<<1, 32, 0, s>>>(B, jump, cnt, jump2, cnt2, sup_pos, sup_idx, found);
    CK(ctx, cudaGetLastError());
    cdr_anchor_kernel<<<(u32)((nsuper + 63) / 64), 64, 0, s>>>(B, jump, cnt, sup_pos, sup_idx, tile_pos, tile_idx, nsuper);
    CK(ctx, cudaGetLastError());
    cdr_emit_kernel<<<(u32)((ntiles + 63) / 64), 64, 0, s>>>(body, B, tile_pos, tile_idx, (ArcEntry *)ctx->d_arc_e.p, count, ntiles);
    CK(ctx, cudaGetLastError());
    ctx->launches += 5;
    CK(ctx, cudaEventRecord(ctx->evs[1], s));
    CK(ctx, cudaMemcpyAsync(h, found, 8, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    CK(ctx, cudaEventElapsedTime(&ctx->arc_ms[2], ctx->evs[0], ctx->evs[1]));
    u64 nfound;
    memcpy(&nfound, h, 8);
    if (nfound < count) { *result = 8; return ZPB_OK; }               // a record of the first `count` does not fit the block
    CK(ctx, cudaMemcpyAsync(h, ctx->d_arc_e.p, count * sizeof(ArcEntry), cudaMemcpyDeviceToHost, s));
    if (names) CK(ctx, cudaMemcpyAsync(names, body, B, cudaMemcpyDeviceToHost, s));
    CK(ctx, cudaStreamSynchronize(s));
    memcpy(entries, h, count * sizeof(ArcEntry));
    *result = 0;
    return ZPB_OK;
}

extern "C" int zpb_last_archive_ms(const zpb_ctx *ctx, float *ms3) {
    if (!ctx || !ms3) return ZPB_E_ARG;
    for (int k = 0; k < 3; ++k) ms3[k] = ctx->arc_ms[k];
    return ZPB_OK;
}

// ------------------------------------------------------------------------------------ file <-> device
// The file side of a device-resident archive: the reference's fread of an entry's bytes / of the directory block
// (/root/reference/lib/zpack_read.c:298-324, 190-223) and its seek + fwrite of payloads, directory and end record
// (/root/reference/lib/zpack_common.c:72-81; lib/zpack_write.c:308, 393, 747, 800), with HBM on the other side.  No
// GPUDirect-Storage driver is assumed: FILE_IO_THREADS host threads pread / pwrite through pinned buffers of
// FILE_IO_CHUNK bytes, each on its own stream, so that disk (page cache) reads, PCIe copies and the next read overlap.
#include <unistd.h>
#define FILE_IO_THREADS 4
#define FILE_IO_CHUNK (8u << 20)

struct FileIoLane {
    cudaStream_t s = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    u8 *buf[2] = {nullptr, nullptr};
    int rc = ZPB_OK;
    std::string err;
};

static int file_io_device(zpb_ctx *ctx, int fd, u64 file_off, u64 size, u8 *d_ptr, bool to_device) {
    if (!ctx || fd < 0 || (!d_ptr && size)) return fail(ctx, ZPB_E_ARG, "null argument");
    if (size == 0) return ZPB_OK;
    if (file_off + size < file_off) return fail(ctx, ZPB_E_ARG, "offset overflow");
    const u64 nchunk = (size + FILE_IO_CHUNK - 1) / FILE_IO_CHUNK;
    const int nlane = (int)std::min<u64>(FILE_IO_THREADS, nchunk);
    std::vector<FileIoLane> lanes(nlane);
    auto work = [&](int k) {
        FileIoLane &L = lanes[k];
        auto bad = [&](int code, const std::string &m) { L.rc = code; L.err = m; };
        if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking) != cudaSuccess)
            return bad(ZPB_E_CUDA, "stream creation failed");
        for (int b = 0; b < 2; ++b)
            if (cudaMallocHost((void **)&L.buf[b], FILE_IO_CHUNK) != cudaSuccess || cudaEventCreateWithFlags(&L.ev[b], cudaEventDisableTiming) != cudaSuccess)
                return bad(ZPB_E_NOMEM, "pinned buffer allocation failed");
        u64 round = 0;
        for (u64 c = (u64)k; c < nchunk && L.rc == ZPB_OK; c += (u64)nlane, ++round) {
            const int b = (int)(round & 1);
            const u64 off = c * FILE_IO_CHUNK, len = std::min<u64>(FILE_IO_CHUNK, size - off);
            if (round >= 2 && cudaEventSynchronize(L.ev[b]) != cudaSuccess) return bad(ZPB_E_CUDA, "event wait failed");
            if (to_device) {
                for (u64 got = 0; got < len;) {
                    const ssize_t r = pread(fd, L.buf[b] + got, len - got, (off_t)(file_off + off + got));
                    if (r <= 0) return bad(ZPB_E_IO, r == 0 ? "file shorter than the requested range" : "pread failed");
                    got += (u64)r;
                }
                if (cudaMemcpyAsync(d_ptr + off, L.buf[b], len, cudaMemcpyHostToDevice, L.s) != cudaSuccess) return bad(ZPB_E_CUDA, "H2D copy failed");
                cudaEventRecord(L.ev[b], L.s);
            } else {
                // two chunks in flight: the copy of this round's chunk is issued, the previous round's chunk is written out
                if (cudaMemcpyAsync(L.buf[b], d_ptr + off, len, cudaMemcpyDeviceToHost, L.s) != cudaSuccess) return bad(ZPB_E_CUDA, "D2H copy failed");
                cudaEventRecord(L.ev[b], L.s);
                if (round >= 1) {
                    const u64 poff = (c - (u64)nlane) * FILE_IO_CHUNK, plen = std::min<u64>(FILE_IO_CHUNK, size - poff);
                    if (cudaEventSynchronize(L.ev[b ^ 1]) != cudaSuccess) return bad(ZPB_E_CUDA, "event wait failed");
                    for (u64 put = 0; put < plen;) {
                        const ssize_t r = pwrite(fd, L.buf[b ^ 1] + put, plen - put, (off_t)(file_off + poff + put));
                        if (r <= 0) return bad(ZPB_E_IO, "pwrite failed");
                        put += (u64)r;
                    }
                }
            }
        }
        if (L.rc == ZPB_OK && !to_device && round >= 1) {          // the lane's last chunk
            const u64 c = (u64)k + (round - 1) * (u64)nlane, off = c * FILE_IO_CHUNK, len = std::min<u64>(FILE_IO_CHUNK, size - off);
            const int b = (int)((round - 1) & 1);
            if (cudaEventSynchronize(L.ev[b]) != cudaSuccess) return bad(ZPB_E_CUDA, "event wait failed");
            for (u64 put = 0; put < len;) {
                const ssize_t r = pwrite(fd, L.buf[b] + put, len - put, (off_t)(file_off + off + put));
                if (r <= 0) return bad(ZPB_E_IO, "pwrite failed");
                put += (u64)r;
            }
        }
        if (L.s && cudaStreamSynchronize(L.s) != cudaSuccess && L.rc == ZPB_OK) bad(ZPB_E_CUDA, "stream synchronize failed");
    };
    std::vector<std::thread> th;
    for (int k = 1; k < nlane; ++k) th.emplace_back(work, k);
    work(0);
    for (auto &t : th) t.join();
    int rc = ZPB_OK;
    for (FileIoLane &L : lanes) {
        if (L.rc != ZPB_OK && rc == ZPB_OK) { rc = L.rc; ctx->err = L.err; g_last_error = L.err; }
        if (L.s) { cudaStreamSynchronize(L.s); cudaStreamDestroy(L.s); }
        for (int b = 0; b < 2; ++b) { if (L.ev[b]) cudaEventDestroy(L.ev[b]); if (L.buf[b]) cudaFreeHost(L.buf[b]); }
    }
    return rc;
}

extern "C" int zpb_file_read_device(zpb_ctx *ctx, int fd, uint64_t file_off, uint64_t size, uint8_t *d_dst) {
    return file_io_device(ctx, fd, file_off, size, d_dst, true);
}
extern "C" int zpb_file_write_device(zpb_ctx *ctx, int fd, uint64_t file_off, uint64_t size, const uint8_t *d_src) {
    return file_io_device(ctx, fd, file_off, size, const_cast<u8 *>(d_src), false);
}
