// pack_api.inl — host side of zpb_pack_* (included by zpb_api.cu).  Filled in by the pack milestone.
extern "C" uint64_t zpb_pack_bound(uint32_t method, uint64_t size) {
    if (method == ZPB_METHOD_LZ4) return 7 + 4 * ((size + 65535) / 65536) + size + 4;
    if (method == ZPB_METHOD_NONE) return size;
    return 0;
}
extern "C" int zpb_pack_device(zpb_ctx *ctx, const uint8_t *, uint64_t, uint8_t *, uint64_t, const zpb_file *,
                               uint64_t n, uint64_t *, uint64_t *, int32_t *status, void *) {
    for (uint64_t i = 0; status && i < n; ++i) status[i] = ZPB_ST_NOT_AVAILABLE;
    return fail(ctx, ZPB_E_ARG, "pack not built yet");
}
extern "C" int zpb_pack_host(zpb_ctx *ctx, const uint8_t *, uint64_t, uint8_t *, uint64_t, const zpb_file *,
                             uint64_t n, uint64_t *, uint64_t *, int32_t *status) {
    for (uint64_t i = 0; status && i < n; ++i) status[i] = ZPB_ST_NOT_AVAILABLE;
    return fail(ctx, ZPB_E_ARG, "pack not built yet");
}
