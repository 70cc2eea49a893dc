// group_api.inl — one box, several GPUs, one call (part of zpb_api.cu).
//
// Entries are independent (own frame, own digest): a batch is cut into contiguous runs of entries in archive order,
// balanced by decoded bytes (the same rule as zpack_b200/shard.py), and every run goes through
// zpb_unpack_host on its own device context from its own host thread.  No collective and no peer traffic: each GPU
// reads only the byte range of the archive its run touches and writes only its own outputs; the only cross-device
// values are the per-entry status / digest scalars the host already owns.  This is what the drop-in's batched read
// (zpack_read_files) drives, so that a buffer-mode reader uses every visible GPU
// (/root/reference/lib/zpack.h:335-341: readers in buffer mode may be used from several threads, one context each).
struct zpb_group {
    std::vector<zpb_ctx *> ctx;
    std::string err;
    std::vector<float> last_ms;     // wall time of each device's share in the last call
};

extern "C" int zpb_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" zpb_group *zpb_group_create(const int *devices, int n) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { g_last_error = "no CUDA device"; return nullptr; }
    std::vector<int> dev;
    if (!devices || n <= 0) { for (int d = 0; d < count; ++d) dev.push_back(d); }
    else dev.assign(devices, devices + n);
    zpb_group *g = new zpb_group();
    for (int d : dev) {
        zpb_ctx *c = zpb_create(d);
        if (!c) { for (zpb_ctx *x : g->ctx) zpb_destroy(x); delete g; return nullptr; }
        g->ctx.push_back(c);
    }
    g->last_ms.assign(g->ctx.size(), 0.f);
    return g;
}

extern "C" void zpb_group_destroy(zpb_group *g) {
    if (!g) return;
    for (zpb_ctx *c : g->ctx) zpb_destroy(c);
    delete g;
}

extern "C" int zpb_group_size(const zpb_group *g) { return g ? (int)g->ctx.size() : 0; }
extern "C" zpb_ctx *zpb_group_ctx(zpb_group *g, int k) { return g && k >= 0 && k < (int)g->ctx.size() ? g->ctx[k] : nullptr; }
extern "C" const char *zpb_group_last_error(const zpb_group *g) { return g ? g->err.c_str() : g_last_error.c_str(); }

// cuts[k] .. cuts[k+1]: the run of `order` (entries in archive order) that device k takes; balanced on decoded bytes
extern "C" int zpb_group_partition(const zpb_entry *entries, uint64_t n, int world, uint64_t *order, uint64_t *cuts) {
    if ((!entries && n) || world < 1 || !order || !cuts) return ZPB_E_ARG;
    std::iota(order, order + n, (uint64_t)0);
    std::sort(order, order + n, [&](uint64_t a, uint64_t b) { return entries[a].src_off < entries[b].src_off; });
    long double total = 0;
    for (uint64_t i = 0; i < n; ++i) total += entries[i].comp_size ? (long double)entries[i].uncomp_size : 0;
    cuts[0] = 0;
    uint64_t k = 0;
    long double acc = 0;
    for (int r = 1; r < world; ++r) {
        const long double target = total * r / world;
        while (k < n) {
            const long double w = entries[order[k]].comp_size ? (long double)entries[order[k]].uncomp_size : 0;
            if (acc + w > target && (target - acc) < (acc + w - target)) break;   // the nearer boundary
            acc += w; ++k;
            if (acc >= target) break;
        }
        cuts[r] = k;
    }
    cuts[world] = n;
    return ZPB_OK;
}

extern "C" int zpb_group_unpack_host(zpb_group *g, const uint8_t *h_archive, uint64_t archive_size, uint8_t *h_out,
                                     uint64_t out_size, const zpb_entry *entries, uint64_t n, int32_t *status,
                                     uint64_t *digest) {
    if (!g || g->ctx.empty() || (!entries && n)) { if (g) g->err = "null argument"; return ZPB_E_ARG; }
    const int W = (int)g->ctx.size();
    if (n == 0) return ZPB_OK;
    if (W == 1) return zpb_unpack_host(g->ctx[0], h_archive, archive_size, h_out, out_size, entries, n, status, digest);
    std::vector<uint64_t> order(n), cuts(W + 1);
    zpb_group_partition(entries, n, W, order.data(), cuts.data());
    std::vector<int> rcs(W, ZPB_OK);
    auto body = [&](int k) {
        const auto t0 = std::chrono::steady_clock::now();
        const uint64_t a = cuts[k], b = cuts[k + 1], m = b - a;
        if (m) {
            std::vector<zpb_entry> sub(m);
            std::vector<int32_t> st(m);
            std::vector<uint64_t> dg(m);
            for (uint64_t i = 0; i < m; ++i) sub[i] = entries[order[a + i]];
            rcs[k] = zpb_unpack_host(g->ctx[k], h_archive, archive_size, h_out, out_size, sub.data(), m, st.data(), dg.data());
            if (rcs[k] == ZPB_OK)
                for (uint64_t i = 0; i < m; ++i) {
                    if (status) status[order[a + i]] = st[i];
                    if (digest) digest[order[a + i]] = dg[i];
                }
        }
        g->last_ms[k] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    std::vector<std::thread> th;
    for (int k = 1; k < W; ++k) th.emplace_back(body, k);
    body(0);
    for (auto &t : th) t.join();
    for (int k = 0; k < W; ++k)
        if (rcs[k] != ZPB_OK) { g->err = "device " + std::to_string(g->ctx[k]->device) + ": " + g->ctx[k]->err; return rcs[k]; }
    return ZPB_OK;
}

// The same for a packing batch: files are cut into contiguous runs balanced by input bytes.
extern "C" int zpb_group_pack_host(zpb_group *g, const uint8_t *h_in, uint64_t in_size, uint8_t *h_out, uint64_t out_size,
                                   const zpb_file *files, uint64_t n, uint64_t *comp_size, uint64_t *digest, int32_t *status) {
    if (!g || g->ctx.empty() || (!files && n)) { if (g) g->err = "null argument"; return ZPB_E_ARG; }
    const int W = (int)g->ctx.size();
    if (n == 0) return ZPB_OK;
    if (W == 1) return zpb_pack_host(g->ctx[0], h_in, in_size, h_out, out_size, files, n, comp_size, digest, status);
    long double total = 0;
    for (uint64_t i = 0; i < n; ++i) total += files[i].size;
    std::vector<uint64_t> cuts(W + 1, n);
    cuts[0] = 0;
    { uint64_t k = 0; long double acc = 0;
      for (int r = 1; r < W; ++r) { while (k < n && acc + files[k].size / 2.0L < total * r / W) acc += files[k++].size; cuts[r] = k; } }
    std::vector<int> rcs(W, ZPB_OK);
    auto body = [&](int k) {
        const uint64_t a = cuts[k], m = cuts[k + 1] - a;
        if (m) rcs[k] = zpb_pack_host(g->ctx[k], h_in, in_size, h_out, out_size, files + a, m, comp_size + a, digest + a, status + a);
    };
    std::vector<std::thread> th;
    for (int k = 1; k < W; ++k) th.emplace_back(body, k);
    body(0);
    for (auto &t : th) t.join();
    for (int k = 0; k < W; ++k)
        if (rcs[k] != ZPB_OK) { g->err = "device " + std::to_string(g->ctx[k]->device) + ": " + g->ctx[k]->err; return rcs[k]; }
    return ZPB_OK;
}

extern "C" int zpb_group_last_ms(const zpb_group *g, float *ms, int cap) {
    if (!g || !ms) return ZPB_E_ARG;
    for (int k = 0; k < cap && k < (int)g->last_ms.size(); ++k) ms[k] = g->last_ms[k];
    return ZPB_OK;
}

extern "C" void *zpb_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr;
}
extern "C" void zpb_host_free(void *p) { if (p) cudaFreeHost(p); }
