// group.cuh -- ca_group: ONE very long IR split by partition range across the GPUs of a node
// (BASELINE configs[4]).  Included at the end of engine.cu (same translation unit: it
// drives the member engines through their internals).
//
// Per period, fused variant (CA_EXCHANGE_P2P), all launches from one host thread:
//   every GPU g : k_forward (input read straight from mapped pinned host memory, 2 KB)
//                 k_mac over its partition range; the last CTA sums the launch's partial spectra and stores
//                 the result into slot g of the ROOT's gather buffer (peer store over NVLink), then
//                 st.release.sys of the period count into the root's flag g         [kernels.cuh: MacArgs]
//   root        : k_inverse waits (ld.acquire.sys) for flags 1..G-1, sums the G spectra in fixed order, C2R,
//                 clamp, dry mix, stores the block straight into mapped pinned host memory
// One CUDA graph per GPU per period; no collective launch, 4 KB per peer over NVLink.
// CA_EXCHANGE_NCCL is the library baseline of the same math: ncclReduce(sum) of the spectra, then the inverse.
#pragma once
#include <dlfcn.h>

namespace {

// minimal NCCL surface, resolved with dlopen so the product library does not depend on libnccl
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        Reduce = (decltype(Reduce))dlsym(lib, "ncclReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Reduce;
    }
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h, stable ABI)

}  // namespace

struct ca_group {
    ca_group_config cfg{};
    uint32_t G = 0;
    std::vector<ca_engine *> eng;  // [0] = root
    std::vector<int> dev;
    float *h_in = nullptr, *h_out = nullptr;  // pinned + portable + mapped: kernels of every GPU read / write them directly
    int *h_err = nullptr;                     // mapped: the root's "peer timed out" flag
    float2 *gather = nullptr;                 // root memory: [G][n_out][B]
    unsigned long long *flags = nullptr;      // root memory: [G]
    std::vector<uint32_t *> counters;         // per device: CTA counter of the MAC launch
    cudaEvent_t done = nullptr;
    // NCCL variant
    NcclApi nccl;
    std::vector<void *> comms;
    std::vector<float2 *> ylocal;             // per device: [n_out][B] sum of the device's partials
    float2 *ysum = nullptr;                   // root: reduced spectrum
    // stats
    std::vector<float> wall;
    uint64_t periods = 0;
    double wall_sum = 0, wall_max = 0;
};

namespace {

void split_range(uint32_t n, uint32_t world, uint32_t rank, uint32_t *begin, uint32_t *count)
{
    const uint32_t base = n / world, rem = n % world;
    *begin = rank * base + std::min(rank, rem);
    *count = base + (rank < rem ? 1u : 0u);
}

int group_create_impl(const ca_group_config *cfg, ca_group *g)
{
    g->cfg = *cfg;
    g->G = cfg->n_devices;
    const uint32_t B = cfg->period, P = (cfg->max_ir_frames + B - 1) / B;
    if (g->G > P) { g_last_error = "more devices than IR partitions"; return CA_ERR_INVALID; }
    int ndev = 0;
    CA_CUDA(cudaGetDeviceCount(&ndev));
    for (uint32_t i = 0; i < g->G; i++) {
        if (cfg->devices[i] < 0 || cfg->devices[i] >= ndev) { g_last_error = "ca_group: device ordinal out of range"; return CA_ERR_INVALID; }
        for (uint32_t j = 0; j < i; j++)
            if (cfg->devices[j] == cfg->devices[i]) { g_last_error = "ca_group: a device appears twice (kernels that wait on one another must not share a GPU)"; return CA_ERR_INVALID; }
        g->dev.push_back(cfg->devices[i]);
    }
    const bool p2p = cfg->exchange == CA_EXCHANGE_P2P;
    if (!p2p && g->G > 1 && !g->nccl.load()) { g_last_error = "libnccl.so.2 not found"; return CA_ERR_UNSUPPORTED; }
    const size_t spec = (size_t)cfg->n_out * B;  // complex values of one spectrum set
    CA_CUDA(cudaHostAlloc(&g->h_in, (size_t)cfg->n_in * B * sizeof(float), cudaHostAllocPortable | cudaHostAllocMapped));
    CA_CUDA(cudaHostAlloc(&g->h_out, (size_t)cfg->n_out * B * sizeof(float), cudaHostAllocPortable | cudaHostAllocMapped));
    CA_CUDA(cudaHostAlloc(&g->h_err, sizeof(int), cudaHostAllocPortable | cudaHostAllocMapped));
    *g->h_err = 0;
    memset(g->h_in, 0, (size_t)cfg->n_in * B * sizeof(float));
    CA_CUDA(cudaSetDevice(g->dev[0]));
    CA_CUDA(cudaMalloc(&g->gather, g->G * spec * sizeof(float2)));
    CA_CUDA(cudaMemset(g->gather, 0, g->G * spec * sizeof(float2)));
    CA_CUDA(cudaMalloc(&g->flags, 8 * sizeof(unsigned long long)));
    CA_CUDA(cudaMemset(g->flags, 0, 8 * sizeof(unsigned long long)));
    CA_CUDA(cudaMalloc(&g->ysum, spec * sizeof(float2)));
    CA_CUDA(cudaEventCreateWithFlags(&g->done, cudaEventDisableTiming));
    g->counters.assign(g->G, nullptr);
    g->ylocal.assign(g->G, nullptr);
    for (uint32_t i = 0; i < g->G; i++) {
        CA_CUDA(cudaSetDevice(g->dev[i]));
        if (i > 0 && p2p) {
            int can = 0;
            CA_CUDA(cudaDeviceCanAccessPeer(&can, g->dev[i], g->dev[0]));
            if (!can) { g_last_error = "ca_group: no peer access to the root device"; return CA_ERR_UNSUPPORTED; }
            const cudaError_t rc = cudaDeviceEnablePeerAccess(g->dev[0], 0);
            if (rc != cudaSuccess && rc != cudaErrorPeerAccessAlreadyEnabled) CA_CUDA(rc);
            (void)cudaGetLastError();
        }
        CA_CUDA(cudaMalloc(&g->counters[i], sizeof(uint32_t)));
        CA_CUDA(cudaMemset(g->counters[i], 0, sizeof(uint32_t)));
        CA_CUDA(cudaMalloc(&g->ylocal[i], spec * sizeof(float2)));
        CA_CUDA(cudaMemset(g->ylocal[i], 0, spec * sizeof(float2)));
        ca_config ec;
        ca_config_init(&ec);
        ec.device = g->dev[i];
        ec.period = B; ec.n_instances = 1; ec.n_in = cfg->n_in; ec.n_out = cfg->n_out;
        ec.max_ir_frames = cfg->max_ir_frames; ec.n_ir_slots = cfg->n_ir_slots;
        ec.max_voices = cfg->max_voices; ec.sample_rate = cfg->sample_rate;
        ec.flags = (cfg->flags & ~(uint32_t)(CA_FLAG_GRAPH | CA_FLAG_PROFILE | CA_FLAG_ASYNC_TIERS | CA_FLAG_RAW_WET)) | (p2p ? (uint32_t)CA_FLAG_GRAPH : 0u);
        split_range(P, g->G, i, &ec.part_begin, &ec.part_count);
        ca_engine *e = nullptr;
        const int rc = ca_create(&ec, &e);
        if (rc) return rc;
        g->eng.push_back(e);
        ca_engine::Link &l = e->link;
        l.gcount = g->counters[i];
        if (g->G == 1) { l = ca_engine::Link{}; continue; }  // one device: a plain engine
        if (p2p) {
            l.gather = g->gather + i * spec;
            l.gflag = g->flags + i;
            l.skip_inverse = i > 0;
            if (i == 0) { l.inv_src = g->gather; l.inv_split = g->G; l.wait_flags = g->flags + 1; l.n_wait = g->G - 1; l.gerr = g->h_err; }
        } else {
            l.gather = g->ylocal[i];
            l.gflag = g->flags + i;  // unused by anyone; keeps the kernel path identical
            l.skip_inverse = i > 0;
            l.defer_inverse = i == 0;
            if (i == 0) { l.inv_src = g->ysum; l.inv_split = 1; }
        }
        drop_graphs(e);  // captured before the link existed
    }
    if (!p2p && g->G > 1) {
        g->comms.assign(g->G, nullptr);
        const int rc = g->nccl.CommInitAll(g->comms.data(), (int)g->G, g->dev.data());
        if (rc) { g_last_error = std::string("ncclCommInitAll: ") + (g->nccl.GetErrorString ? g->nccl.GetErrorString(rc) : "?"); return CA_ERR_CUDA; }
    }
    g->wall.assign(1u << 16, 0.f);
    return CA_OK;
}

}  // namespace

extern "C" {

void ca_group_config_init(ca_group_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->n_devices = 1;
    cfg->period = 256;
    cfg->n_in = cfg->n_out = 2;
    cfg->max_ir_frames = 60 * 48000;
    cfg->n_ir_slots = 2;
    cfg->exchange = CA_EXCHANGE_P2P;
    cfg->sample_rate = 48000.f;
}

int ca_group_destroy(ca_group *g)
{
    if (!g) return CA_OK;
    for (size_t i = 0; i < g->eng.size(); i++) { cudaSetDevice(g->dev[i]); cudaDeviceSynchronize(); }
    for (auto &c : g->comms) if (c) g->nccl.CommDestroy(c);
    for (auto *e : g->eng) ca_destroy(e);
    for (size_t i = 0; i < g->counters.size(); i++) { cudaSetDevice(g->dev[i]); cudaFree(g->counters[i]); cudaFree(g->ylocal[i]); }
    if (!g->dev.empty()) cudaSetDevice(g->dev[0]);
    cudaFree(g->gather); cudaFree(g->flags); cudaFree(g->ysum);
    if (g->done) cudaEventDestroy(g->done);
    cudaFreeHost(g->h_in); cudaFreeHost(g->h_out); cudaFreeHost(g->h_err);
    delete g;
    return CA_OK;
}

int ca_group_create(const ca_group_config *cfg, ca_group **out)
{
    if (!cfg || !out) return CA_ERR_INVALID;
    *out = nullptr;
    if (cfg->struct_size != sizeof(ca_group_config)) { g_last_error = "ca_group_config.struct_size mismatch"; return CA_ERR_INVALID; }
    if (cfg->n_devices < 1 || cfg->n_devices > 8 || cfg->exchange > CA_EXCHANGE_NCCL) { g_last_error = "ca_group: 1..8 devices, exchange P2P or NCCL"; return CA_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        (void)cudaGetLastError();
        g_last_error = "no usable CUDA device (this engine has no CPU fallback)";
        return CA_ERR_CUDA;
    }
    ca_group *g = new (std::nothrow) ca_group();
    if (!g) return CA_ERR_NOMEM;
    const int rc = group_create_impl(cfg, g);
    if (rc) { std::string keep = g_last_error; ca_group_destroy(g); g_last_error = keep; (void)cudaGetLastError(); return rc; }
    *out = g;
    return CA_OK;
}

int ca_group_load_ir(ca_group *g, uint32_t slot, const float *left, const float *right, uint32_t frames)
{
    if (!g) return CA_ERR_INVALID;
    for (auto *e : g->eng) {  // every member keeps only its partition range (IrArgs::frame_off)
        const int rc = ca_load_ir(e, slot, left, right, frames);
        if (rc) return rc;
    }
    return CA_OK;
}

int ca_group_set_params(ca_group *g, uint32_t input, const ca_params *p)
{
    if (!g) return CA_ERR_INVALID;
    for (auto *e : g->eng) {  // the wet path is linear in the IR: the same parameters on every member
        const int rc = ca_set_params(e, 0, input, p);
        if (rc) return rc;
    }
    return CA_OK;
}

int ca_group_set_glide(ca_group *g, uint32_t input, float glide)
{
    if (!g) return CA_ERR_INVALID;
    for (auto *e : g->eng) {
        const int rc = ca_set_glide(e, 0, input, glide);
        if (rc) return rc;
    }
    return CA_OK;
}

int ca_group_reset(ca_group *g)
{
    if (!g) return CA_ERR_INVALID;
    for (uint32_t i = 0; i < g->G; i++) {  // every member restarts at the same period: their counters stay in step
        const int rc = ca_reset(g->eng[i]);  // sets the device, drains, synchronises
        if (rc) return rc;
    }
    return CA_OK;
}

int ca_group_process(ca_group *g, const float *in, float *out, uint32_t nframes)
{
    if (!g || !in || !out) return CA_ERR_INVALID;
    if (nframes != g->cfg.period) return CA_ERR_PERIOD;
    const double t0 = now_us();
    const size_t in_bytes = (size_t)g->cfg.n_in * nframes * sizeof(float), out_bytes = (size_t)g->cfg.n_out * nframes * sizeof(float);
    memcpy(g->h_in, in, in_bytes);  // mapped pinned memory: the forward kernels of every GPU read it in place
    const bool nccl = g->cfg.exchange == CA_EXCHANGE_NCCL && g->G > 1;
    for (uint32_t i = 0; i < g->G; i++) {
        ca_engine *e = g->eng[i];
        CA_CUDA(cudaSetDevice(g->dev[i]));
        const int rc = run_period(e, g->h_in, i == 0 ? g->h_out : e->d_out);
        if (rc) return rc;
    }
    if (nccl) {
        const size_t count = (size_t)g->cfg.n_out * nframes * 2;  // floats of one spectrum set
        g->nccl.GroupStart();
        for (uint32_t i = 0; i < g->G; i++)
            g->nccl.Reduce(g->ylocal[i], i == 0 ? (void *)g->ysum : (void *)g->ylocal[i], count, kNcclFloat32, kNcclSum, 0, g->comms[i], g->eng[i]->stream);
        const int rc = g->nccl.GroupEnd();
        if (rc) { g_last_error = std::string("ncclReduce: ") + (g->nccl.GetErrorString ? g->nccl.GetErrorString(rc) : "?"); return CA_ERR_CUDA; }
        CA_CUDA(cudaSetDevice(g->dev[0]));
        const int rc2 = launch_inverse_phase(g->eng[0], g->h_in, g->h_out);
        if (rc2) return rc2;
    }
    for (uint32_t i = 0; i < g->G; i++) {
        CA_CUDA(cudaSetDevice(g->dev[i]));
        const int rc = run_deferred(g->eng[i]);
        if (rc) return rc;
    }
    CA_CUDA(cudaSetDevice(g->dev[0]));
    CA_CUDA(cudaEventRecord(g->done, g->eng[0]->stream));
    CA_CUDA(cudaEventSynchronize(g->done));
    memcpy(out, g->h_out, out_bytes);
    const double us = now_us() - t0;
    g->wall[g->periods % g->wall.size()] = (float)us;
    g->periods++;
    g->wall_sum += us;
    g->wall_max = std::max(g->wall_max, us);
    return CA_OK;
}

int ca_group_get_stats(ca_group *g, ca_group_stats *s)
{
    if (!g || !s) return CA_ERR_INVALID;
    memset(s, 0, sizeof(*s));
    s->periods = g->periods;
    const size_t n = (size_t)std::min<uint64_t>(g->periods, g->wall.size());
    if (n) {
        std::vector<float> v(g->wall.begin(), g->wall.begin() + n);
        std::sort(v.begin(), v.end());
        s->p50_us = v[n / 2];
        s->p99_us = v[std::min(n - 1, (size_t)std::ceil(0.99 * (double)n))];
        s->max_us = g->wall_max;
        s->mean_us = g->wall_sum / (double)g->periods;
    }
    s->n_devices = g->G; s->exchange = g->cfg.exchange;
    for (uint32_t i = 0; i < g->G; i++) {
        ca_stats es;
        ca_get_stats(g->eng[i], &es);
        s->part_begin[i] = g->eng[i]->k_off; s->part_count[i] = es.partitions;
        s->mac_split[i] = es.mac_split; s->mac_bytes[i] = es.mac_bytes;
        s->gpu_launches += es.gpu_launches;
    }
    s->exchange_bytes_per_peer = g->G > 1 ? (uint64_t)g->cfg.n_out * g->cfg.period * sizeof(float2) : 0;
    s->peer_timeout = *g->h_err;
    return CA_OK;
}

int ca_group_reset_stats(ca_group *g)
{
    if (!g) return CA_ERR_INVALID;
    g->periods = 0; g->wall_sum = g->wall_max = 0;
    return CA_OK;
}

}  // extern "C"
