// engine.cu -- host runtime + C ABI (include/cuda_audio_b200.h) of the B200 convolution engine.
//
// Owns the device memory (IR spectra bank, frequency-domain delay lines, predelay rings,
// parameter blocks), the CUDA stream / graph of the per-period pipeline and the statistics.
// No cuFFT, no CPU fallback: if CUDA is not usable every entry point returns CA_ERR_CUDA.
#include "../../include/cuda_audio_b200.h"
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace ca;

namespace {

thread_local std::string g_last_error;

#define CA_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _rc = (expr);                                                                  \
        if (_rc != cudaSuccess) {                                                                  \
            char _buf[512];                                                                        \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_rc)); \
            g_last_error = _buf;                                                                   \
            return _rc == cudaErrorMemoryAllocation ? CA_ERR_NOMEM : CA_ERR_CUDA;                  \
        }                                                                                          \
    } while (0)

typedef void (*fwd_fn)(const FwdArgs);
typedef void (*ir_fn)(const IrArgs);
typedef void (*mac_fn)(const MacArgs);
typedef void (*inv_fn)(const InvArgs);

struct MacVariant { mac_fn fn; uint32_t smem; int kc; };

template <int BT, int NOUT, int MULT, int NSTAGE>
MacVariant mac_variant()
{
    constexpr int G = kMacConsumers / (BT / 2);
    using Cfg = MacCfg<BT, NOUT, G * MULT, NSTAGE>;
    return MacVariant{k_mac<BT, NOUT, G * MULT, NSTAGE>, Cfg::SMEM_BYTES, G * MULT};
}

// stage = G*MULT rows of (1 + NOUT) arrays of BT complex; BT = 256, NOUT = 2: 12 KB * MULT
template <int BT, int NOUT>
MacVariant mac_pick_v(int variant)
{
    switch (variant) {
    case 0: return mac_variant<BT, NOUT, 4, 4>();   // 192 KB: 1 CTA / SM
    case 2: return mac_variant<BT, NOUT, 2, 3>();   //  72 KB: 3 CTAs / SM
    case 3: return mac_variant<BT, NOUT, 4, 2>();   //  96 KB, longer stages
    case 4: return mac_variant<BT, NOUT, 1, 4>();   //  48 KB: 4 CTAs / SM
    case 5: return mac_variant<BT, NOUT, 2, 6>();   // 144 KB: 1 CTA / SM, deep
    default: return mac_variant<BT, NOUT, 2, 4>();  //  96 KB: 2 CTAs / SM (measured best)
    }
}

MacVariant mac_pick(int bt, int n_out, int variant)
{
    switch (bt) {
    case 32: return n_out == 1 ? mac_pick_v<32, 1>(variant) : mac_pick_v<32, 2>(variant);
    case 64: return n_out == 1 ? mac_pick_v<64, 1>(variant) : mac_pick_v<64, 2>(variant);
    case 128: return n_out == 1 ? mac_pick_v<128, 1>(variant) : mac_pick_v<128, 2>(variant);
    default: return n_out == 1 ? mac_pick_v<256, 1>(variant) : mac_pick_v<256, 2>(variant);
    }
}

struct FftFns { fwd_fn fwd; ir_fn ir; inv_fn inv; };
FftFns fft_pick(int R)
{
    switch (R) {
    case 1: return {k_forward<1>, k_ir_fft<1>, k_inverse<1>};
    case 2: return {k_forward<2>, k_ir_fft<2>, k_inverse<2>};
    case 4: return {k_forward<4>, k_ir_fft<4>, k_inverse<4>};
    case 8: return {k_forward<8>, k_ir_fft<8>, k_inverse<8>};
    case 16: return {k_forward<16>, k_ir_fft<16>, k_inverse<16>};
    case 32: return {k_forward<32>, k_ir_fft<32>, k_inverse<32>};
    default: return {nullptr, nullptr, nullptr};
    }
}

double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct ca_engine {
    ca_config cfg{};
    int device = 0;
    uint32_t B = 0, R = 0, P = 0, Lring = 0, k_off = 0, tiles = 1, bt = 0;
    uint32_t n_inst = 0, n_active = 0, n_in = 0, n_out = 0, n_split = 1, nv = 2, ring_len = 16384, ring_out = 0;
    cudaStream_t stream = nullptr;
    // device memory
    unsigned char *d_arena = nullptr;  // [H | X] contiguous (one L2 access-policy window)
    size_t arena_bytes = 0, h_bytes = 0, x_bytes = 0;
    float2 *d_H = nullptr, *d_X = nullptr, *d_Ypart = nullptr, *d_tw = nullptr;
    float *d_ring = nullptr, *d_in = nullptr, *d_out = nullptr;
    InParamDev *d_par = nullptr;
    ItemState *d_st = nullptr;
    Ctl *d_ctl = nullptr;
    uint64_t device_bytes = 0;
    // pinned host staging
    float *h_in = nullptr, *h_out = nullptr;
    InParamDev *h_upload[2] = {nullptr, nullptr};
    cudaEvent_t upload_done[2] = {nullptr, nullptr};
    int upload_idx = 0;
    // parameters (host shadow)
    std::mutex par_mutex;
    std::vector<InParamDev> par;
    std::vector<ca_params> user;
    std::atomic<bool> par_dirty{true};
    std::vector<uint8_t> ir_loaded;
    // kernels
    FftFns fft{};
    MacVariant mac{};
    // graph
    cudaGraphExec_t gexec = nullptr;
    const float *g_in = nullptr;
    float *g_out = nullptr;
    uint32_t g_active = 0;
    // profiling
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double prof_us[3] = {0, 0, 0};
    uint64_t prof_n = 0;
    // stats
    std::vector<float> wall;  // ring of host wall times (us)
    uint64_t periods = 0, xruns = 0, launches = 0;
    double wall_sum = 0, wall_max = 0, deadline_us = 0;
    // pinned-pointer cache
    const void *pin_ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    bool pin_val[4] = {false, false, false, false};
    int pin_next = 0;
};

namespace {

bool is_pinned(ca_engine *e, const void *p)
{
    for (int i = 0; i < 4; i++)
        if (e->pin_ptr[i] == p) return e->pin_val[i];
    cudaPointerAttributes at{};
    bool pinned = false;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess) pinned = (at.type == cudaMemoryTypeHost);
    else (void)cudaGetLastError();
    e->pin_ptr[e->pin_next] = p;
    e->pin_val[e->pin_next] = pinned;
    e->pin_next = (e->pin_next + 1) & 3;
    return pinned;
}

void fill_dev_param(InParamDev &d, const ca_params &u)
{
    d.wet = u.wet; d.dry = u.dry; d.level = u.level; d.panWet = u.panWet; d.panDry = u.panDry;
    d.predelay = u.predelay; d.select = u.select;
}

int flush_params(ca_engine *e)
{
    if (!e->par_dirty.load(std::memory_order_acquire)) return CA_OK;
    std::unique_lock<std::mutex> lk(e->par_mutex, std::try_to_lock);
    if (!lk.owns_lock()) return CA_OK;  // a setter is mid-update: pick it up next period (never block the RT thread)
    const int b = e->upload_idx;
    e->upload_idx ^= 1;
    CA_CUDA(cudaEventSynchronize(e->upload_done[b]));
    const size_t bytes = e->par.size() * sizeof(InParamDev);
    memcpy(e->h_upload[b], e->par.data(), bytes);
    e->par_dirty.store(false, std::memory_order_release);
    lk.unlock();
    CA_CUDA(cudaMemcpyAsync(e->d_par, e->h_upload[b], bytes, cudaMemcpyHostToDevice, e->stream));
    CA_CUDA(cudaEventRecord(e->upload_done[b], e->stream));
    return CA_OK;
}

int launch_kernels(ca_engine *e, const float *d_in, float *d_out, bool profile)
{
    const uint32_t n_items = e->n_active * e->n_in;
    const uint32_t n_alloc = e->n_inst * e->n_in;
    FwdArgs fa{d_in, e->d_ring, e->d_X, e->d_par, e->d_st, e->d_ctl, e->d_tw, e->d_tw + e->B,
               n_items, n_alloc, e->n_in, e->nv, e->Lring, e->ring_len, e->ring_out};
    MacArgs ma{e->d_X, e->d_H, e->d_Ypart, e->d_par, e->d_st, e->d_ctl, n_alloc, e->n_in, e->nv, e->Lring, e->P, e->B, e->k_off,
               1u, 1u, e->n_split, (e->cfg.flags & CA_FLAG_STREAMING) ? 1u : 0u};
    InvArgs ia{e->d_Ypart, d_in, d_out, nullptr, e->d_par, e->d_ctl, e->d_tw, e->d_tw + e->B, e->n_split, e->n_in, e->n_out, 0u};
    if (profile) CA_CUDA(cudaEventRecord(e->ev[0], e->stream));
    e->fft.fwd<<<(n_items * e->nv + kFwdWarps - 1) / kFwdWarps, kFwdWarps * 32, 0, e->stream>>>(fa);
    if (profile) CA_CUDA(cudaEventRecord(e->ev[1], e->stream));
    e->mac.fn<<<dim3(e->n_split, e->tiles, e->n_active), kMacThreads, e->mac.smem, e->stream>>>(ma);
    if (profile) CA_CUDA(cudaEventRecord(e->ev[2], e->stream));
    e->fft.inv<<<e->n_active * e->n_out, kInvThreads, 0, e->stream>>>(ia);
    if (profile) CA_CUDA(cudaEventRecord(e->ev[3], e->stream));
    CA_CUDA(cudaGetLastError());
    e->launches += 3;
    return CA_OK;
}

int run_period(ca_engine *e, const float *d_in, float *d_out)
{
    int rc = flush_params(e);
    if (rc) return rc;
    const bool profile = (e->cfg.flags & CA_FLAG_PROFILE) != 0;
    if ((e->cfg.flags & CA_FLAG_GRAPH) && !profile) {
        if (!e->gexec || e->g_in != d_in || e->g_out != d_out || e->g_active != e->n_active) {
            if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
            cudaGraph_t g = nullptr;
            CA_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            rc = launch_kernels(e, d_in, d_out, false);
            e->launches -= 3;
            cudaError_t erc = cudaStreamEndCapture(e->stream, &g);
            if (rc) return rc;
            CA_CUDA(erc);
            CA_CUDA(cudaGraphInstantiate(&e->gexec, g, 0));
            cudaGraphDestroy(g);
            e->g_in = d_in; e->g_out = d_out; e->g_active = e->n_active;
        }
        CA_CUDA(cudaGraphLaunch(e->gexec, e->stream));
        e->launches += 3;
        return CA_OK;
    }
    rc = launch_kernels(e, d_in, d_out, profile);
    if (rc) return rc;
    if (profile) {
        CA_CUDA(cudaEventSynchronize(e->ev[3]));
        for (int i = 0; i < 3; i++) {
            float ms = 0;
            CA_CUDA(cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
            e->prof_us[i] += 1e3 * ms;
        }
        e->prof_n++;
    }
    return CA_OK;
}

void record_wall(ca_engine *e, double us)
{
    e->wall[e->periods % e->wall.size()] = (float)us;
    e->periods++;
    e->wall_sum += us;
    if (us > e->wall_max) e->wall_max = us;
    if (e->deadline_us > 0 && us > e->deadline_us) e->xruns++;
}

bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

}  // namespace

extern "C" {

int ca_api_version(void) { return CA_API_VERSION; }

const char *ca_strerror(int code)
{
    switch (code) {
    case CA_OK: return "ok";
    case CA_ERR_INVALID: return "invalid argument";
    case CA_ERR_CUDA: return "CUDA failure";
    case CA_ERR_NOMEM: return "out of memory";
    case CA_ERR_STATE: return "invalid state (IR slot not loaded?)";
    case CA_ERR_UNSUPPORTED: return "unsupported configuration";
    case CA_ERR_PERIOD: return "nframes does not match the configured period";
    default: return "unknown error";
    }
}

const char *ca_last_error_string(void) { return g_last_error.c_str(); }

void ca_config_init(ca_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->period = 256;
    cfg->n_instances = 1;
    cfg->n_in = cfg->n_out = 2;
    cfg->max_ir_frames = 512 * 256 - 1024;  // CONV_DEFAULT_FFTSIZE - default nframes, conv.h:10-12,63
    cfg->n_ir_slots = 2;
    cfg->sample_rate = 48000.f;
}

int ca_destroy(ca_engine *e)
{
    if (!e) return CA_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->gexec) cudaGraphExecDestroy(e->gexec);
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->upload_done) if (ev) cudaEventDestroy(ev);
    cudaFree(e->d_arena); cudaFree(e->d_Ypart); cudaFree(e->d_tw); cudaFree(e->d_ring);
    cudaFree(e->d_in); cudaFree(e->d_out); cudaFree(e->d_par); cudaFree(e->d_st); cudaFree(e->d_ctl);
    cudaFreeHost(e->h_in); cudaFreeHost(e->h_out); cudaFreeHost(e->h_upload[0]); cudaFreeHost(e->h_upload[1]);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return CA_OK;
}

static int create_impl(const ca_config *cfg, ca_engine *e)
{
    e->cfg = *cfg;
    e->device = cfg->device;
    CA_CUDA(cudaSetDevice(cfg->device));
    e->B = cfg->period; e->R = cfg->period / 32;
    e->n_inst = e->n_active = cfg->n_instances; e->n_in = cfg->n_in; e->n_out = cfg->n_out;
    const uint32_t P_total = (cfg->max_ir_frames + e->B - 1) / e->B;
    if (cfg->part_count) {
        if (cfg->part_begin + cfg->part_count > P_total) { g_last_error = "partition shard exceeds the IR"; return CA_ERR_INVALID; }
        e->P = cfg->part_count; e->k_off = cfg->part_begin;
    } else { e->P = P_total; e->k_off = 0; }
    e->Lring = e->k_off + e->P;
    e->bt = std::min<uint32_t>(e->B, 256);
    e->tiles = e->B / e->bt;
    e->fft = fft_pick((int)e->R);
    int variant = 1;
    if (const char *v = getenv("CA_MAC_VARIANT")) variant = atoi(v);
    e->mac = mac_pick((int)e->bt, (int)e->n_out, variant);
    e->nv = cfg->max_voices ? cfg->max_voices : 2u;
    CA_CUDA(cudaFuncSetAttribute((const void *)e->mac.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->mac.smem));

    // split of the row list per instance: enough CTAs to cover the SMs when few instances run
    // (latency schedule), 1 when the batch alone fills the machine.
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    uint32_t split = cfg->mac_split;
    if (!split) {
        const uint64_t ctas = (uint64_t)e->n_inst * e->tiles;
        split = ctas >= (uint64_t)2 * sms ? 1u : (uint32_t)std::min<uint64_t>(32, (2 * (uint64_t)sms + ctas - 1) / ctas);
    }
    const uint32_t rows = e->P * e->n_in;  // steady state: one voice per input
    e->n_split = std::max<uint32_t>(1, std::min(split, std::max<uint32_t>(1, rows / (uint32_t)e->mac.kc)));
    // time-domain ring per voice: predelay reach + the two blocks of the overlap-save window
    e->ring_len = 1;
    while (e->ring_len < kMaxPredelay + 2 * e->B) e->ring_len <<= 1;
    e->ring_out = std::max(e->Lring, e->ring_len / e->B) + 2;

    CA_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (auto &ev : e->ev) CA_CUDA(cudaEventCreate(&ev));
    for (auto &ev : e->upload_done) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

    const size_t n_items = (size_t)e->n_inst * e->n_in;
    e->h_bytes = (size_t)cfg->n_ir_slots * e->n_out * e->P * e->B * sizeof(float2);
    e->x_bytes = n_items * e->nv * e->Lring * e->B * sizeof(float2);
    e->arena_bytes = e->h_bytes + e->x_bytes;
    CA_CUDA(cudaMalloc(&e->d_arena, e->arena_bytes));
    e->d_H = reinterpret_cast<float2 *>(e->d_arena);
    e->d_X = reinterpret_cast<float2 *>(e->d_arena + e->h_bytes);
    CA_CUDA(cudaMemsetAsync(e->d_arena, 0, e->arena_bytes, e->stream));
    const size_t yp_bytes = (size_t)e->n_inst * e->n_split * e->n_out * e->B * sizeof(float2);
    CA_CUDA(cudaMalloc(&e->d_Ypart, yp_bytes));
    CA_CUDA(cudaMemsetAsync(e->d_Ypart, 0, yp_bytes, e->stream));
    const size_t ring_bytes = n_items * e->nv * e->ring_len * sizeof(float);
    CA_CUDA(cudaMalloc(&e->d_ring, ring_bytes));
    CA_CUDA(cudaMemsetAsync(e->d_ring, 0, ring_bytes, e->stream));
    const size_t in_bytes = n_items * e->B * sizeof(float), out_bytes = (size_t)e->n_inst * e->n_out * e->B * sizeof(float);
    CA_CUDA(cudaMalloc(&e->d_in, in_bytes));
    CA_CUDA(cudaMalloc(&e->d_out, out_bytes));
    CA_CUDA(cudaMalloc(&e->d_par, n_items * sizeof(InParamDev)));
    CA_CUDA(cudaMalloc(&e->d_st, 2 * n_items * sizeof(ItemState)));
    CA_CUDA(cudaMemsetAsync(e->d_st, 0, 2 * n_items * sizeof(ItemState), e->stream));
    CA_CUDA(cudaMalloc(&e->d_ctl, sizeof(Ctl)));
    CA_CUDA(cudaMemsetAsync(e->d_ctl, 0, sizeof(Ctl), e->stream));
    CA_CUDA(cudaMallocHost(&e->h_in, in_bytes));
    CA_CUDA(cudaMallocHost(&e->h_out, out_bytes));
    for (auto &u : e->h_upload) CA_CUDA(cudaMallocHost(&u, n_items * sizeof(InParamDev)));
    e->device_bytes = e->arena_bytes + yp_bytes + ring_bytes + in_bytes + out_bytes + n_items * (sizeof(InParamDev) + 2 * sizeof(ItemState));

    // twiddles, fp64 -> fp32: [W_M^n, n < M | W_2M^k, k < M]
    {
        const uint32_t M = e->B;
        std::vector<float2> tw(2 * M);
        for (uint32_t n = 0; n < M; n++) {
            const double a = -2.0 * M_PI * (double)n / (double)M, b = -M_PI * (double)n / (double)M;
            tw[n] = make_float2((float)cos(a), (float)sin(a));
            tw[M + n] = make_float2((float)cos(b), (float)sin(b));
        }
        CA_CUDA(cudaMalloc(&e->d_tw, tw.size() * sizeof(float2)));
        CA_CUDA(cudaMemcpy(e->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }

    // parameter defaults == Convolution::CC::value defaults (conv.h:42-50)
    e->par.assign(n_items, InParamDev{});
    e->user.assign(n_items, ca_params{});
    for (size_t i = 0; i < n_items; i++) {
        ca_params &u = e->user[i];
        u.select = 0; u.predelay = 0; u.speed = 100; u.vsteps = -1;
        u.dry = 0.5f; u.wet = 0.5f; u.panDry = 0.f; u.panWet = 0.f; u.level = 1.0f;
        fill_dev_param(e->par[i], u);
    }
    e->ir_loaded.assign(cfg->n_ir_slots, 0);
    e->wall.assign(1u << 16, 0.f);
    e->deadline_us = cfg->sample_rate > 0 ? 1e6 * (double)e->B / (double)cfg->sample_rate : 0.0;

    if (cfg->flags & CA_FLAG_L2_PERSIST) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, cfg->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, cfg->device);
        if (max_persist > 0 && max_window > 0) {
            const size_t want = std::min<size_t>(e->arena_bytes, (size_t)max_persist);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = e->d_arena;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(e->arena_bytes, (size_t)max_window);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)av.accessPolicyWindow.num_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
        }
    }
    CA_CUDA(cudaStreamSynchronize(e->stream));
    return CA_OK;
}

int ca_create(const ca_config *cfg, ca_engine **out)
{
    if (!cfg || !out) return CA_ERR_INVALID;
    *out = nullptr;
    if (cfg->struct_size != sizeof(ca_config)) { g_last_error = "ca_config.struct_size mismatch"; return CA_ERR_INVALID; }
    if (!is_pow2(cfg->period) || cfg->period < 32 || cfg->period > 1024) { g_last_error = "period must be a power of two in [32, 1024]"; return CA_ERR_INVALID; }
    if (cfg->n_in < 1 || cfg->n_in > 2 || cfg->n_out < 1 || cfg->n_out > 2) { g_last_error = "n_in / n_out must be 1 or 2"; return CA_ERR_INVALID; }
    if (cfg->max_voices > (uint32_t)kMaxVoices) { g_last_error = "max_voices must be <= 4"; return CA_ERR_INVALID; }
    if (!cfg->n_instances || !cfg->max_ir_frames || !cfg->n_ir_slots) { g_last_error = "n_instances, max_ir_frames, n_ir_slots must be > 0"; return CA_ERR_INVALID; }
    if (cfg->n_tiers > 1) { g_last_error = "non-uniform tiers are not available in this build"; return CA_ERR_UNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
        (void)cudaGetLastError();
        g_last_error = "no usable CUDA device (this engine has no CPU fallback)";
        return CA_ERR_CUDA;
    }
    ca_engine *e = new (std::nothrow) ca_engine();
    if (!e) return CA_ERR_NOMEM;
    int rc = create_impl(cfg, e);
    if (rc) { std::string keep = g_last_error; ca_destroy(e); g_last_error = keep; (void)cudaGetLastError(); return rc; }
    *out = e;
    return CA_OK;
}

static int load_ir_dev(ca_engine *e, uint32_t slot, const float *d_left, const float *d_right, uint32_t frames, uint32_t stride)
{
    if (!e || !d_left || slot >= e->cfg.n_ir_slots) return CA_ERR_INVALID;
    if (e->n_out == 2 && !d_right) return CA_ERR_INVALID;
    CA_CUDA(cudaSetDevice(e->device));
    IrArgs a{};
    a.stride = stride;
    a.h[0] = d_left; a.h[1] = d_right ? d_right : d_left;
    a.H = e->d_H + (size_t)slot * e->n_out * e->P * e->B;
    a.twM = e->d_tw; a.tw2M = e->d_tw + e->B;
    a.frames = std::min(frames, e->cfg.max_ir_frames);  // truncation like conv.cu:239
    a.P = e->P; a.n_out = e->n_out; a.k_begin = e->k_off;
    a.scale = 1.0f / (2.0f * (float)e->B);
    const uint32_t items = e->n_out * e->P;
    e->fft.ir<<<(items + kFwdWarps - 1) / kFwdWarps, kFwdWarps * 32, 0, e->stream>>>(a);
    CA_CUDA(cudaGetLastError());
    CA_CUDA(cudaStreamSynchronize(e->stream));
    e->launches += 1;
    e->ir_loaded[slot] = 1;
    return CA_OK;
}

int ca_load_ir_device(ca_engine *e, uint32_t slot, const float *d_left, const float *d_right, uint32_t frames)
{
    return load_ir_dev(e, slot, d_left, d_right, frames, 1);
}

int ca_load_ir_interleaved_device(ca_engine *e, uint32_t slot, const float *d_lr, uint32_t frames)
{
    if (!d_lr) return CA_ERR_INVALID;
    return load_ir_dev(e, slot, d_lr, d_lr + 1, frames, 2);
}

int ca_load_ir(ca_engine *e, uint32_t slot, const float *left, const float *right, uint32_t frames)
{
    if (!e || !left || slot >= e->cfg.n_ir_slots || !frames) return CA_ERR_INVALID;
    if (e->n_out == 2 && !right) return CA_ERR_INVALID;
    CA_CUDA(cudaSetDevice(e->device));
    const uint32_t n = std::min(frames, e->cfg.max_ir_frames);
    float *d = nullptr;
    CA_CUDA(cudaMalloc(&d, (size_t)2 * n * sizeof(float)));
    cudaError_t rc = cudaMemcpy(d, left, (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
    if (rc == cudaSuccess && right) rc = cudaMemcpy(d + n, right, (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
    int r = CA_OK;
    if (rc != cudaSuccess) { g_last_error = cudaGetErrorString(rc); r = CA_ERR_CUDA; }
    else r = ca_load_ir_device(e, slot, d, right ? d + n : nullptr, n);
    cudaFree(d);
    return r;
}

int ca_set_params(ca_engine *e, uint32_t instance, uint32_t input, const ca_params *p)
{
    if (!e || !p || instance >= e->n_inst || input >= e->n_in) return CA_ERR_INVALID;
    if (p->select >= e->cfg.n_ir_slots || p->predelay >= CA_MAX_PREDELAY) return CA_ERR_INVALID;
    if (!e->ir_loaded[p->select]) return CA_ERR_STATE;  // the reference would dereference nullptr (conv.cu:340)
    std::lock_guard<std::mutex> lk(e->par_mutex);
    const size_t i = (size_t)instance * e->n_in + input;
    e->user[i] = *p;
    e->user[i].vsteps = -1;
    InParamDev &d = e->par[i];
    fill_dev_param(d, *p);
    if (p->vsteps >= 0) { d.vsteps_cmd = (uint32_t)p->vsteps; d.vsteps_seq++; }
    e->par_dirty.store(true, std::memory_order_release);
    return CA_OK;
}

int ca_get_params(ca_engine *e, uint32_t instance, uint32_t input, ca_params *p)
{
    if (!e || !p || instance >= e->n_inst || input >= e->n_in) return CA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->par_mutex);
    *p = e->user[(size_t)instance * e->n_in + input];
    return CA_OK;
}

int ca_set_glide(ca_engine *e, uint32_t instance, uint32_t input, float g)
{
    if (!e || instance >= e->n_inst || input >= e->n_in) return CA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->par_mutex);
    InParamDev &d = e->par[(size_t)instance * e->n_in + input];
    d.glide_cmd = g;
    d.glide_seq++;
    e->par_dirty.store(true, std::memory_order_release);
    return CA_OK;
}

int ca_set_active(ca_engine *e, uint32_t n)
{
    if (!e || !n || n > e->n_inst) return CA_ERR_INVALID;
    e->n_active = n;
    return CA_OK;
}

int ca_process_device(ca_engine *e, const float *d_in, float *d_out, uint32_t nframes)
{
    if (!e || !d_in || !d_out) return CA_ERR_INVALID;
    if (nframes != e->B) return CA_ERR_PERIOD;
    const double t0 = now_us();
    int rc = run_period(e, d_in, d_out);
    if (rc) return rc;
    record_wall(e, now_us() - t0);
    return CA_OK;
}

int ca_process(ca_engine *e, const float *in, float *out, uint32_t nframes)
{
    if (!e || !in || !out) return CA_ERR_INVALID;  // the reference silently returns on null ports (conv.cu:297)
    if (nframes != e->B) return CA_ERR_PERIOD;
    const double t0 = now_us();
    const size_t in_bytes = (size_t)e->n_active * e->n_in * e->B * sizeof(float);
    const size_t out_bytes = (size_t)e->n_active * e->n_out * e->B * sizeof(float);
    const float *src = in;
    if (!is_pinned(e, in)) { memcpy(e->h_in, in, in_bytes); src = e->h_in; }
    float *dst = is_pinned(e, out) ? out : e->h_out;
    CA_CUDA(cudaMemcpyAsync(e->d_in, src, in_bytes, cudaMemcpyHostToDevice, e->stream));
    int rc = run_period(e, e->d_in, e->d_out);
    if (rc) return rc;
    CA_CUDA(cudaMemcpyAsync(dst, e->d_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
    CA_CUDA(cudaStreamSynchronize(e->stream));
    if (dst != out) memcpy(out, e->h_out, out_bytes);
    record_wall(e, now_us() - t0);
    return CA_OK;
}

int ca_sync(ca_engine *e)
{
    if (!e) return CA_ERR_INVALID;
    CA_CUDA(cudaStreamSynchronize(e->stream));
    return CA_OK;
}

void *ca_stream(ca_engine *e) { return e ? (void *)e->stream : nullptr; }

int ca_get_stats(ca_engine *e, ca_stats *s)
{
    if (!e || !s) return CA_ERR_INVALID;
    memset(s, 0, sizeof(*s));
    s->periods = e->periods; s->xruns = e->xruns;
    const size_t n = (size_t)std::min<uint64_t>(e->periods, e->wall.size());
    if (n) {
        std::vector<float> v(e->wall.begin(), e->wall.begin() + n);
        std::sort(v.begin(), v.end());
        s->p50_us = v[n / 2];
        s->p99_us = v[std::min(n - 1, (size_t)std::ceil(0.99 * (double)n))];
        s->max_us = e->wall_max;
        s->mean_us = e->wall_sum / (double)e->periods;
    }
    if (e->prof_n) {
        s->fwd_us = e->prof_us[0] / (double)e->prof_n;
        s->mac_us = e->prof_us[1] / (double)e->prof_n;
        s->inv_us = e->prof_us[2] / (double)e->prof_n;
        s->total_us = s->fwd_us + s->mac_us + s->inv_us;
    }
    s->gpu_launches = e->launches;
    // the MAC streams every IR partition spectrum and every FDL slot once
    s->mac_bytes = (uint64_t)8 * e->B * e->P * (uint64_t)(e->n_in * e->n_out + e->n_in) * e->n_active;
    s->partitions = e->P; s->mac_split = e->n_split; s->device_bytes = e->device_bytes;
    return CA_OK;
}

int ca_reset_stats(ca_engine *e)
{
    if (!e) return CA_ERR_INVALID;
    e->periods = e->xruns = 0; e->wall_sum = e->wall_max = 0;
    e->prof_us[0] = e->prof_us[1] = e->prof_us[2] = 0; e->prof_n = 0;
    return CA_OK;
}

int ca_host_alloc(void **p, size_t bytes)
{
    if (!p) return CA_ERR_INVALID;
    CA_CUDA(cudaMallocHost(p, bytes));
    return CA_OK;
}

int ca_host_free(void *p)
{
    CA_CUDA(cudaFreeHost(p));
    return CA_OK;
}

}  // extern "C"
