// fake_engine.cpp -- TEST DOUBLE of the C ABI (include/cuda_audio_b200.h) and of the handful of CUDA runtime calls the
// host mirror makes, so the mirror's THREADING (cuda-audio_b200/host/{convolution,shared_engine}.cpp: rendezvous of
// the engine.shared members, prepare() against a running callback, rebuilds, members that stop calling) runs on a
// machine without a GPU under ThreadSanitizer / AddressSanitizer.  It is not an engine: "processing" multiplies
// every input block by the first tap of the selected IR's left channel.  What it checks is the contract the header
// states for callers: ca_process / ca_load_ir / ca_reset / ca_destroy of one engine never overlap, nothing is called
// on a destroyed engine, instance / slot / input indices are in range, buffers are big enough (every byte is touched).
// Only tests/hostsim links it; the product never does.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/cuda_audio_b200.h"
#include "fake_engine.h"

namespace {
std::atomic<uint64_t> g_violations{0}, g_created{0}, g_destroyed{0}, g_processed{0}, g_loads{0};
std::atomic<int> g_processDelayUs{0}, g_createDelayUs{0};
std::atomic<int> g_failCreate{0};
std::atomic<uint64_t> g_onDevice[16], g_groups{0};

void violation(const char *what)
{
    g_violations.fetch_add(1);
    fprintf(stderr, "fake_engine: CONTRACT VIOLATION: %s\n", what);
}

constexpr uint32_t kMagicLive = 0xCA11AB1Eu, kMagicDead = 0xDEADC0DEu;
std::mutex g_graveyardMutex;
std::vector<ca_engine *> *g_graveyard = new std::vector<ca_engine *>();  // destroyed engines stay allocated (and reachable, also at exit): a late call finds kMagicDead
}  // namespace

struct ca_engine {
    std::atomic<uint32_t> magic{kMagicLive};
    ca_config cfg;
    std::atomic<int> exclusive{0};  // ca_process / ca_load_ir / ca_reset / ca_destroy in flight
    std::vector<std::atomic<float>> tap;      // [slot] first tap of the left IR; NaN = never loaded
    std::vector<std::atomic<uint32_t>> select;  // [instance * n_in + input]
    explicit ca_engine(const ca_config &c) : cfg(c), tap(c.n_ir_slots), select((size_t)c.n_instances * c.n_in)
    {
        for (auto &t : tap) t.store(__builtin_nanf(""));
        for (auto &s : select) s.store(0);
    }
};

namespace {
struct Exclusive {
    ca_engine *e;
    Exclusive(ca_engine *eng, const char *who) : e(eng)
    {
        if (e->exclusive.fetch_add(1) != 0) violation(who);
    }
    ~Exclusive() { e->exclusive.fetch_sub(1); }
};
bool alive(ca_engine *e, const char *who)
{
    if (!e) { violation(who); return false; }
    if (e->magic.load() != kMagicLive) { violation(who); return false; }
    return true;
}
}  // namespace

extern "C" {

uint64_t fake_violations(void) { return g_violations.load(); }
uint64_t fake_engines_created(void) { return g_created.load(); }
uint64_t fake_engines_destroyed(void) { return g_destroyed.load(); }
uint64_t fake_periods_processed(void) { return g_processed.load(); }
uint64_t fake_ir_loads(void) { return g_loads.load(); }
uint64_t fake_engines_on_device(int d) { return d >= 0 && d < 16 ? g_onDevice[d].load() : 0; }
uint64_t fake_groups_created(void) { return g_groups.load(); }
void fake_set_process_delay_us(int us) { g_processDelayUs.store(us); }
void fake_set_create_delay_us(int us) { g_createDelayUs.store(us); }
void fake_fail_next_creates(int n) { g_failCreate.store(n); }

int ca_api_version(void) { return CA_API_VERSION; }
const char *ca_strerror(int code) { return code == CA_OK ? "ok" : "error"; }
const char *ca_last_error_string(void) { return ""; }

void ca_config_init(ca_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->period = 256;
    cfg->n_instances = 1;
    cfg->n_in = cfg->n_out = 2;
    cfg->n_ir_slots = 1;
}

int ca_config_auto_tiers(ca_config *cfg, uint32_t, uint32_t)
{
    cfg->n_tiers = 1;
    cfg->tier_block[0] = cfg->period;
    return CA_OK;
}

int ca_create(const ca_config *cfg, ca_engine **out)
{
    if (!cfg || !out || cfg->struct_size != sizeof(ca_config) || !cfg->n_instances || !cfg->n_ir_slots || !cfg->period) { violation("ca_create: bad config"); return CA_ERR_INVALID; }
    if (int us = g_createDelayUs.load()) std::this_thread::sleep_for(std::chrono::microseconds(us));
    if (g_failCreate.load() > 0 && g_failCreate.fetch_sub(1) > 0) return CA_ERR_NOMEM;
    if (cfg->device < 0 || cfg->device >= 16) { violation("ca_create: device out of range"); return CA_ERR_INVALID; }
    *out = new ca_engine(*cfg);
    g_created.fetch_add(1);
    g_onDevice[cfg->device].fetch_add(1);
    return CA_OK;
}

int ca_destroy(ca_engine *e)
{
    if (!alive(e, "ca_destroy on a dead engine")) return CA_ERR_INVALID;
    {
        Exclusive x(e, "ca_destroy while another call is in flight");
        e->magic.store(kMagicDead);
    }
    g_destroyed.fetch_add(1);
    std::lock_guard<std::mutex> lk(g_graveyardMutex);
    g_graveyard->push_back(e);
    return CA_OK;
}

int ca_load_ir(ca_engine *e, uint32_t slot, const float *left, const float *right, uint32_t frames)
{
    if (!alive(e, "ca_load_ir on a dead engine")) return CA_ERR_INVALID;
    Exclusive x(e, "ca_load_ir overlaps another call");
    if (slot >= e->cfg.n_ir_slots || !left || !frames) { violation("ca_load_ir: bad slot / buffer"); return CA_ERR_INVALID; }
    if (frames > e->cfg.max_ir_frames) { violation("ca_load_ir: IR longer than max_ir_frames"); return CA_ERR_INVALID; }
    float sum = 0;  // touch every sample: ASan finds short buffers
    for (uint32_t i = 0; i < frames; i++) sum += left[i] + (right ? right[i] : 0.f);
    (void)sum;
    e->tap[slot].store(left[0]);
    g_loads.fetch_add(1);
    return CA_OK;
}

int ca_set_params(ca_engine *e, uint32_t instance, uint32_t input, const ca_params *p)
{
    if (!alive(e, "ca_set_params on a dead engine")) return CA_ERR_INVALID;
    if (instance >= e->cfg.n_instances || input >= e->cfg.n_in || !p) { violation("ca_set_params: index out of range"); return CA_ERR_INVALID; }
    if (p->select >= e->cfg.n_ir_slots) { violation("ca_set_params: select out of range"); return CA_ERR_STATE; }
    e->select[(size_t)instance * e->cfg.n_in + input].store(p->select);  // lock-free by contract: no Exclusive here
    return CA_OK;
}

int ca_reset(ca_engine *e)
{
    if (!alive(e, "ca_reset on a dead engine")) return CA_ERR_INVALID;
    Exclusive x(e, "ca_reset overlaps another call");
    return CA_OK;
}

int ca_process(ca_engine *e, const float *in, float *out, uint32_t nframes)
{
    if (!alive(e, "ca_process on a dead engine")) return CA_ERR_INVALID;
    Exclusive x(e, "ca_process overlaps another call");
    if (nframes != e->cfg.period) return CA_ERR_PERIOD;
    if (!in || !out) { violation("ca_process: null buffer"); return CA_ERR_INVALID; }
    if (int us = g_processDelayUs.load()) std::this_thread::sleep_for(std::chrono::microseconds(us));
    const uint32_t n_in = e->cfg.n_in;
    for (uint32_t k = 0; k < e->cfg.n_instances; k++)
        for (uint32_t c = 0; c < n_in; c++) {
            const uint32_t sel = e->select[(size_t)k * n_in + c].load();
            float g = e->tap[sel].load();
            if (g != g) g = 0.f;  // slot never loaded
            const float *x = in + ((size_t)k * n_in + c) * nframes;
            float *y = out + ((size_t)k * n_in + c) * nframes;
            for (uint32_t i = 0; i < nframes; i++) y[i] = g * x[i];
        }
    g_processed.fetch_add(1);
    return CA_OK;
}

// ---- ca_group (one IR over several GPUs): the double is one 1-instance engine that remembers the devices it was given ----
void ca_group_config_init(ca_group_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->n_devices = 1;
    cfg->period = 256;
    cfg->n_in = cfg->n_out = 2;
    cfg->n_ir_slots = 2;
}

int ca_group_create(const ca_group_config *gc, ca_group **out)
{
    if (!gc || !out || gc->struct_size != sizeof(*gc) || gc->n_devices < 1 || gc->n_devices > 8) { violation("ca_group_create: bad config"); return CA_ERR_INVALID; }
    for (uint32_t i = 0; i < gc->n_devices; i++)
        for (uint32_t j = 0; j < i; j++)
            if (gc->devices[i] == gc->devices[j]) { violation("ca_group_create: a device appears twice"); return CA_ERR_INVALID; }
    ca_config cfg;
    ca_config_init(&cfg);
    cfg.device = gc->devices[0];
    cfg.period = gc->period; cfg.n_in = gc->n_in; cfg.n_out = gc->n_out;
    cfg.max_ir_frames = gc->max_ir_frames; cfg.n_ir_slots = gc->n_ir_slots;
    ca_engine *e = nullptr;
    const int rc = ca_create(&cfg, &e);
    if (rc) return rc;
    for (uint32_t i = 1; i < gc->n_devices; i++) g_onDevice[gc->devices[i] & 15].fetch_add(1);
    g_groups.fetch_add(1);
    *out = reinterpret_cast<ca_group *>(e);
    return CA_OK;
}
int ca_group_destroy(ca_group *g) { return ca_destroy(reinterpret_cast<ca_engine *>(g)); }
int ca_group_load_ir(ca_group *g, uint32_t slot, const float *l, const float *r, uint32_t frames) { return ca_load_ir(reinterpret_cast<ca_engine *>(g), slot, l, r, frames); }
int ca_group_set_params(ca_group *g, uint32_t input, const ca_params *p) { return ca_set_params(reinterpret_cast<ca_engine *>(g), 0, input, p); }
int ca_group_reset(ca_group *g) { return ca_reset(reinterpret_cast<ca_engine *>(g)); }
int ca_group_process(ca_group *g, const float *in, float *out, uint32_t nframes) { return ca_process(reinterpret_cast<ca_engine *>(g), in, out, nframes); }

// "pinned" memory: plain heap, so AddressSanitizer sees every access after ca_host_free
int ca_host_alloc(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? CA_OK : CA_ERR_NOMEM; }
int ca_host_free(void *p) { free(p); return CA_OK; }

// ---- the CUDA runtime calls of host/wavfile.cpp and host/convolution.cpp: "device" memory is the heap ----
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { memcpy(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "fake"; }

}  // extern "C"
