#include "shared_engine.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>

namespace {
std::mutex g_registryMutex;
std::shared_ptr<SharedEngine> g_open;  // the group that still has room
}  // namespace

std::shared_ptr<SharedEngine> SharedEngine::join(Convolution *c, const EngineOptions &opt, int *index)
{
    std::lock_guard<std::mutex> lk(g_registryMutex);
    if (!g_open || g_open->_members.size() >= g_open->_opt.shared || g_open->_opt.shared != opt.shared)
        g_open = std::shared_ptr<SharedEngine>(new SharedEngine(opt));
    *index = (int)g_open->_members.size();
    g_open->_members.push_back(c);
    g_open->_dirty.store(true);
    return g_open;
}

SharedEngine::~SharedEngine()
{
    if (_engine) ca_destroy(_engine);
    if (_in) ca_host_free(_in);
    if (_out) ca_host_free(_out);
}

void SharedEngine::leave(Convolution *c)
{
    // the slot stays (indices of the others must not move); the batch simply stops waiting for it
    std::lock_guard<std::mutex> lk(_buildMutex);
    for (auto &m : _members)
        if (m == c) m = nullptr;
}

void SharedEngine::irChanged(Convolution *) { _dirty.store(true, std::memory_order_release); }

bool SharedEngine::buildNow(size_t period, float sampleRate)
{
    std::lock_guard<std::mutex> lk(_buildMutex);
    if (_engine && _period == period && !_dirty.load(std::memory_order_acquire)) return true;
    return build(period, sampleRate);
}

bool SharedEngine::build(size_t period, float sampleRate)
{
    if (_engine) { ca_destroy(_engine); _engine = nullptr; }
    _ok.store(false);
    size_t slots = 1, longest = 1;
    for (auto *m : _members) {
        if (!m || m->_irs.empty()) continue;
        slots = std::max(slots, m->_irs.rbegin()->first + 1);
        for (auto &kv : m->_irs) longest = std::max(longest, kv.second.left.size());
    }
    const uint32_t n = (uint32_t)_members.size();
    ca_config cfg;
    ca_config_init(&cfg);
    cfg.device = _opt.device;
    cfg.period = (uint32_t)period;
    cfg.n_instances = n;
    cfg.n_in = cfg.n_out = 2;
    cfg.max_ir_frames = (uint32_t)longest;
    cfg.n_ir_slots = (uint32_t)(slots * n);
    cfg.flags = _opt.flags;
    cfg.max_voices = 3;
    cfg.sample_rate = sampleRate;
    if (_opt.autoTiers && ca_config_auto_tiers(&cfg, _opt.tierGrowth, _opt.tierMaxBlock) != CA_OK) return false;
    if (ca_create(&cfg, &_engine) != CA_OK) { _engine = nullptr; return false; }
    _slotsPerMember = slots;
    _period = period;
    for (uint32_t i = 0; i < n; i++) {
        Convolution *m = _members[i];
        if (!m) continue;
        for (auto &kv : m->_irs)
            if (ca_load_ir(_engine, (uint32_t)(i * slots + kv.first), kv.second.left.data(), kv.second.right.data(), (uint32_t)kv.second.left.size()) != CA_OK) {
                ca_destroy(_engine); _engine = nullptr; return false;
            }
        m->_havePushed = false;
        if (!m->_irs.empty()) m->pushParamsTo(_engine, i, i * slots, true);
    }
    if (_in) ca_host_free(_in);
    if (_out) ca_host_free(_out);
    _in = _out = nullptr;
    const size_t bytes = (size_t)n * 2 * period * sizeof(float);
    if (ca_host_alloc((void **)&_in, bytes) || ca_host_alloc((void **)&_out, bytes)) { ca_destroy(_engine); _engine = nullptr; return false; }
    memset(_in, 0, bytes);
    memset(_out, 0, bytes);
    if (ca_process(_engine, _in, _out, (uint32_t)period) != CA_OK) { ca_destroy(_engine); _engine = nullptr; return false; }  // warm-up
    ca_reset(_engine);  // ... which must not count as the first step of the fade-in glide
    _dirty.store(false, std::memory_order_release);
    _ok.store(true, std::memory_order_release);
    return true;
}

bool SharedEngine::process(Convolution *c, int idx, const float *in1, const float *in2, float *L, float *R, size_t nframes)
{
    const uint64_t gen = _generation.load(std::memory_order_acquire);
    const bool usable = _ok.load(std::memory_order_acquire) && _period == nframes && !_dirty.load(std::memory_order_acquire);
    if (usable) {
        float *dst = _in + (size_t)idx * 2 * nframes;
        memcpy(dst, in1, nframes * sizeof(float));
        memcpy(dst + nframes, in2, nframes * sizeof(float));
        c->pushParamsTo(_engine, (uint32_t)idx, (size_t)idx * _slotsPerMember, false);  // lock-free hand-off: any thread
    }
    int live = 0;
    for (auto *m : _members) live += m ? 1 : 0;
    if (_arrived.fetch_add(1, std::memory_order_acq_rel) + 1 >= live) {
        // last to arrive: (re)build if needed, run the whole batch, release the others
        _arrived.store(0, std::memory_order_relaxed);
        bool ok = usable;
        if (!usable) {
            std::lock_guard<std::mutex> lk(_buildMutex);
            ok = build(nframes, c->samplerate ? (float)c->samplerate : c->_sampleRate);
            if (ok) {  // this cycle's inputs were not staged: it is answered with silence, the next one runs
                _generation.store(gen + 1, std::memory_order_release);
                return false;
            }
        }
        if (ok) ok = ca_process(_engine, _in, _out, (uint32_t)nframes) == CA_OK;
        _ok.store(ok || !usable ? _ok.load() : false);
        _generation.store(gen + 1, std::memory_order_release);
        if (!ok) return false;
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        for (int spins = 0; _generation.load(std::memory_order_acquire) == gen; spins++) {
            if (spins > 2000) {
                std::this_thread::yield();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) return false;  // the others never came
            }
        }
        if (!usable) return false;
    }
    const float *src = _out + (size_t)idx * 2 * nframes;
    memcpy(L, src, nframes * sizeof(float));
    memcpy(R, src + nframes, nframes * sizeof(float));
    return true;
}
