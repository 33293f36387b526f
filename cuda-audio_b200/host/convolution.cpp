#include "convolution.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>

#include "settings_file.h"
#include "shared_engine.h"

// ---- engine options -------------------------------------------------------------------------
static bool truthy(const char *v) { return v && (v[0] == '1' || v[0] == 't' || v[0] == 'T' || v[0] == 'y' || v[0] == 'Y'); }

EngineOptions EngineOptions::fromEnv()
{
    EngineOptions o;
    if (const char *v = getenv("CA_ENGINE_DEVICE")) o.device = atoi(v);
    if (const char *v = getenv("CA_ENGINE_TIERS")) o.autoTiers = std::string(v) == "auto";
    if (const char *v = getenv("CA_ENGINE_PERIOD")) o.period = (uint32_t)atoi(v);
    if (const char *v = getenv("CA_ENGINE_SHARED")) o.shared = (uint32_t)atoi(v);
    if (const char *v = getenv("CA_ENGINE_SHARED_TIMEOUT_MS")) o.sharedTimeoutMs = (uint32_t)atoi(v);
    if (const char *v = getenv("CA_ENGINE_SHARED_LATENCY")) o.sharedLatency = (uint32_t)atoi(v) ? 1u : 0u;
    if (const char *v = getenv("CA_ENGINE_GPUS")) o.gpus = (uint32_t)std::max(1, atoi(v));
    if (const char *v = getenv("CA_ENGINE_IR_SPLIT")) o.irSplit = (uint32_t)atoi(v);
    if (const char *v = getenv("CA_ENGINE_EXCHANGE")) o.exchange = std::string(v) == "nccl" ? CA_EXCHANGE_NCCL : CA_EXCHANGE_P2P;
    if (o.irSplit) o.shared = 0;  // one object IS the whole group of GPUs
    if (truthy(getenv("CA_ENGINE_ASYNC_TIERS"))) o.flags |= CA_FLAG_ASYNC_TIERS;
    if (truthy(getenv("CA_ENGINE_REF_QUIRKS"))) o.flags |= CA_FLAG_REF_QUIRKS;
    if (o.shared > 1) o.flags = (o.flags & ~(uint32_t)CA_FLAG_GRAPH) | CA_FLAG_STREAMING;  // batches: host-driven launches + PDL
    return o;
}

EngineOptions EngineOptions::fromSettings(Settings &st)
{
    EngineOptions o = fromEnv();
    if (st.has("engine.device")) o.device = (int)st.u32("engine.device");
    if (st.has("engine.tiers")) o.autoTiers = st.str("engine.tiers") == "auto";
    if (st.has("engine.tier_growth")) o.tierGrowth = st.u32("engine.tier_growth");
    if (st.has("engine.tier_max_block")) o.tierMaxBlock = st.u32("engine.tier_max_block");
    if (st.has("engine.period")) o.period = st.u32("engine.period");
    if (st.has("engine.shared")) o.shared = st.u32("engine.shared");
    if (st.has("engine.shared_timeout_ms")) o.sharedTimeoutMs = st.u32("engine.shared_timeout_ms");
    if (st.has("engine.shared_latency")) o.sharedLatency = st.u32("engine.shared_latency") ? 1u : 0u;
    if (st.has("engine.gpus")) o.gpus = std::max<uint32_t>(1, st.u32("engine.gpus"));
    if (st.has("engine.ir_split")) o.irSplit = st.u32("engine.ir_split");
    if (st.has("engine.exchange")) o.exchange = st.str("engine.exchange") == "nccl" ? CA_EXCHANGE_NCCL : CA_EXCHANGE_P2P;
    if (o.irSplit) o.shared = 0;
    auto flag = [&](const char *key, uint32_t bit, bool dflt) {
        const bool on = st.has(key) ? truthy(st.str(key).c_str()) : dflt;
        o.flags = on ? (o.flags | bit) : (o.flags & ~bit);
    };
    flag("engine.graph", CA_FLAG_GRAPH, (o.flags & CA_FLAG_GRAPH) != 0);
    flag("engine.async_tiers", CA_FLAG_ASYNC_TIERS, (o.flags & CA_FLAG_ASYNC_TIERS) != 0);
    flag("engine.l2_persist", CA_FLAG_L2_PERSIST, (o.flags & CA_FLAG_L2_PERSIST) != 0);
    // the reference's DC / Nyquist bins (conv.cu:47-73) reproduced for its fftSize: bit-compatible sound on any IR
    flag("engine.ref_quirks", CA_FLAG_REF_QUIRKS, (o.flags & CA_FLAG_REF_QUIRKS) != 0);
    if (o.shared > 1) o.flags = (o.flags & ~(uint32_t)CA_FLAG_GRAPH) | CA_FLAG_STREAMING;
    return o;
}

static EngineOptions &defaultOptions()
{
    static EngineOptions o = EngineOptions::fromEnv();
    return o;
}

// engine.gpus: objects are dealt onto the GPUs in construction order (main.cu:31-39 constructs conv.count/2 of them)
static std::atomic<uint32_t> g_constructed{0};
static EngineOptions nextOptions()
{
    EngineOptions o = defaultOptions();
    if (o.gpus > 1) o.device += (int)((g_constructed.fetch_add(1) % o.gpus) * std::max<uint32_t>(1, o.irSplit));
    return o;
}

void Convolution::setDefaultOptions(const EngineOptions &o)
{
    defaultOptions() = o;
    g_constructed.store(0);
}

void Convolution::setOptions(const EngineOptions &o)
{
    if (built() || _shared) { fail(CA_ERR_STATE, "setOptions: the engine is already built"); return; }
    _opt = o;
    if (_opt.shared > 1) _shared = SharedEngine::join(this, _opt, &_sharedIdx);
}

Convolution::Convolution(const std::string &name, size_t fftSize) : JackClient(name), capture{nullptr, nullptr}, playback{nullptr, nullptr}, _fftSize(fftSize), _opt(nextOptions())
{
    for (auto &h : _hasIR) h.store(false, std::memory_order_relaxed);
    if (_opt.shared > 1) _shared = SharedEngine::join(this, _opt, &_sharedIdx);
}

Convolution::~Convolution()
{
    stop();  // no callback may be running or arrive from here on
    if (_shared) _shared->leave(this);
    destroyEngine();
    if (_in) ca_host_free(_in);
    if (_out) ca_host_free(_out);
}

void Convolution::fail(int code, const char *what)
{
    _lastError = code;
    _lastErrorText = std::string(what) + ": " + ca_strerror(code) + " (" + ca_last_error_string() + ")";
    Log::error(name, "%s", _lastErrorText.c_str());
}

// conv.cu:197-204.  The engine is built HERE (allocation, graph capture, IR transforms, warm-up), before the
// client is activated, so the first real-time callback finds it ready.
void Convolution::onStart()
{
    _stopping.store(false, std::memory_order_seq_cst);
    const size_t period = _opt.period ? _opt.period : (handle ? (size_t)jack_get_buffer_size(handle) : 0);
    // a shared batch is built by the member that completes it (earlier ones would build an engine that is rebuilt when the
    // next member joins); a group that never fills up is built at its first rendezvous
    if (period && numIRs() && (!_shared || _shared->full())) buildNow(period);
    activate();
    playback[0] = addOutput("playback_1");
    playback[1] = addOutput("playback_2");
    capture[0] = addInput("capture_1");
    capture[1] = addInput("capture_2");
}

// JackClient::stop(): no more callbacks for this object; the other members of a shared batch stop waiting for it
void Convolution::onStop()
{
    _stopping.store(true, std::memory_order_seq_cst);  // a callback that is still delivered must not join the batch again
    if (_shared) _shared->standDown(_sharedIdx);
}

bool Convolution::buildNow(size_t period)
{
    if (_shared) return _shared->buildNow(period, samplerate ? (float)samplerate : _sampleRate);
    std::lock_guard<std::mutex> lk(_engineMutex);
    if (built() && _period == period) return true;
    destroyEngine();
    if (!buildEngine(period)) return false;
    pushParams(true);
    // one silent period: first-launch costs (module load, graph upload) are paid here, not in the callback
    memset(_in, 0, 2 * period * sizeof(float));
    const int rc = processBlock((uint32_t)period);
    if (rc) { fail(rc, "ca_process (warm-up)"); return false; }
    resetEngine();  // the warm-up period must not count as the first step of the fade-in glide (conv.cu:15-32)
    return true;
}

// ---- the engine behind this object: a ca_engine of its own, or (engine.ir_split) a ca_group over several GPUs ----
void Convolution::destroyEngine()
{
    if (_engine) ca_destroy(_engine);
    if (_group) ca_group_destroy(_group);
    _engine = nullptr;
    _group = nullptr;
}

int Convolution::loadIR(size_t idx, const HostIR &ir)
{
    return _group ? ca_group_load_ir(_group, (uint32_t)idx, ir.left.data(), ir.right.data(), (uint32_t)ir.left.size())
                  : ca_load_ir(_engine, (uint32_t)idx, ir.left.data(), ir.right.data(), (uint32_t)ir.left.size());
}

int Convolution::processBlock(uint32_t nframes)
{
    return _group ? ca_group_process(_group, _in, _out, nframes) : ca_process(_engine, _in, _out, nframes);  // pinned staging: used in place
}

int Convolution::resetEngine() { return _group ? ca_group_reset(_group) : ca_reset(_engine); }

// conv.cu:207-253: the stereo IR `wav` (device float2 frames, half scale) becomes bank entry idx,
// truncated to fftSize - nframes frames.  Synchronous: wav may be destroyed on return.
void Convolution::prepare(size_t idx, const WavFile &wav, size_t nframes)
{
    if (!wav.buffer || !wav.numFrames) { fail(CA_ERR_INVALID, "prepare: empty wav"); return; }
    if (idx >= kMaxIRs) { fail(CA_ERR_INVALID, "prepare: IR index too large"); return; }
    const size_t cap = _fftSize > nframes ? _fftSize - nframes : 0;
    const size_t n = std::min(wav.numFrames, cap);
    if (!n) { fail(CA_ERR_INVALID, "prepare: fftSize too small"); return; }
    std::vector<float2> host(n);
    cudaSetDevice(_opt.device);
    if (cudaMemcpy(host.data(), wav.buffer, n * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) {
        (void)cudaGetLastError();
        fail(CA_ERR_CUDA, "prepare: cannot read the wav buffer");
        return;
    }
    HostIR ir;
    ir.left.resize(n);
    ir.right.resize(n);
    for (size_t i = 0; i < n; i++) { ir.left[i] = host[i].x; ir.right[i] = host[i].y; }
    auto store = [&] {  // under the lock that guards _irs
        _irs[idx] = std::move(ir);
        _hasIR[idx].store(true, std::memory_order_release);
        _firstIR.store(_irs.begin()->first, std::memory_order_release);
        _numIRs.store(_irs.size(), std::memory_order_release);
        _minPrepareFrames = std::min(_minPrepareFrames, nframes);
    };
    if (_shared) {
        std::lock_guard<std::mutex> lk(_shared->irMutex());  // a rebuild of the group only try_locks: never blocks its RT threads
        store();
        _shared->irChanged(this);
        return;
    }
    // A running engine: the load (or the rebuild) happens here, on the caller's thread, under the engine lock;
    // the real-time thread only try_locks and answers the periods it loses with silence (skippedPeriods()).
    // The reference's prepare() is not thread safe at all (conv.cu:206 "TODO make thread safe").
    std::lock_guard<std::mutex> lk(_engineMutex);
    store();
    const HostIR &cur = _irs[idx];
    if (built()) {
        if (idx < _engineSlots && n <= _engineCapFrames && loadIR(idx, cur) == CA_OK) return;
        const size_t period = _period;
        destroyEngine();
        if (buildEngine(period)) pushParams(true);
    } else if (_opt.period) {
        // period known up front (engine.period): build as soon as there is an IR, rebuild above when the bank grows
        if (buildEngine(_opt.period)) pushParams(true);
    }
}

bool Convolution::buildEngine(size_t period)
{
    if (_irs.empty()) { fail(CA_ERR_STATE, "onProcess: no IR prepared"); return false; }
    size_t longest = 1;
    for (auto &kv : _irs) longest = std::max(longest, kv.second.left.size());
    const uint32_t slots = (uint32_t)(_irs.rbegin()->first + 1);
    const float rate = samplerate ? (float)samplerate : _sampleRate;
    int rc;
    if (_opt.irSplit) {
        // one IR over several GPUs (BASELINE configs[4]): the reference's only answer to a long IR is a bigger fftSize
        // on one GPU (conv.cu:239, conv.h:63)
        ca_group_config gc;
        ca_group_config_init(&gc);
        gc.n_devices = std::min<uint32_t>(_opt.irSplit, 8);
        for (uint32_t g = 0; g < gc.n_devices; g++) gc.devices[g] = _opt.device + (int)g;
        gc.period = (uint32_t)period;
        gc.n_in = gc.n_out = 2;
        gc.max_ir_frames = (uint32_t)longest;
        gc.n_ir_slots = slots;
        gc.flags = _opt.flags & (uint32_t)(CA_FLAG_STREAMING | CA_FLAG_L2_PERSIST);
        gc.exchange = _opt.exchange;
        gc.max_voices = 3;
        gc.sample_rate = rate;
        rc = ca_group_create(&gc, &_group);
        if (rc) { _group = nullptr; fail(rc, "ca_group_create"); return false; }
    } else {
        ca_config cfg;
        ca_config_init(&cfg);
        cfg.device = _opt.device;
        cfg.period = (uint32_t)period;
        cfg.n_instances = 1;
        cfg.n_in = cfg.n_out = 2;
        cfg.max_ir_frames = (uint32_t)longest;
        cfg.n_ir_slots = slots;
        cfg.flags = _opt.flags;
        if (cfg.flags & CA_FLAG_REF_QUIRKS) {
            if (_fftSize % period == 0 && _fftSize >= 2 * period) cfg.ref_fft_size = (uint32_t)_fftSize;  // Convolution(name, fftSize), conv.cu:142
            else cfg.flags &= ~(uint32_t)CA_FLAG_REF_QUIRKS;  // the reference itself only works for such sizes
        }
        cfg.max_voices = 3;  // old IR + new IR + one more switch in flight during a cross-fade
        cfg.sample_rate = rate;
        if (_opt.autoTiers && ca_config_auto_tiers(&cfg, _opt.tierGrowth, _opt.tierMaxBlock) != CA_OK) { fail(CA_ERR_INVALID, "ca_config_auto_tiers"); return false; }
        rc = ca_create(&cfg, &_engine);
        if (rc) { _engine = nullptr; fail(rc, "ca_create"); return false; }
    }
    for (auto &kv : _irs) {
        rc = loadIR(kv.first, kv.second);
        if (rc) { fail(rc, "ca_load_ir"); destroyEngine(); return false; }
    }
    _period = period;
    _engineSlots = slots;
    _engineCapFrames = longest;
    if (_in) ca_host_free(_in);
    if (_out) ca_host_free(_out);
    _in = _out = nullptr;
    if (ca_host_alloc((void **)&_in, 2 * period * sizeof(float)) || ca_host_alloc((void **)&_out, 2 * period * sizeof(float))) { fail(CA_ERR_NOMEM, "ca_host_alloc"); destroyEngine(); return false; }
    memset(_in, 0, 2 * period * sizeof(float));
    memset(_out, 0, 2 * period * sizeof(float));
    _havePushed = false;
    return true;
}

// cc[i].value is plain shared state like in the reference (conv.cu:339-427 reads it every period while the MIDI thread
// and user code write it whenever they like): whatever changed since the last period is forwarded to the engine.
// These two accessors are the only places the engine side touches it; they are kept out of ThreadSanitizer's view
// because that race is the interface, not a defect.
#if defined(__GNUC__) || defined(__clang__)
#define CA_PLAIN_SHARED __attribute__((noinline, no_sanitize("thread")))
#else
#define CA_PLAIN_SHARED
#endif
CA_PLAIN_SHARED static Convolution::CC::Value readValue(const Convolution::CC::Value &v) { return v; }
CA_PLAIN_SHARED static void writeBack(Convolution::CC::Value &v, bool fixSelect, size_t select, bool step)
{
    if (fixSelect) v.select = select;
    if (step && v.vsteps > 0) v.vsteps--;  // conv.cu:345,353: one step per period
}

void Convolution::pushParams(bool force) { pushParamsTo(_engine, 0, 0, force); }

// instance / slotBase: position of this object inside a shared batched engine (0, 0 for its own engine)
void Convolution::pushParamsTo(ca_engine *e, uint32_t instance, size_t slotBase, bool force)
{
    for (int i = 0; i < 2; i++) {
        CC::Value v = readValue(cc[i].value);
        CC::Value &p = _pushed[i];
        // the engine counts the glide down on the device; mirror it so `vsteps` stays observable
        const size_t expected = _havePushed ? p.vsteps : (size_t)-1;
        const bool vstepsChanged = v.vsteps != expected;
        const bool changed = force || !_havePushed || vstepsChanged || v.select != p.select || v.predelay != p.predelay || v.dry != p.dry ||
                             v.wet != p.wet || v.panDry != p.panDry || v.panWet != p.panWet || v.level != p.level || v.speed != p.speed;
        bool fixSelect = false;
        if (changed) {
            if (v.select >= kMaxIRs || !_hasIR[v.select].load(std::memory_order_acquire)) {
                const size_t keep = _havePushed ? p.select : _firstIR.load(std::memory_order_acquire);
                Log::error(name, "select %zu has no IR; keeping %zu", v.select, keep);  // reference: nullptr deref (conv.cu:340)
                v.select = keep;
                fixSelect = true;
            }
            ca_params q;
            q.select = (uint32_t)(slotBase + v.select);
            q.predelay = (uint32_t)std::min<size_t>(v.predelay, CA_MAX_PREDELAY - 1);
            q.speed = (uint32_t)v.speed;
            q.vsteps = (vstepsChanged || !_havePushed) ? (int32_t)v.vsteps : -1;
            q.dry = v.dry; q.wet = v.wet; q.panDry = v.panDry; q.panWet = v.panWet; q.level = v.level;
            const int rc = (_group && e == nullptr) ? ca_group_set_params(_group, (uint32_t)i, &q) : ca_set_params(e, instance, (uint32_t)i, &q);
            if (rc) fail(rc, "ca_set_params");
            p = v;
        }
        writeBack(cc[i].value, fixSelect, v.select, true);
        if (v.vsteps > 0) v.vsteps--;
        p.vsteps = v.vsteps;
    }
    _havePushed = true;
}

// conv.cu:287-466
void Convolution::onProcess(size_t nframes)
{
    auto *IN1 = capture[0] ? (float *)jack_port_get_buffer(capture[0], (jack_nframes_t)nframes) : nullptr;
    auto *IN2 = capture[1] ? (float *)jack_port_get_buffer(capture[1], (jack_nframes_t)nframes) : nullptr;
    auto *L = playback[0] ? (float *)jack_port_get_buffer(playback[0], (jack_nframes_t)nframes) : nullptr;
    auto *R = playback[1] ? (float *)jack_port_get_buffer(playback[1], (jack_nframes_t)nframes) : nullptr;
    if (!IN1 || !IN2 || !L || !R) return;  // conv.cu:297

    const auto t0 = std::chrono::steady_clock::now();
    auto silence = [&] { memset(L, 0, nframes * sizeof(float)); memset(R, 0, nframes * sizeof(float)); };
    if (_shared) {
        if (_stopping.load(std::memory_order_seq_cst) || !_shared->process(this, _sharedIdx, IN1, IN2, L, R, nframes)) { silence(); _skipped++; }
    } else {
        std::unique_lock<std::mutex> lk(_engineMutex, std::try_to_lock);
        if (!lk.owns_lock()) { silence(); _skipped++; return; }  // prepare() is loading an IR: never block the RT thread
        if (!built() || _period != nframes) {
            // no buildNow() / onStart() with a known buffer size came first (plain harness): build here
            destroyEngine();
            if (!buildEngine(nframes)) { silence(); return; }
        }
        pushParams(false);
        memcpy(_in, IN1, nframes * sizeof(float));
        memcpy(_in + nframes, IN2, nframes * sizeof(float));
        const int rc = processBlock((uint32_t)nframes);
        if (rc) { fail(rc, "ca_process"); silence(); return; }
        memcpy(L, _out, nframes * sizeof(float));
        memcpy(R, _out + nframes, nframes * sizeof(float));
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (++_nruns > 0) _runtimeMs += ms;  // conv.cu:462
}

// handleCC, conv.cu:255-276
static void handleCC(Convolution::CC &cc, uint8_t m1, uint8_t m2, int v, size_t nb)
{
    if (cc.message != m1) return;
    if (cc.select == m2) { cc.value.select = (size_t)v * nb / 0x80; cc.value.vsteps = cc.value.speed; Log::info("conv", "Selected IR %zu", cc.value.select); }
    if (cc.predelay == m2) cc.value.predelay = (size_t)v * CONV_MAX_PREDELAY / 0x80;
    if (cc.dry == m2) cc.value.dry = v / 128.0f;
    if (cc.wet == m2) cc.value.wet = v / 128.0f;
    if (cc.panDry == m2) cc.value.panDry = v / 64.0f - 1;
    if (cc.panWet == m2) cc.value.panWet = v / 64.0f - 1;
    if (cc.level == m2) cc.value.level = v / 128.0f;
    if (cc.speed == m2) {
        cc.value.speed = ((size_t)v * CONV_MAX_SPEED) / 0x80;
        if (cc.value.vsteps > cc.value.speed) cc.value.vsteps = cc.value.speed;
    }
}

// conv.cu:278-285
void Convolution::onMidiMessage(const RawMidi::Device *sender, const uint8_t *buffer, size_t len)
{
    if (len < 3) return;
    for (auto &c : cc)
        if (c.device == sender) handleCC(c, buffer[0], buffer[1], buffer[2], numIRs());
}
