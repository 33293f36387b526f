#include "shared_engine.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <thread>

namespace {
std::mutex g_registryMutex;
std::map<int, std::shared_ptr<SharedEngine>> g_openByDevice;  // per GPU (engine.gpus): the group that still has room

constexpr uint64_t kCountMask = 0xffff;
inline uint64_t genOf(uint64_t s) { return s >> 16; }
inline int countOf(uint64_t s) { return (int)(s & kCountMask); }

struct Inside {  // a member is inside process() for the lifetime of this object
    std::atomic<int> &n;
    explicit Inside(std::atomic<int> &c) : n(c) { n.fetch_add(1, std::memory_order_seq_cst); }
    ~Inside() { n.fetch_sub(1, std::memory_order_seq_cst); }
};
}  // namespace

SharedEngine::SharedEngine(const EngineOptions &opt) : _opt(opt)
{
    const size_t cap = std::max<size_t>(opt.shared, 1);
    _members.reserve(cap);
    _active.reset(new std::atomic<bool>[cap]);
    _arrivedGen.reset(new std::atomic<uint64_t>[cap]);
    _stagedGen.reset(new std::atomic<uint64_t>[cap]);
    for (size_t i = 0; i < cap; i++) { _active[i].store(false); _arrivedGen[i].store(0); _stagedGen[i].store(0); }
}

std::shared_ptr<SharedEngine> SharedEngine::join(Convolution *c, const EngineOptions &opt, int *index)
{
    std::lock_guard<std::mutex> lk(g_registryMutex);
    std::shared_ptr<SharedEngine> &g_open = g_openByDevice[opt.device];
    if (!g_open || g_open->_members.size() >= g_open->_opt.shared || g_open->_opt.shared != opt.shared)
        g_open = std::shared_ptr<SharedEngine>(new SharedEngine(opt));
    std::lock_guard<std::mutex> bl(g_open->_buildMutex);  // a rebuild at the rendezvous walks _members
    *index = (int)g_open->_members.size();
    g_open->_members.push_back(c);
    g_open->_dirty.store(true, std::memory_order_release);  // the engine has one instance too few
    // Expected at the rendezvous from the start: a host that calls all members every cycle (jackd, a harness with a
    // barrier) finds the whole batch waiting for the last one in the very first cycle.  Members that are activated
    // later are set aside after sharedTimeoutMs, once, and take part when their callbacks begin.
    g_open->_active[*index].store(true, std::memory_order_release);
    g_open->_live.fetch_add(1, std::memory_order_seq_cst);
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
    for (size_t i = 0; i < _members.size(); i++) {
        if (_members[i] != c) continue;
        _members[i] = nullptr;
        bool was = true;
        if (_active[i].compare_exchange_strong(was, false)) _live.fetch_sub(1, std::memory_order_seq_cst);
    }
}

bool SharedEngine::full()
{
    std::lock_guard<std::mutex> lk(_buildMutex);
    return _members.size() >= std::max<size_t>(_opt.shared, 1);
}

void SharedEngine::standDown(int idx)
{
    if (idx < 0 || (size_t)idx >= std::max<size_t>(_opt.shared, 1)) return;
    bool was = true;
    if (_active[idx].compare_exchange_strong(was, false)) _live.fetch_sub(1, std::memory_order_seq_cst);
}

void SharedEngine::irChanged(Convolution *) { _dirty.store(true, std::memory_order_release); }

bool SharedEngine::buildNow(size_t period, float sampleRate)
{
    std::lock_guard<std::mutex> lk(_buildMutex);
    if (_engine && _ok.load(std::memory_order_acquire) && _period == period && !_dirty.load(std::memory_order_acquire)) return true;
    // Nobody may be inside process() while the engine and its buffers are replaced: new arrivals see _exclusive
    // and answer with silence, waiting members leave the rendezvous, a running batch finishes.
    _exclusive.fetch_add(1, std::memory_order_seq_cst);
    _evict.store(true, std::memory_order_seq_cst);
    while (_inside.load(std::memory_order_seq_cst) != 0) std::this_thread::yield();
    const bool ok = build(period, sampleRate);
    _state.store((genOf(_state.load(std::memory_order_relaxed)) + 1) << 16, std::memory_order_release);  // abandoned cycle, if any
    _evict.store(false, std::memory_order_seq_cst);
    _exclusive.fetch_sub(1, std::memory_order_seq_cst);
    return ok;
}

bool SharedEngine::build(size_t period, float sampleRate)
{
    _rebuilds.fetch_add(1, std::memory_order_relaxed);
    _ok.store(false, std::memory_order_release);
    if (_engine) { ca_destroy(_engine); _engine = nullptr; }
    size_t slots = 1, longest = 1;
    for (auto *m : _members) {
        if (!m || m->_irs.empty()) continue;
        slots = std::max(slots, m->_irs.rbegin()->first + 1);
        for (auto &kv : m->_irs) longest = std::max(longest, kv.second.left.size());
    }
    const uint32_t n = (uint32_t)_members.size();
    // from here on a failure leaves _dirty set: the next rendezvous (or buildNow) tries again
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
    _builtMembers = n;
    // The IR set is sampled now: a prepare() that lands while the loads below run raises _dirty again afterwards.
    _dirty.store(false, std::memory_order_release);
    auto fail = [&] { ca_destroy(_engine); _engine = nullptr; _dirty.store(true, std::memory_order_release); return false; };
    for (uint32_t i = 0; i < n; i++) {
        Convolution *m = _members[i];
        if (!m) continue;
        for (auto &kv : m->_irs)
            if (ca_load_ir(_engine, (uint32_t)(i * slots + kv.first), kv.second.left.data(), kv.second.right.data(), (uint32_t)kv.second.left.size()) != CA_OK) return fail();
        m->_havePushed = false;
        if (!m->_irs.empty()) m->pushParamsTo(_engine, i, i * slots, true);
    }
    if (_in) ca_host_free(_in);
    if (_out) ca_host_free(_out);
    _in = _out = nullptr;
    const size_t bytes = (size_t)n * 2 * period * sizeof(float) * (_opt.sharedLatency ? 2 : 1);  // pipelined: one buffer per generation parity
    if (ca_host_alloc((void **)&_in, bytes) || ca_host_alloc((void **)&_out, bytes)) return fail();
    memset(_in, 0, bytes);
    memset(_out, 0, bytes);
    if (ca_process(_engine, _in, _out, (uint32_t)period) != CA_OK) return fail();  // warm-up
    ca_reset(_engine);  // ... which must not count as the first step of the fade-in glide
    _ok.store(true, std::memory_order_release);
    return true;
}

// A member that has not arrived for sharedTimeoutMs (its JACK client was stopped, or never activated) is set
// aside so the others keep their deadline; it takes part again with its next call.
void SharedEngine::dropStalled(uint64_t gen)
{
    const size_t cap = std::max<size_t>(_opt.shared, 1);
    for (size_t i = 0; i < cap; i++) {
        if (!_active[i].load(std::memory_order_acquire) || _arrivedGen[i].load(std::memory_order_acquire) == gen + 1) continue;
        bool was = true;
        if (_active[i].compare_exchange_strong(was, false)) {
            _live.fetch_sub(1, std::memory_order_seq_cst);
            _dropped.fetch_add(1, std::memory_order_relaxed);
        }
    }
}

// On the thread of the member that found everybody arrived.  Ends generation `gen`.
void SharedEngine::runBatch(Convolution *c, size_t nframes, uint64_t gen)
{
    bool ok = false;
    if (_ok.load(std::memory_order_acquire) && _period == nframes && !_dirty.load(std::memory_order_acquire)) {
        const size_t half = _opt.sharedLatency ? _builtMembers * 2 * nframes * (gen & 1) : 0;  // pipelined: buffers by generation parity
        ok = ca_process(_engine, _in + half, _out + half, (uint32_t)nframes) == CA_OK;
        _batches.fetch_add(1, std::memory_order_relaxed);
        if (!ok) { _ok.store(false, std::memory_order_release); _dirty.store(true, std::memory_order_release); }
    } else {
        // (Re)build here, on a real-time thread: prepare() on a live group, a period change, or no buildNow() before
        // the first cycle.  This cycle is answered with silence by everybody (its inputs were not staged), and so are
        // the cycles that arrive while the build runs.  Anybody inside process() who is not waiting at this rendezvous
        // (between entry and arrival, or a member set aside as stalled that wakes up now) may still touch the
        // buffers: then the rebuild waits for the next cycle.
        _exclusive.fetch_add(1, std::memory_order_seq_cst);
        // (pipelined members touch the buffers inside the staging section only, and are not inside between their calls)
        if (_staging.load(std::memory_order_seq_cst) == 0 && (_opt.sharedLatency || _inside.load(std::memory_order_seq_cst) == countOf(_state.load(std::memory_order_seq_cst)))) {
            std::unique_lock<std::mutex> lk(_buildMutex, std::try_to_lock);
            if (lk.owns_lock()) build(nframes, c->samplerate ? (float)c->samplerate : c->_sampleRate);
        }
        _exclusive.fetch_sub(1, std::memory_order_seq_cst);
    }
    if (ok) _okGen.store(gen + 1, std::memory_order_release);
    _state.store((gen + 1) << 16, std::memory_order_release);
}

// Has every member that takes part handed in its block for generation `gen`?  The rendezvous counts arrivals (its
// members stay inside process() until the generation ends, so the count cannot go stale).  A pipelined member is gone
// as soon as it has arrived and may stand down, leave or be set aside before the generation ends: there the answer is
// taken from the members' own marks.
bool SharedEngine::everybodyArrived(uint64_t s, uint64_t gen) const
{
    if (genOf(s) != gen) return false;
    if (!_opt.sharedLatency) return countOf(s) >= _live.load(std::memory_order_seq_cst);
    const size_t cap = std::max<size_t>(_opt.shared, 1);
    bool any = false;
    for (size_t i = 0; i < cap; i++) {
        if (!_active[i].load(std::memory_order_seq_cst)) continue;
        if (_arrivedGen[i].load(std::memory_order_acquire) != gen + 1) return false;
        any = true;
    }
    return any;
}

// Wait until generation `gen` has ended; whoever finds everybody arrived runs its batch.  false: buildNow() on another
// thread wants everybody out (this cycle is silence).
bool SharedEngine::waitGeneration(Convolution *c, size_t nframes, uint64_t gen)
{
    auto t0 = std::chrono::steady_clock::now();
    for (int spins = 0;; spins++) {
        uint64_t s = _state.load(std::memory_order_acquire);
        if (genOf(s) != gen) return true;  // the batch ran
        if (everybodyArrived(s, gen)) {
            bool expected = false;
            if (_runner.compare_exchange_strong(expected, true, std::memory_order_seq_cst)) {
                if (everybodyArrived(_state.load(std::memory_order_seq_cst), gen)) runBatch(c, nframes, gen);
                _runner.store(false, std::memory_order_seq_cst);
                continue;
            }
        }
        // buildNow() on another thread waits for everybody to leave: this cycle is silence.  (A rebuild by this
        // cycle's runner is waited for instead: a driver that is not paced by a clock -- an offline render, one
        // loop per member -- would otherwise run through its input while the build lasts.)
        if (_evict.load(std::memory_order_seq_cst)) return false;
        if (spins > 2000) {
            std::this_thread::yield();
            if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(_opt.sharedTimeoutMs ? _opt.sharedTimeoutMs : 200)) {
                dropStalled(gen);
                t0 = std::chrono::steady_clock::now();
            }
        }
    }
}

// Entry of a member into the section in which it may touch the batch's buffers.  false: a build is in progress.
bool SharedEngine::enterStaging(int idx)
{
    _staging.fetch_add(1, std::memory_order_seq_cst);
    if (_exclusive.load(std::memory_order_seq_cst) != 0) {  // never block, never touch the buffers
        _staging.fetch_sub(1, std::memory_order_seq_cst);
        return false;
    }
    if (!_active[idx].load(std::memory_order_acquire)) {  // back after having been set aside (stopped client, stall)
        _live.fetch_add(1, std::memory_order_seq_cst);
        _active[idx].store(true, std::memory_order_seq_cst);
        // A runner that decided before it could see us is reading the input buffer: stage only after its batch.
        // (It cannot be rebuilding: _staging != 0.  Steady-state members never wait here.)
        while (_runner.load(std::memory_order_seq_cst)) std::this_thread::yield();
    }
    return true;
}

bool SharedEngine::process(Convolution *c, int idx, const float *in1, const float *in2, float *L, float *R, size_t nframes)
{
    if (idx < 0 || (size_t)idx >= std::max<size_t>(_opt.shared, 1)) return false;
    Inside inside(_inside);
    if (_opt.sharedLatency) return processPipelined(c, idx, in1, in2, L, R, nframes);
    if (!enterStaging(idx)) return false;
    uint64_t s = _state.load(std::memory_order_acquire);
    const uint64_t gen = genOf(s);
    const bool usable = _ok.load(std::memory_order_acquire) && !_dirty.load(std::memory_order_acquire) && _period == nframes && (size_t)idx < _builtMembers;
    if (usable) {
        float *dst = _in + (size_t)idx * 2 * nframes;
        memcpy(dst, in1, nframes * sizeof(float));
        memcpy(dst + nframes, in2, nframes * sizeof(float));
        c->pushParamsTo(_engine, (uint32_t)idx, (size_t)idx * _slotsPerMember, false);  // lock-free hand-off: any thread
    }
    // arrive -- unless this generation's batch already ran without us (we had been set aside, or we joined while the
    // runner was deciding): then this period is silence and the next one is in step again
    // (the buffers are not touched again before the generation ends: leave the staging section BEFORE arriving, or a
    // runner that sees the last arrival could still find _staging != 0 and put a rebuild off for no reason)
    _staging.fetch_sub(1, std::memory_order_seq_cst);
    bool counted = false;
    while (genOf(s) == gen) {
        if (_state.compare_exchange_weak(s, s + 1, std::memory_order_acq_rel, std::memory_order_acquire)) { counted = true; break; }
    }
    if (!counted) return false;
    _arrivedGen[idx].store(gen + 1, std::memory_order_release);
    if (!waitGeneration(c, nframes, gen)) return false;
    if (!usable || _okGen.load(std::memory_order_acquire) != gen + 1) return false;
    const float *src = _out + (size_t)idx * 2 * nframes;
    memcpy(L, src, nframes * sizeof(float));
    memcpy(R, src + nframes, nframes * sizeof(float));
    return true;
}

// engine.shared_latency 1: nobody waits for anybody inside a cycle.  A member hands in its block for generation g, takes
// the output of generation g - 1 (one period of extra latency for every member alike) and returns; the member that
// completes the set runs the batch before it returns.  For hosts that call their clients ONE AFTER THE OTHER on one
// thread (jack1, a plain loop), where a rendezvous inside the callback can never be met.  Input and output are double
// buffered by generation parity.  A member that comes back before the generation it already handed a block to has
// ended (a driver without a clock running ahead; a member of the set stalled) waits for that end first, with the same
// timeout / set-aside rules as the rendezvous.
bool SharedEngine::processPipelined(Convolution *c, int idx, const float *in1, const float *in2, float *L, float *R, size_t nframes)
{
    uint64_t gen = genOf(_state.load(std::memory_order_acquire));
    if (_active[idx].load(std::memory_order_acquire) && _arrivedGen[idx].load(std::memory_order_acquire) == gen + 1) {
        if (!waitGeneration(c, nframes, gen)) return false;
    }
    if (!enterStaging(idx)) return false;
    uint64_t s = _state.load(std::memory_order_acquire);
    gen = genOf(s);
    const bool usable = _ok.load(std::memory_order_acquire) && !_dirty.load(std::memory_order_acquire) && _period == nframes && (size_t)idx < _builtMembers;
    const size_t half = _builtMembers * 2 * nframes;  // floats of one generation's buffer
    // the block this member handed in one generation ago, processed by that generation's batch
    const bool have = usable && gen > 0 && _stagedGen[idx].load(std::memory_order_relaxed) == gen && _okGen.load(std::memory_order_acquire) == gen;
    if (have) {
        const float *src = _out + ((gen - 1) & 1) * half + (size_t)idx * 2 * nframes;
        memcpy(L, src, nframes * sizeof(float));
        memcpy(R, src + nframes, nframes * sizeof(float));
    }
    if (usable) {
        float *dst = _in + (gen & 1) * half + (size_t)idx * 2 * nframes;
        memcpy(dst, in1, nframes * sizeof(float));
        memcpy(dst + nframes, in2, nframes * sizeof(float));
        c->pushParamsTo(_engine, (uint32_t)idx, (size_t)idx * _slotsPerMember, false);
        _stagedGen[idx].store(gen + 1, std::memory_order_relaxed);
    }
    _staging.fetch_sub(1, std::memory_order_seq_cst);
    bool counted = false;
    while (genOf(s) == gen) {
        if (_state.compare_exchange_weak(s, s + 1, std::memory_order_acq_rel, std::memory_order_acquire)) { counted = true; break; }
    }
    if (!counted) return have;
    _arrivedGen[idx].store(gen + 1, std::memory_order_release);
    // the member that completes the set runs the batch now, on its own thread; everybody else is already gone
    if (everybodyArrived(_state.load(std::memory_order_seq_cst), gen)) {
        bool expected = false;
        if (_runner.compare_exchange_strong(expected, true, std::memory_order_seq_cst)) {
            if (everybodyArrived(_state.load(std::memory_order_seq_cst), gen)) runBatch(c, nframes, gen);
            _runner.store(false, std::memory_order_seq_cst);
        }
    }
    return have;
}
