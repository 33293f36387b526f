// shared_engine.h -- `engine.shared N`: N Convolution objects (the conv.count/2 JACK clients main.cu:31-39
// creates, one real-time callback thread each) are the N instances of ONE batched ca_engine, so a JACK cycle
// costs one set of kernel launches for all of them instead of N.  Every object keeps the reference's
// surface (prepare / onProcess / cc[] / ports); its onProcess() copies the period's input into the batch's
// pinned buffer and meets the others at a rendezvous; the last one to arrive runs ca_process() for the
// whole batch, everybody copies its own output block out.
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <vector>

#include "convolution.h"

class SharedEngine {
public:
    // process-wide: objects are dealt into groups of opt.shared members in construction order
    static std::shared_ptr<SharedEngine> join(Convolution *c, const EngineOptions &opt, int *index);
    ~SharedEngine();

    void leave(Convolution *c);
    void irChanged(Convolution *c);  // a member (re)prepared an IR: rebuild at the next rendezvous / buildNow
    bool buildNow(size_t period, float sampleRate);
    // one period of member `idx`; false: silence (build failure, or the others never arrived)
    bool process(Convolution *c, int idx, const float *in1, const float *in2, float *L, float *R, size_t nframes);
    ca_engine *engine() const { return _engine; }
    size_t size() const { return _members.size(); }

private:
    explicit SharedEngine(const EngineOptions &opt) : _opt(opt) {}
    bool build(size_t period, float sampleRate);  // under _buildMutex

    EngineOptions _opt;
    std::vector<Convolution *> _members;  // index = instance
    std::mutex _buildMutex;
    ca_engine *_engine = nullptr;
    size_t _period = 0, _slotsPerMember = 0;
    float *_in = nullptr, *_out = nullptr;  // pinned [instance][2][period]
    std::atomic<bool> _dirty{true};
    std::atomic<int> _arrived{0};
    std::atomic<uint64_t> _generation{0};
    std::atomic<bool> _ok{false};
};
