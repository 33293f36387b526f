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
    // member idx's client was stopped: the others no longer wait for it (it takes part again with its next call)
    void standDown(int idx);
    void irChanged(Convolution *c);  // a member (re)prepared an IR: rebuild at the next rendezvous / buildNow
    std::mutex &irMutex() { return _buildMutex; }  // members change their IR set under it (a build walks every member's IRs)
    // Build (or rebuild) now, from any thread that is NOT inside process(): waits until every member has left
    // process() (members that arrive meanwhile answer their period with silence), then builds.
    bool buildNow(size_t period, float sampleRate);
    // one period of member `idx`; false: silence (engine not built / being rebuilt, build failure)
    bool process(Convolution *c, int idx, const float *in1, const float *in2, float *L, float *R, size_t nframes);
    ca_engine *engine() const { return _engine; }
    size_t size() const { return _members.size(); }
    bool full();  // every one of the opt.shared seats is taken
    // diagnostics
    uint64_t batches() const { return _batches.load(std::memory_order_relaxed); }   // ca_process calls for the whole batch
    uint64_t rebuilds() const { return _rebuilds.load(std::memory_order_relaxed); } // engine (re)builds so far
    uint64_t dropped() const { return _dropped.load(std::memory_order_relaxed); }   // members set aside because they stopped arriving
    int taking_part() const { return _live.load(std::memory_order_relaxed); }

private:
    explicit SharedEngine(const EngineOptions &opt);
    bool build(size_t period, float sampleRate);  // under _buildMutex, nobody else inside process()'s buffer accesses
    void runBatch(Convolution *c, size_t nframes, uint64_t gen);
    bool everybodyArrived(uint64_t state, uint64_t gen) const;
    bool waitGeneration(Convolution *c, size_t nframes, uint64_t gen);
    bool enterStaging(int idx);
    bool processPipelined(Convolution *c, int idx, const float *in1, const float *in2, float *L, float *R, size_t nframes);
    void dropStalled(uint64_t gen);

    EngineOptions _opt;
    std::vector<Convolution *> _members;  // index = instance; only touched under _buildMutex
    std::mutex _buildMutex;
    // engine state: written by build() only (exclusive, see process()), read by the members
    ca_engine *_engine = nullptr;
    size_t _period = 0, _slotsPerMember = 0, _builtMembers = 0;
    float *_in = nullptr, *_out = nullptr;  // pinned [instance][2][period]
    std::atomic<bool> _dirty{true};
    std::atomic<bool> _ok{false};
    // rendezvous: generation << 16 | members that arrived in this generation.  The batch of a generation runs once,
    // on the thread of whichever waiting member finds everybody arrived; it ends the generation.
    std::atomic<uint64_t> _state{0};
    std::atomic<bool> _runner{false};
    std::atomic<uint64_t> _okGen{0};  // generation + 1 of the last batch that ran and succeeded
    // members are expected at the rendezvous from join() until they stop arriving (sharedTimeoutMs), stand down or
    // leave(); a member that was set aside takes part again with its next call
    std::unique_ptr<std::atomic<bool>[]> _active;
    std::unique_ptr<std::atomic<uint64_t>[]> _arrivedGen;  // generation + 1 of the member's last arrival
    std::unique_ptr<std::atomic<uint64_t>[]> _stagedGen;   // engine.shared_latency 1: generation + 1 the member last handed a block to
    std::atomic<int> _live{0};
    // build exclusion: a builder raises _exclusive, then needs _inside == 0 (buildNow: waits) or _staging == 0
    // (rebuild at the rendezvous: gives up for this cycle); members raise their counter first, then read _exclusive
    std::atomic<int> _exclusive{0}, _inside{0}, _staging{0};
    std::atomic<bool> _evict{false};  // buildNow() is waiting for _inside == 0: members waiting at the rendezvous leave
    std::atomic<uint64_t> _batches{0}, _rebuilds{0}, _dropped{0};
};
