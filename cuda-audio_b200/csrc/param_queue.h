// param_queue.h -- parameter hand-off of the C ABI (plain C++, no CUDA: also compiled by tests/hostsim under ThreadSanitizer).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>

#include "../../include/cuda_audio_b200.h"

// Parameter hand-off: ca_set_params / ca_set_glide may be called from ANY thread (MIDI thread, UI, the
// real-time thread itself) and never take a lock: commands go through a bounded multi-producer /
// single-consumer ring (Vyukov's sequence-numbered slots) that the processing thread drains at the top of
// the next period ("set_params lock-free from any thread").
struct ParamCmd {
    uint32_t item = 0, kind = 0;  // kind 0: parameter block, 1: glide jump
    ca_params p{};
    float glide = 0.f;
};

class ParamQueue {
public:
    ParamQueue() { resize(1u << 13); }
    ~ParamQueue() { delete[] slots_; }
    // before the engine is shared: room for a full set-up pass (parameters + glide of every item, twice)
    void resize(uint64_t min_slots)
    {
        delete[] slots_;
        kSlots = 1u << 13;
        while (kSlots < min_slots) kSlots <<= 1;
        slots_ = new Slot[kSlots];
        for (uint64_t i = 0; i < kSlots; i++) slots_[i].seq.store(i, std::memory_order_relaxed);
        enq_.store(0, std::memory_order_relaxed);
        deq_ = 0;
    }
    bool push(const ParamCmd &c)
    {
        uint64_t pos = enq_.load(std::memory_order_relaxed);
        for (;;) {
            Slot &s = slots_[pos & (kSlots - 1)];
            const int64_t dif = (int64_t)s.seq.load(std::memory_order_acquire) - (int64_t)pos;
            if (dif == 0) {
                if (enq_.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) { s.cmd = c; s.seq.store(pos + 1, std::memory_order_release); return true; }
            } else if (dif < 0) return false;  // full
            else pos = enq_.load(std::memory_order_relaxed);
        }
    }
    bool pop(ParamCmd *c)  // single consumer
    {
        Slot &s = slots_[deq_ & (kSlots - 1)];
        if (s.seq.load(std::memory_order_acquire) != deq_ + 1) return false;
        *c = s.cmd;
        s.seq.store(deq_ + kSlots, std::memory_order_release);
        deq_++;
        return true;
    }

private:
    struct Slot { std::atomic<uint64_t> seq; ParamCmd cmd; };
    uint64_t kSlots = 0;
    Slot *slots_ = nullptr;
    std::atomic<uint64_t> enq_{0};
    uint64_t deq_ = 0;
};

// What ca_get_params answers: the last block handed to ca_set_params per (instance, input), readable from any thread
// while any number of threads set it.  A seqlock per item whose payload is words of relaxed atomics (no data race in
// the language's sense either): writers take the item by moving its sequence from even to odd (two writers of the same
// item at the same instant: the second spins for the few stores of the first), readers retry while it is odd or moved.
class ParamMirror {
public:
    static constexpr size_t kWords = sizeof(ca_params) / sizeof(uint32_t);
    static_assert(sizeof(ca_params) % sizeof(uint32_t) == 0, "ca_params is made of 32-bit fields");
    void resize(size_t n_items)
    {
        n_ = n_items;
        seq_.reset(new std::atomic<uint32_t>[n_items]);
        words_.reset(new std::atomic<uint32_t>[n_items * kWords]);
        for (size_t i = 0; i < n_items; i++) seq_[i].store(0, std::memory_order_relaxed);
        for (size_t i = 0; i < n_items * kWords; i++) words_[i].store(0, std::memory_order_relaxed);
    }
    size_t size() const { return n_; }
    void set(size_t item, const ca_params &p)
    {
        uint32_t w[kWords];
        memcpy(w, &p, sizeof(p));
        std::atomic<uint32_t> &sq = seq_[item];
        uint32_t s = sq.load(std::memory_order_relaxed);
        for (;;) {
            if (s & 1u) { s = sq.load(std::memory_order_relaxed); continue; }
            if (sq.compare_exchange_weak(s, s + 1, std::memory_order_acquire, std::memory_order_relaxed)) break;
        }
        // release stores: none of them may become visible before the odd sequence (no fence: ThreadSanitizer models none)
        for (size_t k = 0; k < kWords; k++) words_[item * kWords + k].store(w[k], std::memory_order_release);
        sq.store(s + 2, std::memory_order_release);
    }
    void get(size_t item, ca_params *p) const
    {
        uint32_t w[kWords];
        const std::atomic<uint32_t> &sq = seq_[item];
        for (;;) {
            const uint32_t a = sq.load(std::memory_order_acquire);
            for (size_t k = 0; k < kWords; k++) w[k] = words_[item * kWords + k].load(std::memory_order_acquire);  // the re-check below stays below
            if (!(a & 1u) && sq.load(std::memory_order_relaxed) == a) break;
        }
        memcpy(p, w, sizeof(*p));
    }

private:
    size_t n_ = 0;
    std::unique_ptr<std::atomic<uint32_t>[]> seq_, words_;
};
