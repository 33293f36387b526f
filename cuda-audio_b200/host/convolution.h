// convolution.h -- host-side mirror of the reference's `Convolution` class (src/conv.h:30-86)
// on top of the B200 engine's C ABI (include/cuda_audio_b200.h).
//
// Same public surface, same meaning, so code written against the reference keeps working:
//   Convolution(name, fftSize)              conv.cu:142   fftSize now only bounds the IR length
//                                                         (IR <= fftSize - nframes, conv.cu:239)
//   prepare(idx, wav, nframes = 1024)       conv.cu:207   stereo IR -> bank slot idx
//   onProcess(nframes)                      conv.cu:287   one JACK period, 2 in -> 2 out
//   onStart()                               conv.cu:197   activate + register the 4 ports
//   cc[2] (CC numbers + .value block)       conv.h:33-50  read every period, written by anyone
//   capture[2], playback[2]                 conv.h:56-57
//   onMidiMessage(sender, buffer, len)      conv.cu:278   CC -> parameter mapping (handleCC)
//   avgRuntime()                            conv.h:61     mean ms per onProcess, first 10 skipped
// Differences, all deliberate: errors are reported (lastError()) instead of assert-aborting;
// an out-of-range `select` is ignored instead of dereferencing nullptr (conv.cu:340); the
// destructor frees everything (the reference leaks, conv.h:53-54).
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cuda_audio_b200.h"
#include "jack_client.h"
#include "rawmidi.h"
#include "wavfile.h"

#ifndef CONV_DEFAULT_FFTSIZE
#define CONV_DEFAULT_FFTSIZE (512 * 256)
#endif
#ifndef CONV_MAX_SPEED
#define CONV_MAX_SPEED CA_MAX_SPEED
#endif
#ifndef CONV_MAX_PREDELAY
#define CONV_MAX_PREDELAY CA_MAX_PREDELAY
#endif

class Settings;
class SharedEngine;

// How a Convolution object maps onto the B200 engine (not in the reference: its engine IS conv.cu).
// Settings-file keys (all optional, read by EngineOptions::fromSettings; ca_live / ca_render pass them on):
//   engine.device <n>            CUDA device ordinal
//   engine.tiers uniform|auto    uniform partitions (default) or non-uniform tiers (ca_config_auto_tiers)
//   engine.tier_growth <n>, engine.tier_max_block <n>
//   engine.async_tiers true      long tiers on a low-priority stream, two periods of slack (CA_FLAG_ASYNC_TIERS)
//   engine.graph false           host-driven launches instead of one CUDA graph per period
//   engine.l2_persist true       pin IR spectra + delay lines in L2 (CA_FLAG_L2_PERSIST)
//   engine.period <frames>       build the engine at prepare()/onStart() time for this period (default: JACK's buffer size)
//   engine.shared <N>            every N consecutive Convolution objects of the process are the N instances of
//                                ONE batched engine (one set of kernel launches per JACK cycle for all of them)
//   engine.shared_timeout_ms <n> how long the others wait for a member that stopped calling (default 200)
//   engine.shared_latency 1      shared batch without a rendezvous: every member's output is one period late, nobody waits
//                                inside a cycle -- for hosts that call their clients one after the other (jack1)
//   engine.gpus <G>              the process's Convolution objects are dealt round-robin onto G GPUs starting at
//                                engine.device (independent streams, no collective); shared batches
//                                form per GPU
//   engine.ir_split <G>          ONE object's IR is split by partition range over G GPUs (engine.device ..
//                                engine.device + G - 1) behind the same prepare / onProcess surface: the ca_group of
//                                the C ABI (BASELINE configs[4]); engine.exchange p2p|nccl picks the
//                                exchange (default p2p: fused into the MAC kernel over NVLink)
// The same options can come from the environment (CA_ENGINE_TIERS, CA_ENGINE_SHARED, CA_ENGINE_PERIOD,
// CA_ENGINE_DEVICE, CA_ENGINE_ASYNC_TIERS) for drivers that construct Convolution objects themselves.
struct EngineOptions {
    int device = 0;
    uint32_t flags = CA_FLAG_GRAPH;
    bool autoTiers = false;
    uint32_t tierGrowth = 0, tierMaxBlock = 0;
    uint32_t period = 0;
    uint32_t shared = 0;
    uint32_t gpus = 1;        // engine.gpus
    uint32_t irSplit = 0;     // engine.ir_split: > 0 = this many GPUs behind one object (ca_group)
    uint32_t exchange = 0;    // engine.exchange: ca_exchange (CA_EXCHANGE_P2P / CA_EXCHANGE_NCCL)
    uint32_t sharedLatency = 0;  // engine.shared_latency: 0 = rendezvous inside the cycle, 1 = hand in / take out, one period later
    uint32_t sharedTimeoutMs = 200;  // engine.shared: a member that has not arrived for this long is set aside until it calls again
    static EngineOptions fromEnv();
    static EngineOptions fromSettings(Settings &settings);
};

class Convolution : public JackClient, public RawMidi::MessageHandler {
public:
    struct CC {
        RawMidi::Device *device = nullptr;
        uint8_t message = 0;
        uint8_t select = 0, predelay = 0, dry = 0, wet = 0, speed = 0, panDry = 0, panWet = 0, level = 0;
        struct Value {
            size_t select = 0;    // IR index
            size_t predelay = 0;  // samples, [0, 8192)
            size_t speed = 100;   // glide length in periods
            size_t vsteps = 0;    // glide countdown
            float dry = 0.5f, wet = 0.5f;
            float panDry = 0.0f, panWet = 0.0f;
            float level = 1.0f;
        } value;
    } cc[2];

    explicit Convolution(const std::string &name = "Conv", size_t fftSize = CONV_DEFAULT_FFTSIZE);
    ~Convolution() override;

    JackPort capture[2];
    JackPort playback[2];

    void onProcess(size_t nframes) override;
    void onStart() override;
    void onStop() override;
    double avgRuntime() const { return _nruns > 0 ? _runtimeMs / _nruns : 0.0; }

    void prepare(size_t idx, const WavFile &wav, size_t nframes = 1024);

    void onMidiMessage(const RawMidi::Device *sender, const uint8_t *buffer, size_t len) override;

    // --- additions (not in the reference) ---
    int lastError() const { return _lastError; }          // ca_error of the last failing call, 0 = none
    const std::string &lastErrorText() const { return _lastErrorText; }
    ca_engine *engine() const { return _engine; }          // null until the engine is built (buildNow / onStart / first onProcess)
    ca_group *group() const { return _group; }             // engine.ir_split: the multi-GPU group that stands in for the engine
    void setDevice(int device) { _opt.device = device; }   // before the engine is built
    void setFlags(uint32_t flags) { _opt.flags = flags; }  // ca_flags, before the engine is built
    void setSampleRate(float fs) { _sampleRate = fs; }
    void setOptions(const EngineOptions &o);               // before the engine is built
    const EngineOptions &options() const { return _opt; }
    static void setDefaultOptions(const EngineOptions &o); // options of Convolution objects constructed from now on
    // Build the engine NOW, outside the real-time thread (allocation, graph capture, IR transforms, one silent
    // warm-up period).  onStart() calls it with JACK's buffer size; harnesses that know the period call it after
    // the last prepare().  Without it the first onProcess() builds the engine (not real-time safe).
    bool buildNow(size_t period);
    size_t numIRs() const { return _numIRs.load(std::memory_order_acquire); }
    static constexpr size_t kMaxIRs = 1024;                // prepare(idx, ...) accepts idx < kMaxIRs
    std::shared_ptr<SharedEngine> sharedGroup() const { return _shared; }  // engine.shared: the batch this object is an instance of
    uint64_t skippedPeriods() const { return _skipped; }   // periods answered with silence because prepare() held the engine

private:
    friend class SharedEngine;
    struct HostIR { std::vector<float> left, right; };
    bool buildEngine(size_t period);
    void pushParams(bool force);
    void pushParamsTo(ca_engine *e, uint32_t instance, size_t slotBase, bool force);
    void fail(int code, const char *what);

    size_t _fftSize;
    // time-domain IRs kept on the host so the engine can be (re)built.  Changed and walked only under _engineMutex (or the
    // shared group's irMutex()); the real-time and MIDI threads look at the lock-free summary below instead.
    std::map<size_t, HostIR> _irs;
    std::atomic<bool> _hasIR[kMaxIRs];
    std::atomic<size_t> _numIRs{0}, _firstIR{0};
    size_t _minPrepareFrames = 1024;
    ca_engine *_engine = nullptr;
    ca_group *_group = nullptr;  // engine.ir_split: used instead of _engine (same lock, same life cycle)
    bool built() const { return _engine || _group; }
    void destroyEngine();
    int loadIR(size_t idx, const HostIR &ir);
    int processBlock(uint32_t nframes);
    int resetEngine();
    std::mutex _engineMutex;  // prepare() on a live engine vs onProcess(): the RT thread only try_locks (silence on contention)
    size_t _period = 0, _engineSlots = 0, _engineCapFrames = 0;
    EngineOptions _opt;
    std::shared_ptr<SharedEngine> _shared;  // engine.shared: this object is instance _sharedIdx of a batched engine
    int _sharedIdx = -1;
    float _sampleRate = 48000.f;
    CC::Value _pushed[2];
    bool _havePushed = false;
    float *_in = nullptr, *_out = nullptr;  // pinned planar staging [2][period] (ca_host_alloc): no copy inside ca_process
    uint64_t _skipped = 0;
    std::atomic<bool> _stopping{false};  // between onStop() and the next onStart(): callbacks that still arrive answer with silence
    double _runtimeMs = 0;
    int _nruns = -10;  // conv.h:80: discard the first couple of runs
    int _lastError = 0;
    std::string _lastErrorText;
};
