/*
 * oracle/ref_harness/harness.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Headless driver for the UNMODIFIED reference (limitz/cuda-audio src/conv.cu, wav.cu,
 * log.cu, jackclient.cu compiled from /root/reference/src where they lie; see Makefile).
 * It supplies a fake libjack (a port is a struct holding a caller-owned float buffer),
 * constructs the reference's own `Convolution` (conv.h:30-86) and drives it through its
 * public API only: `prepare()` (conv.cu:207), `onProcess()` (conv.cu:287), the public
 * `cc[2]`, `capture[2]`, `playback[2]` members (conv.h:51,56-57) and `avgRuntime()`.
 * `selectGpu()` (gpu.cu:38) is never called: it asserts on sm_100.
 *
 * Exposed as a C ABI so tests/ and bench.py (--impl reference) can call it with ctypes.
 */
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "conv.h" /* the reference's header, via -I$(REFSRC) */

/* ---------------- fake libjack ------------------------------------------------------------ */
struct _jack_port { float *buf; };
struct _jack_client { int unused; };
extern "C" {
void *jack_port_get_buffer(jack_port_t *port, jack_nframes_t) { return port ? port->buf : nullptr; }
jack_port_t *jack_port_register(jack_client_t *, const char *, const char *, unsigned long, unsigned long) { return new _jack_port{nullptr}; }
int jack_activate(jack_client_t *) { return 0; }
jack_client_t *jack_client_open(const char *, jack_options_t, jack_status_t *status, ...) { if (status) *status = (jack_status_t)0; return new _jack_client{0}; }
int jack_set_process_callback(jack_client_t *, JackProcessCallback, void *) { return 0; }
void jack_on_shutdown(jack_client_t *, JackShutdownCallback, void *) {}
jack_nframes_t jack_get_sample_rate(jack_client_t *) { return 48000; }
jack_nframes_t jack_get_buffer_size(jack_client_t *) { return 0; }  /* unknown: the harness never calls start() */
int jack_client_close(jack_client_t *c) { delete c; return 0; }
int jack_connect(jack_client_t *, const char *, const char *) { return 0; }
const char *jack_port_name(const jack_port_t *) { return "fake"; }
}

/* ---------------- helpers ------------------------------------------------------------------ */
static void put_u32(FILE *f, uint32_t v) { fwrite(&v, 4, 1, f); }
static void put_u16(FILE *f, uint16_t v) { fwrite(&v, 2, 1, f); }

/* minimal 16-bit stereo wav of `frames` silent frames: lets the reference's own WavFile
 * constructor (wav.cu:46-118) size its buffer; the float data is then injected. */
static std::string write_dummy_wav(size_t frames)
{
    char path[] = "/tmp/ca_ref_dummy_XXXXXX";
    int fd = mkstemp(path);
    FILE *f = fdopen(fd, "wb");
    uint32_t dataBytes = (uint32_t)(frames * 4);
    fwrite("RIFF", 1, 4, f); put_u32(f, 36 + dataBytes); fwrite("WAVE", 1, 4, f);
    fwrite("fmt ", 1, 4, f); put_u32(f, 16); put_u16(f, 1); put_u16(f, 2); put_u32(f, 48000);
    put_u32(f, 48000 * 4); put_u16(f, 4); put_u16(f, 16);
    fwrite("data", 1, 4, f); put_u32(f, dataBytes);
    std::vector<char> zeros(1 << 16, 0);
    for (size_t left = dataBytes; left;) { size_t n = left < zeros.size() ? left : zeros.size(); fwrite(zeros.data(), 1, n, f); left -= n; }
    fclose(f);
    return path;
}

struct RefInstance {
    Convolution conv;
    _jack_port in[2], out[2];
    size_t nIR = 0;
    RefInstance(const std::string &name, size_t fftSize) : conv(name, fftSize)
    {
        conv.capture[0] = &in[0]; conv.capture[1] = &in[1];
        conv.playback[0] = &out[0]; conv.playback[1] = &out[1];
    }
};

extern "C" {

int ref_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }

void *ref_create(size_t fftSize, int device)
{
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    return new RefInstance("refconv", fftSize);
}

void ref_destroy(void *h) { /* the reference has no destructor that frees (conv.h:53-54) */ (void)h; }

/* prepare() with exact fp32 IR data (no PCM quantisation) */
int ref_prepare(void *h, size_t idx, const float *left, const float *right, size_t frames, size_t nframes)
{
    auto *r = (RefInstance *)h;
    std::string path = write_dummy_wav(frames);
    {
        WavFile w(path);
        if (w.numFrames != frames) { unlink(path.c_str()); return -1; }
        std::vector<float2> host(frames);
        for (size_t i = 0; i < frames; i++) host[i] = make_float2(left[i], right[i]);
        cudaMemcpy(w.buffer, host.data(), frames * sizeof(float2), cudaMemcpyHostToDevice);
        r->conv.prepare(idx, w, nframes);
    }
    unlink(path.c_str());
    if (idx + 1 > r->nIR) r->nIR = idx + 1;
    return 0;
}

/* prepare() from a real wav file through the reference's own PCM decoder */
int ref_prepare_wav(void *h, size_t idx, const char *path, size_t nframes)
{
    auto *r = (RefInstance *)h;
    WavFile w(path);
    r->conv.prepare(idx, w, nframes);
    if (idx + 1 > r->nIR) r->nIR = idx + 1;
    return (int)w.numFrames;
}

/* decode a wav with the reference's WavFile (wav.cu) and return planar floats */
long ref_wav_decode(const char *path, float *left, float *right, size_t maxFrames)
{
    WavFile w(path);
    size_t n = w.numFrames < maxFrames ? w.numFrames : maxFrames;
    std::vector<float2> host(n);
    cudaMemcpy(host.data(), w.buffer, n * sizeof(float2), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < n; i++) { left[i] = host[i].x; right[i] = host[i].y; }
    return (long)w.numFrames;
}

void ref_set_cc(void *h, int input, size_t select, size_t predelay, size_t speed, long vsteps,
                float dry, float wet, float panDry, float panWet, float level)
{
    auto &v = ((RefInstance *)h)->conv.cc[input].value;
    v.select = select; v.predelay = predelay; v.speed = speed;
    if (vsteps >= 0) v.vsteps = (size_t)vsteps;
    v.dry = dry; v.wet = wet; v.panDry = panDry; v.panWet = panWet; v.level = level;
}

/* feed a MIDI CC through the reference's own handler (conv.cu:255-285) */
void ref_midi_cc(void *h, int input, uint8_t ccNumberField, uint8_t value)
{
    auto *r = (RefInstance *)h;
    static RawMidi::Device dev("fake");
    auto &cc = r->conv.cc[input];
    cc.device = &dev; cc.message = 176;
    /* give every parameter its own CC number 1..8 in declaration order */
    cc.select = 1; cc.predelay = 2; cc.dry = 3; cc.wet = 4; cc.speed = 5; cc.panDry = 6; cc.panWet = 7; cc.level = 8;
    uint8_t msg[3] = {176, ccNumberField, value};
    auto *other = &r->conv.cc[1 - input];
    RawMidi::Device *saved = other->device; other->device = nullptr;
    r->conv.onMidiMessage(&dev, msg, 3);
    other->device = saved;
}

void ref_get_cc(void *h, int input, size_t *select, size_t *predelay, size_t *speed, size_t *vsteps,
                float *dry, float *wet, float *panDry, float *panWet, float *level)
{
    auto &v = ((RefInstance *)h)->conv.cc[input].value;
    *select = v.select; *predelay = v.predelay; *speed = v.speed; *vsteps = v.vsteps;
    *dry = v.dry; *wet = v.wet; *panDry = v.panDry; *panWet = v.panWet; *level = v.level;
}

int ref_process(void *h, const float *in1, const float *in2, float *L, float *R, size_t nframes)
{
    auto *r = (RefInstance *)h;
    r->in[0].buf = (float *)in1; r->in[1].buf = (float *)in2;
    r->out[0].buf = L; r->out[1].buf = R;
    r->conv.onProcess(nframes);
    return 0;
}

/* many periods in one call (avoids ctypes overhead per period) */
int ref_render(void *h, const float *in1, const float *in2, float *L, float *R, size_t nframes, size_t periods)
{
    for (size_t t = 0; t < periods; t++)
        ref_process(h, in1 + t * nframes, in2 + t * nframes, L + t * nframes, R + t * nframes, nframes);
    return 0;
}

double ref_avg_runtime_ms(void *h) { return ((RefInstance *)h)->conv.avgRuntime(); }

/* ---- throughput / latency harness: K instances on K host threads (one JACK client per
 * instance in the reference, main.cu:32-39), lock-stepped per period by a barrier.
 * wall_us[p] = wall time from period start until ALL K instances finished period p. */
struct SpinBarrier {
    std::atomic<int> count{0}; std::atomic<int> gen{0}; int n;
    explicit SpinBarrier(int n) : n(n) {}
    void wait()
    {
        int g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) { count.store(0, std::memory_order_relaxed); gen.fetch_add(1, std::memory_order_release); }
        else { int spins = 0; while (gen.load(std::memory_order_acquire) == g) { if (++spins > 2000) std::this_thread::yield(); } }
    }
};

int ref_bench(void **handles, int K, size_t nframes, int warmup, int periods, double *wall_us)
{
    std::vector<std::vector<float>> in1(K), in2(K), oL(K), oR(K);
    for (int k = 0; k < K; k++) {
        in1[k].resize(nframes); in2[k].resize(nframes); oL[k].resize(nframes); oR[k].resize(nframes);
        unsigned s = 12345u + (unsigned)k;
        for (size_t i = 0; i < nframes; i++) { s = s * 1664525u + 1013904223u; in1[k][i] = ((int)(s >> 9) % 2000 - 1000) * 1e-4f; s = s * 1664525u + 1013904223u; in2[k][i] = ((int)(s >> 9) % 2000 - 1000) * 1e-4f; }
    }
    SpinBarrier bar(K);
    std::vector<std::chrono::steady_clock::time_point> t0(periods + warmup), t1(periods + warmup);
    auto worker = [&](int k) {
        for (int p = 0; p < warmup + periods; p++) {
            bar.wait();
            if (k == 0) t0[p] = std::chrono::steady_clock::now();
            ref_process(handles[k], in1[k].data(), in2[k].data(), oL[k].data(), oR[k].data(), nframes);
            bar.wait();
            if (k == 0) t1[p] = std::chrono::steady_clock::now();
        }
    };
    std::vector<std::thread> th;
    for (int k = 0; k < K; k++) th.emplace_back(worker, k);
    for (auto &t : th) t.join();
    for (int p = 0; p < periods; p++) wall_us[p] = std::chrono::duration<double, std::micro>(t1[warmup + p] - t0[warmup + p]).count();
    return 0;
}

} /* extern "C" */
