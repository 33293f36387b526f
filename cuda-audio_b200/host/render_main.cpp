// ca_render -- headless wav-in / wav-out harness for the B200 convolution engine.
//
// It replaces the reference's JACK-bound executable (src/main.cu) for offline use and follows
// the same setup sequence (main.cu:18-93): read settings.txt-style keys, build conv.count/2
// `Convolution` instances, load each input's CC numbers and initial values, prepare() every IR
// listed in the index file, start() the JACK client and connect 2 capture + 2 playback ports.
// Instead of a running jackd, the in-process headless JACK (headless_jack.cpp) drives the
// process callback once per period over an input wav (or a synthetic signal) and the playback
// buffers are collected into an output wav.
//
//   ca_render --settings settings.txt --in dry.wav --out wet.wav [--period 256]
//   ca_render --ir hall.wav --in dry.wav --out wet.wav --wet 1 --dry 0 --period 256
//   ca_render --synthetic-ir 1.0 --synthetic-in 10 --rate 44100 --mono --out out.wav --json stats.json
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iterator>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "convolution.h"
#include "headless_jack.h"
#include "settings_file.h"
#include "wavfile.h"

namespace {

struct Options {
    std::string settings, in, out, ir, json;
    size_t period = 256, fftSize = 0;
    unsigned rate = 0;
    double synthIr = 0, synthIn = 0;
    bool mono = false;
    int bits = 32, device = -1, warmup = 0;  // device -1: engine.device / engine.gpus of the settings file (default 0)
    float wet = -1, dry = -1, level = -1, panWet = -2, panDry = -2;
    long predelay = -1;
    bool resample = false;
    unsigned resampleTo = 0;
};

void usage()
{
    fprintf(stderr,
            "usage: ca_render [--settings FILE] [--ir WAV | --synthetic-ir SEC] [--in WAV | --synthetic-in SEC] --out WAV\n"
            "                 [--period N] [--rate HZ] [--fft-size N] [--mono] [--bits 16|24|32] [--device N]\n"
            "                 [--wet X] [--dry X] [--level X] [--pan-wet X] [--pan-dry X] [--predelay N]\n"
            "                 [--warmup PERIODS] [--json FILE] [--resample]\n"
            "       ca_render [--resample-to HZ] --dump-wav WAV SCALE OUT.f32 | --dump-settings FILE | --midi-parse FILE\n");
}

// exponentially decaying Gaussian noise, T60 = 0.8 x length, unit energy
std::vector<float> synth_ir(size_t frames, unsigned seed)
{
    std::mt19937 rng(seed);
    std::normal_distribution<float> g(0.f, 1.f);
    std::vector<float> h(frames);
    double e = 0;
    for (size_t n = 0; n < frames; n++) { h[n] = g(rng) * (float)exp(-6.91 * (double)n / (0.8 * (double)frames)); e += (double)h[n] * h[n]; }
    const float s = (float)(1.0 / sqrt(e));
    for (auto &v : h) v *= s;
    return h;
}

std::vector<float> synth_audio(size_t frames, unsigned seed)
{
    std::mt19937 rng(seed);
    std::normal_distribution<float> g(0.f, 0.1f);
    std::vector<float> x(frames);
    for (auto &v : x) v = std::min(0.9f, std::max(-0.9f, g(rng)));
    return x;
}

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
        if (a == "--settings") o.settings = next();
        else if (a == "--in") o.in = next();
        else if (a == "--out") o.out = next();
        else if (a == "--ir") o.ir = next();
        else if (a == "--json") o.json = next();
        else if (a == "--period") o.period = (size_t)atol(next());
        else if (a == "--fft-size") o.fftSize = (size_t)atol(next());
        else if (a == "--rate") o.rate = (unsigned)atol(next());
        else if (a == "--synthetic-ir") o.synthIr = atof(next());
        else if (a == "--synthetic-in") o.synthIn = atof(next());
        else if (a == "--mono") o.mono = true;
        else if (a == "--bits") o.bits = atoi(next());
        else if (a == "--device") o.device = atoi(next());
        else if (a == "--warmup") o.warmup = atoi(next());
        else if (a == "--wet") o.wet = (float)atof(next());
        else if (a == "--dry") o.dry = (float)atof(next());
        else if (a == "--level") o.level = (float)atof(next());
        else if (a == "--pan-wet") o.panWet = (float)atof(next());
        else if (a == "--pan-dry") o.panDry = (float)atof(next());
        else if (a == "--predelay") o.predelay = atol(next());
        else if (a == "--resample") o.resample = true;             // IRs whose rate differs from the run's are converted on load
        else if (a == "--resample-to") o.resampleTo = (unsigned)atol(next());  // for --dump-wav (give it first)
        else if (a == "--dump-settings") {  // parser check (no GPU needed): one "key=value" line per key, then typed reads
            Settings st;
            try { st.open(next()); } catch (std::exception &e) { fprintf(stderr, "ca_render: %s\n", e.what()); return 1; }
            for (auto &kv : st) printf("%s=%s\n", kv.first.c_str(), kv.second.value.c_str());
            return 0;
        } else if (a == "--midi-parse") {   // stream-parser check (no GPU needed): one hex line per complete message
            std::ifstream f(next(), std::ios::binary);
            std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            MidiParser parser([](const uint8_t *m, size_t len) { for (size_t k = 0; k < len; k++) printf("%02x%s", m[k], k + 1 == len ? "\n" : " "); });
            for (size_t k = 0; k < bytes.size(); k += 5) parser.feed(bytes.data() + k, std::min<size_t>(5, bytes.size() - k));  // odd chunking on purpose
            return 0;
        } else if (a == "--midi-listen") {  // device-thread check: open PATH, print CC -> parameter changes of a Convolution for SECONDS
            const char *path = next();
            const double seconds = atof(next());
            Convolution c("midi_probe", 1024);
            RawMidi::Device dev(path);
            auto &cc = c.cc[0];
            cc.device = &dev; dev.handler = &c;
            cc.message = 0xB0; cc.select = 20; cc.predelay = 21; cc.dry = 22; cc.wet = 23; cc.speed = 24; cc.panDry = 25; cc.panWet = 26; cc.level = 27;
            if (!dev.start()) { fprintf(stderr, "ca_render: %s\n", dev.error.c_str()); return 1; }
            printf("listening\n"); fflush(stdout);
            usleep((useconds_t)(seconds * 1e6));
            dev.stop();
            printf("{\"predelay\": %zu, \"dry\": %.6f, \"wet\": %.6f, \"speed\": %zu, \"panDry\": %.6f, \"panWet\": %.6f, \"level\": %.6f}\n",
                   cc.value.predelay, cc.value.dry, cc.value.wet, cc.value.speed, cc.value.panDry, cc.value.panWet, cc.value.level);
            return 0;
        } else if (a == "--dump-wav") {     // decoder check (no GPU needed): header + raw float32 samples to a file
            const char *path = next();
            const float scale = (float)atof(next());
            const char *outPath = next();
            WavData w = wav_read(path, scale);
            if (!w.ok()) { fprintf(stderr, "ca_render: %s\n", w.error.c_str()); return 1; }
            if (o.resampleTo) w = wav_resample(w, o.resampleTo);
            printf("{\"channels\": %u, \"rate\": %u, \"bits\": %u, \"format\": %u, \"frames\": %zu}\n", w.channels, w.sampleRate, w.bitsPerSample, w.audioFormat, w.frames);
            FILE *f = fopen(outPath, "wb");
            if (!f) return 1;
            for (auto &c : w.ch) fwrite(c.data(), sizeof(float), c.size(), f);
            fclose(f);
            return 0;
        }
        else { usage(); return 2; }
    }
    if (o.out.empty() && o.json.empty()) { usage(); return 2; }

    // ---- input signal ----
    WavData input;
    if (!o.in.empty()) {
        input = wav_read(o.in, 1.0f);
        if (!input.ok()) { fprintf(stderr, "ca_render: %s\n", input.error.c_str()); return 1; }
    } else {
        if (o.synthIn <= 0) o.synthIn = 5.0;
        input.sampleRate = o.rate ? o.rate : 48000;
        input.channels = o.mono ? 1 : 2;
        input.frames = (size_t)(o.synthIn * input.sampleRate);
        for (int c = 0; c < input.channels; c++) input.ch.push_back(synth_audio(input.frames, 2000 + c));
    }
    const unsigned rate = o.rate ? o.rate : input.sampleRate;
    hj_set_sample_rate(rate);
    hj_set_buffer_size((unsigned)o.period);  // JackClient::start() -> Convolution::onStart() builds the engine for it

    // ---- instances (main.cu:25-93) ----
    Settings settings;
    size_t numInstances = 1;
    if (!o.settings.empty()) {
        try { settings.open(o.settings); } catch (std::exception &e) { fprintf(stderr, "ca_render: %s\n", e.what()); return 1; }
        const uint32_t count = settings.u32("conv.count");
        if (count % 2) { fprintf(stderr, "ca_render: conv.count must be a multiple of 2\n"); return 1; }
        numInstances = count / 2;
        Convolution::setDefaultOptions(EngineOptions::fromSettings(settings));  // engine.* keys
    }
    // optional IR sample-rate conversion: --resample (to the run's rate) or `resample <Hz>` in the settings file
    unsigned irRate = o.resample ? rate : 0;
    if (!o.settings.empty() && settings.has("resample")) irRate = settings.u32("resample");
    std::vector<std::unique_ptr<Convolution>> inst;
    for (size_t n = 0; n < numInstances; n++) {
        size_t fftSize = o.fftSize ? o.fftSize : CONV_DEFAULT_FFTSIZE;
        if (o.settings.empty() && !o.fftSize) {  // size the IR capacity to the IR (the reference needs a hand-picked fftSize)
            size_t frames = (size_t)((o.synthIr > 0 ? o.synthIr : 1.0) * rate);
            if (!o.ir.empty()) {
                WavData probe = wav_read(o.ir, 0.5f);
                if (probe.ok()) frames = irRate && probe.sampleRate ? (size_t)std::ceil((double)probe.frames * irRate / probe.sampleRate) : probe.frames;
            }
            fftSize = 1;
            while (fftSize < frames + o.period) fftSize <<= 1;
        }
        if (!o.settings.empty()) {
            const uint32_t fs1 = settings.u32("conv[%d].fftSize", (int)(n * 2)), fs2 = settings.u32("conv[%d].fftSize", (int)(n * 2 + 1));
            if (fs1 != fs2) { fprintf(stderr, "ca_render: a convolution pair needs identical fft sizes\n"); return 1; }
            if (!o.fftSize) fftSize = fs1;
        }
        auto c = std::make_unique<Convolution>(std::string("cudaconv_") + char('1' + (int)n), fftSize);
        if (o.device >= 0) c->setDevice(o.device);  // --device overrides engine.device / engine.gpus
        c->setSampleRate((float)rate);
        for (int i = 0; i < 2; i++) {
            const int idx = (int)(n * 2 + i);
            auto &cc = c->cc[i];
            if (!o.settings.empty()) {
                cc.message = settings.u8("conv[%d].cc.message", idx);
                cc.select = settings.u8("conv[%d].cc.select", idx);
                cc.predelay = settings.u8("conv[%d].cc.predelay", idx);
                cc.dry = settings.u8("conv[%d].cc.dry", idx);
                cc.wet = settings.u8("conv[%d].cc.wet", idx);
                cc.speed = settings.u8("conv[%d].cc.speed", idx);
                cc.panDry = settings.u8("conv[%d].cc.panDry", idx);
                cc.panWet = settings.u8("conv[%d].cc.panWet", idx);
                cc.level = settings.u8("conv[%d].cc.level", idx);
                cc.value.select = settings.u32("conv[%d].value.select", idx);
                cc.value.predelay = settings.u32("conv[%d].value.predelay", idx);
                cc.value.dry = settings.f32("conv[%d].value.dry", idx);
                cc.value.wet = settings.f32("conv[%d].value.wet", idx);
                cc.value.speed = settings.u32("conv[%d].value.speed", idx);
                cc.value.panDry = settings.f32("conv[%d].value.panDry", idx);
                cc.value.panWet = settings.f32("conv[%d].value.panWet", idx);
                cc.value.level = settings.f32("conv[%d].value.level", idx);
                std::ifstream index(settings.str("conv[%d].index", idx));
                std::string path;
                for (size_t j = 0; std::getline(index, path); j++) {
                    if (path.empty()) continue;
                    WavFile w(path, irRate);
                    if (!w.error.empty()) { fprintf(stderr, "ca_render: %s\n", w.error.c_str()); return 1; }
                    c->prepare(j, w, o.period);
                }
            }
            if (o.wet >= 0) cc.value.wet = o.wet;
            if (o.dry >= 0) cc.value.dry = o.dry;
            if (o.level >= 0) cc.value.level = o.level;
            if (o.panWet >= -1) cc.value.panWet = o.panWet;
            if (o.panDry >= -1) cc.value.panDry = o.panDry;
            if (o.predelay >= 0) cc.value.predelay = (size_t)o.predelay;
        }
        if (o.settings.empty()) {
            if (!o.ir.empty()) {
                WavFile w(o.ir, irRate);
                if (!w.error.empty()) { fprintf(stderr, "ca_render: %s\n", w.error.c_str()); return 1; }
                c->prepare(0, w, o.period);
            } else {
                const size_t frames = (size_t)((o.synthIr > 0 ? o.synthIr : 1.0) * rate);
                for (int s = 0; s < 2; s++) {
                    std::vector<float> l = synth_ir(frames, 1000 + 2 * s), r = o.mono ? l : synth_ir(frames, 1001 + 2 * s);
                    WavFile w(l.data(), r.data(), frames, rate);
                    c->prepare((size_t)s, w, o.period);
                }
                c->cc[0].value.select = 0;
                c->cc[1].value.select = 1;
            }
        }
        if (c->lastError()) { fprintf(stderr, "ca_render: %s\n", c->lastErrorText().c_str()); return 1; }
        c->start();
        if (!c->isRunning()) return 1;
        for (int i = 0; i < 2; i++) {  // main.cu:83-90
            const int idx = (int)(n * 2 + i);
            const std::string inPort = o.settings.empty() ? "system:capture_" + std::to_string(idx + 1) : settings.str("conv[%d].input", idx);
            const std::string outPort = o.settings.empty() ? "system:playback_" + std::to_string(idx + 1) : settings.str("conv[%d].output", idx);
            jack_connect(c->handle, inPort.c_str(), jack_port_name(c->capture[i]));
            jack_connect(c->handle, jack_port_name(c->playback[i]), outPort.c_str());
        }
        inst.push_back(std::move(c));
    }

    // ---- run: "system:capture_k" = input channel k-1, "system:playback_k" = output channel k-1 ----
    const size_t B = o.period;
    const size_t periods = (input.frames + B - 1) / B;
    const size_t outChannels = o.mono ? 1 : 2 * numInstances;
    std::vector<std::vector<float>> out(outChannels, std::vector<float>(periods * B, 0.f));
    std::vector<float> silence(B, 0.f), scratch(B, 0.f), scratch2(B, 0.f);
    std::vector<std::vector<float>> inPadded(input.channels);
    for (int c = 0; c < input.channels; c++) { inPadded[c] = input.ch[c]; inPadded[c].resize(periods * B, 0.f); }
    auto channelOf = [](const char *peer, const char *prefix) -> int {
        if (!peer) return -1;
        const char *p = strstr(peer, prefix);
        return p ? atoi(p + strlen(prefix)) - 1 : -1;
    };
    std::vector<double> wall;
    wall.reserve(periods);
    // one JACK cycle for every client.  With `engine.shared` the clients are instances of one batched engine
    // and meet at a rendezvous inside onProcess(): their callbacks must run on separate threads, as JACK's do.
    const bool threaded = inst.size() > 1 && inst[0]->options().shared > 1;
    auto cycle_all = [&] {
        if (!threaded) { for (auto &c : inst) hj_cycle(c->handle, (jack_nframes_t)B); return; }
        std::vector<std::thread> th;
        for (auto &c : inst) th.emplace_back([&c, B] { hj_cycle(c->handle, (jack_nframes_t)B); });
        for (auto &t : th) t.join();
    };
    for (int w = 0; w < o.warmup; w++) {  // silent warm-up periods (lets the wet glide converge)
        for (auto &c : inst)
            for (int i = 0; i < 2; i++) { hj_port_set_buffer(c->capture[i], silence.data()); hj_port_set_buffer(c->playback[i], i == 0 ? scratch.data() : scratch2.data()); }
        cycle_all();
    }
    for (size_t t = 0; t < periods; t++) {
        const auto t0 = std::chrono::steady_clock::now();
        for (auto &c : inst) {
            for (int i = 0; i < 2; i++) {
                int ic = channelOf(hj_port_peer(c->handle, c->capture[i]), "capture_");
                if (o.mono) ic = (i == 0) ? 0 : -1;
                else if (input.channels == 1 && ic >= 0) ic = 0;
                hj_port_set_buffer(c->capture[i], (ic >= 0 && ic < input.channels) ? inPadded[ic].data() + t * B : silence.data());
                int oc = channelOf(hj_port_peer(c->handle, c->playback[i]), "playback_");
                if (o.mono) oc = (i == 0) ? 0 : -1;
                hj_port_set_buffer(c->playback[i], (oc >= 0 && oc < (int)outChannels) ? out[oc].data() + t * B : scratch.data());
            }
        }
        cycle_all();
        wall.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    }
    int rc = 0;
    for (auto &c : inst)
        if (c->lastError()) { fprintf(stderr, "ca_render: %s\n", c->lastErrorText().c_str()); rc = 1; }
    for (auto &ch : out) ch.resize(input.frames);
    if (!o.out.empty() && !wav_write(o.out, out, rate, o.bits)) { fprintf(stderr, "ca_render: cannot write %s\n", o.out.c_str()); rc = 1; }

    std::vector<double> sorted = wall;
    std::sort(sorted.begin(), sorted.end());
    const double p50 = sorted.empty() ? 0 : sorted[sorted.size() / 2], p99 = sorted.empty() ? 0 : sorted[std::min(sorted.size() - 1, (size_t)(0.99 * sorted.size()))];
    char js[1024];
    snprintf(js, sizeof(js),
             "{\"instances\": %zu, \"period\": %zu, \"rate\": %u, \"periods\": %zu, \"deadline_us\": %.1f, \"p50_us\": %.1f, \"p99_us\": %.1f, "
             "\"avg_runtime_ms\": %.4f, \"irs\": %zu, \"frames\": %zu}",
             numInstances, B, rate, periods, 1e6 * (double)B / rate, p50, p99, inst[0]->avgRuntime(), inst[0]->numIRs(), input.frames);
    if (!o.json.empty()) { FILE *f = fopen(o.json.c_str(), "w"); if (f) { fputs(js, f); fputc('\n', f); fclose(f); } }
    printf("%s\n", js);
    for (auto &c : inst) {
        Log::info(c->name, "Average convolution runtime: %f", c->avgRuntime());  // main.cu:106
        c->stop();
    }
    return rc;
}
