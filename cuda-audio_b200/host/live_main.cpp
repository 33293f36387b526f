// ca_live -- the reference's executable (src/main.cu) on the B200 engine: read settings.txt, build
// conv.count/2 `Convolution` JACK clients, load the IR index files, connect the ports, run until
// Enter is pressed, print the average runtime (main.cu:18-116).  libjack is resolved at run time
// (jack_dl.cpp).  MIDI control: `conv[i].cc.device hw:C,D` opens the kernel's raw MIDI device
// (rawmidi.cpp, no libasound) and routes control changes through Convolution::onMidiMessage exactly
// like main.cu:43-52,86-87; a device that cannot be opened is reported and the instance runs from
// the settings file's values alone.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "convolution.h"
#include "settings_file.h"
#include "wavfile.h"

int main(int argc, char **argv)
{
    const std::string path = argc > 1 ? argv[1] : "settings.txt";
    Settings settings;
    try { settings.open(path); } catch (std::exception &e) { fprintf(stderr, "ca_live: %s\n", e.what()); return 1; }
    uint32_t count = 0;
    try { count = settings.u32("conv.count"); } catch (std::exception &) { fprintf(stderr, "ca_live: conv.count missing\n"); return 1; }
    if (count % 2) { fprintf(stderr, "ca_live: conv.count must be a multiple of 2\n"); return 1; }
    Convolution::setDefaultOptions(EngineOptions::fromSettings(settings));  // engine.* keys (convolution.h)
    // `resample <Hz>`: convert IRs recorded at another rate (the shipped library is 44.1 kHz) to the JACK
    // server's rate on load; absent / 0 keeps the reference's behaviour (IR played at the server's rate)
    const uint32_t irRate = settings.has("resample") ? settings.u32("resample") : 0;
    std::map<std::string, std::unique_ptr<RawMidi::Device>> midiDevices;
    std::vector<std::unique_ptr<Convolution>> inst;
    for (uint32_t n = 0; n < count / 2; n++) {
        const uint32_t fs1 = settings.u32("conv[%d].fftSize", (int)(2 * n)), fs2 = settings.u32("conv[%d].fftSize", (int)(2 * n + 1));
        if (fs1 != fs2) { fprintf(stderr, "ca_live: a convolution pair needs identical fft sizes\n"); return 1; }
        auto c = std::make_unique<Convolution>(std::string("cudaconv_") + char('1' + (int)n), fs1);
        for (int i = 0; i < 2; i++) {
            const int idx = (int)(2 * n + i);
            auto &cc = c->cc[i];
            if (settings.has("conv[%d].cc.device", idx)) {
                const std::string deviceId = settings.str("conv[%d].cc.device", idx);
                if (!deviceId.empty()) {
                    auto &dev = midiDevices[deviceId];
                    if (!dev) dev = std::make_unique<RawMidi::Device>(deviceId);
                    cc.device = dev.get();
                    dev->handler = c.get();  // one controller per convolution pair, like main.cu:50-51
                }
                cc.message = settings.u8("conv[%d].cc.message", idx);
                cc.select = settings.u8("conv[%d].cc.select", idx);
                cc.predelay = settings.u8("conv[%d].cc.predelay", idx);
                cc.dry = settings.u8("conv[%d].cc.dry", idx);
                cc.wet = settings.u8("conv[%d].cc.wet", idx);
                cc.speed = settings.u8("conv[%d].cc.speed", idx);
                cc.panDry = settings.u8("conv[%d].cc.panDry", idx);
                cc.panWet = settings.u8("conv[%d].cc.panWet", idx);
                cc.level = settings.u8("conv[%d].cc.level", idx);
            }
            auto &v = cc.value;
            v.select = settings.u32("conv[%d].value.select", idx);
            v.predelay = settings.u32("conv[%d].value.predelay", idx);
            v.dry = settings.f32("conv[%d].value.dry", idx);
            v.wet = settings.f32("conv[%d].value.wet", idx);
            v.speed = settings.u32("conv[%d].value.speed", idx);
            v.panDry = settings.f32("conv[%d].value.panDry", idx);
            v.panWet = settings.f32("conv[%d].value.panWet", idx);
            v.level = settings.f32("conv[%d].value.level", idx);
            std::ifstream index(settings.str("conv[%d].index", idx));
            std::string wav;
            for (size_t j = 0; std::getline(index, wav); j++) {
                if (wav.empty()) continue;
                WavFile w(wav, irRate);
                if (!w.error.empty()) { fprintf(stderr, "ca_live: %s\n", w.error.c_str()); return 1; }
                c->prepare(j, w);
            }
        }
        c->start();
        if (!c->isRunning()) { fprintf(stderr, "ca_live: cannot start JACK client %s (is jackd running?)\n", c->name.c_str()); return 1; }
        for (int i = 0; i < 2; i++) {
            const int idx = (int)(2 * n + i);
            jack_connect(c->handle, settings.str("conv[%d].input", idx).c_str(), jack_port_name(c->capture[i]));
            jack_connect(c->handle, jack_port_name(c->playback[i]), settings.str("conv[%d].output", idx).c_str());
            RawMidi::Device *d = c->cc[i].device;
            if (d && !d->isOpen && !d->start()) fprintf(stderr, "ca_live: %s (continuing without MIDI control)\n", d->error.c_str());
        }
        inst.push_back(std::move(c));
    }
    printf("ca_live: %zu instance(s) running; press Enter to stop\n", inst.size());
    std::cin.get();
    for (auto &kv : midiDevices) kv.second->stop();
    for (auto &c : inst) {
        if (c->isRunning()) c->stop();
        printf("%s: average convolution runtime %.4f ms\n", c->name.c_str(), c->avgRuntime());
    }
    return 0;
}
