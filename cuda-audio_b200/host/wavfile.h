// wavfile.h -- RIFF/WAVE I/O for the convolution engine.
//
// `WavFile` keeps the reference's surface (src/wav.h:5-13): path, numFrames and a DEVICE
// float2* buffer of (L, R) frames at the reference's HALF-scale convention (int16 / 65536,
// int24 / 2^24; wav.cu:13-14,24-41), so Convolution::prepare(idx, wav) is source compatible.
// Unlike the reference it walks the chunk list (skips LIST / cue / fact ..., finds `fmt ` and
// `data` wherever they are), accepts mono (duplicated to both channels), PCM 16/24/32 and
// IEEE float32, decodes on the host (KBs..MBs, one-off) and reports errors instead of asserting.
// `WavData` / wav_read / wav_write are the planar host-side helpers the headless harness uses.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

struct WavData {
    uint32_t sampleRate = 0;
    uint16_t channels = 0, bitsPerSample = 0, audioFormat = 0;
    size_t frames = 0;
    std::vector<std::vector<float>> ch;  // planar, full scale [-1, 1)
    std::string error;                   // empty on success
    bool ok() const { return error.empty(); }
};

// scale: 1.0 = full scale; 0.5 = the reference's IR convention
WavData wav_read(const std::string &path, float scale = 1.0f);
// Band-limited sample-rate conversion (Kaiser-windowed sinc, 32 zero crossings per side, about -100 dB
// stop band, evaluated in fp64): the shipped IR library is 44.1 kHz while the x86 script runs JACK at
// 48 kHz (run_x64_86.sh:4); the reference plays such IRs 8.8 % fast / sharp.  Opt-in (`resample = 1`
// in the settings file, --resample on the harness): off, the reference's behaviour is kept.
// Gain 1 at DC; output frames = ceil(frames * toRate / fromRate).
WavData wav_resample(const WavData &in, uint32_t toRate);
// bits: 16 / 24 (PCM) or 32 (IEEE float)
bool wav_write(const std::string &path, const std::vector<std::vector<float>> &planar, uint32_t sampleRate, int bits = 32);

class WavFile {
public:
    std::string path;
    size_t numFrames = 0;
    float2 *buffer = nullptr;  // device, (L, R) per frame, half scale
    uint32_t sampleRate = 0;
    std::string error;

    // resampleTo: 0 = keep the file's rate (reference behaviour), else convert to that rate on load
    explicit WavFile(const std::string &path, uint32_t resampleTo = 0);
    // from planar host data (no file): used by the harness for synthetic IRs
    WavFile(const float *left, const float *right, size_t frames, uint32_t sampleRate = 48000);
    ~WavFile();
    WavFile(const WavFile &) = delete;
    WavFile &operator=(const WavFile &) = delete;

private:
    void upload(const float *left, const float *right, size_t frames);
};
