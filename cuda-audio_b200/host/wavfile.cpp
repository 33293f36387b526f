#include "wavfile.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "logger.h"

namespace {
uint32_t rd32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
}  // namespace

WavData wav_read(const std::string &path, float scale)
{
    WavData w;
    std::ifstream is(path, std::ifstream::binary);
    if (!is) { w.error = "cannot open " + path; return w; }
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    if (file.size() < 12 || memcmp(file.data(), "RIFF", 4) || memcmp(file.data() + 8, "WAVE", 4)) { w.error = "not a RIFF/WAVE file: " + path; return w; }
    size_t pos = 12, dataPos = 0, dataLen = 0;
    bool haveFmt = false;
    uint16_t blockAlign = 0;
    while (pos + 8 <= file.size()) {
        const unsigned char *ck = file.data() + pos;
        const size_t len = rd32(ck + 4);
        const size_t body = pos + 8;
        if (!memcmp(ck, "fmt ", 4) && len >= 16 && body + 16 <= file.size()) {
            const unsigned char *f = file.data() + body;
            w.audioFormat = rd16(f); w.channels = rd16(f + 2); w.sampleRate = rd32(f + 4);
            blockAlign = rd16(f + 12); w.bitsPerSample = rd16(f + 14);
            if (w.audioFormat == 0xFFFE && len >= 26) w.audioFormat = rd16(f + 24);  // WAVE_FORMAT_EXTENSIBLE sub-format
            haveFmt = true;
        } else if (!memcmp(ck, "data", 4)) {
            dataPos = body;
            dataLen = std::min(len, file.size() - body);  // tolerate truncated / streamed files
            if (haveFmt) break;
        }
        pos = body + len + (len & 1);  // chunks are word aligned
    }
    if (!haveFmt || !dataPos) { w.error = "missing fmt or data chunk: " + path; return w; }
    const int bytes = w.bitsPerSample / 8;
    const bool isFloat = w.audioFormat == 3 && bytes == 4;
    const bool isPcm = w.audioFormat == 1 && (bytes == 2 || bytes == 3 || bytes == 4);
    if (!w.channels || (!isFloat && !isPcm)) { w.error = "unsupported sample format in " + path; return w; }
    // a header whose blockAlign is smaller than one frame would walk the sample loop past the data chunk
    const size_t frameBytes = (size_t)bytes * w.channels;
    if (!blockAlign) blockAlign = (uint16_t)frameBytes;
    if (blockAlign < frameBytes) { w.error = "blockAlign smaller than channels x bytes per sample in " + path; return w; }
    w.frames = dataLen / blockAlign;
    if (w.frames && (w.frames - 1) * blockAlign + frameBytes > dataLen) w.frames--;  // last frame must lie inside the chunk
    w.ch.assign(w.channels, std::vector<float>(w.frames));
    const unsigned char *d = file.data() + dataPos;
    for (size_t n = 0; n < w.frames; n++)
        for (int c = 0; c < w.channels; c++) {
            const unsigned char *s = d + n * blockAlign + (size_t)c * bytes;
            float v;
            if (isFloat) { uint32_t u = rd32(s); memcpy(&v, &u, 4); v *= scale; }
            else if (bytes == 2) v = (float)(int16_t)rd16(s) / 32768.0f * scale;         // == /65536 at scale 0.5 (wav.cu:13)
            else if (bytes == 3) {
                const int32_t i = (int32_t)(((uint32_t)s[0] << 8) | ((uint32_t)s[1] << 16) | ((uint32_t)s[2] << 24)) / 256;  // wav.cu:27-37
                v = (float)i / 8388608.0f * scale;                                          // == /2^24 at scale 0.5 (wav.cu:40)
            } else v = (float)((double)(int32_t)rd32(s) / 2147483648.0) * scale;
            w.ch[c][n] = v;
        }
    return w;
}

namespace {
double bessel_i0(double x)
{
    double sum = 1.0, term = 1.0;
    for (int k = 1; k < 64; k++) {
        term *= (x / (2.0 * k)) * (x / (2.0 * k));
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}
}  // namespace

WavData wav_resample(const WavData &in, uint32_t toRate)
{
    WavData out = in;
    if (!in.ok() || !toRate || !in.sampleRate || toRate == in.sampleRate) return out;
    const double ratio = (double)toRate / (double)in.sampleRate;
    const double fc = std::min(1.0, ratio);        // cutoff relative to the input Nyquist
    const int zc = 32;                             // zero crossings per side (at the cutoff rate)
    const double beta = 10.0, i0b = bessel_i0(beta);
    const double half = zc / fc;                   // half width of the kernel in input samples
    out.sampleRate = toRate;
    out.frames = (size_t)std::ceil((double)in.frames * ratio - 1e-9);
    for (auto &c : out.ch) c.assign(out.frames, 0.0f);
    std::vector<double> acc(in.ch.size());
    for (size_t m = 0; m < out.frames; m++) {
        const double t = (double)m / ratio;        // position in input samples
        const long n0 = (long)std::ceil(t - half), n1 = (long)std::floor(t + half);
        double wsum = 0.0;
        std::fill(acc.begin(), acc.end(), 0.0);
        for (long n = n0; n <= n1; n++) {
            const double d = (double)n - t, u = d / half;
            const double win = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - u * u))) / i0b;
            const double a = M_PI * fc * d;
            const double w = fc * (std::fabs(a) < 1e-12 ? 1.0 : std::sin(a) / a) * win;
            wsum += w;
            if (n < 0 || n >= (long)in.frames) continue;
            for (size_t c = 0; c < in.ch.size(); c++) acc[c] += w * (double)in.ch[c][(size_t)n];
        }
        // normalising by the kernel's own sum makes the DC gain exactly 1 at every output phase
        for (size_t c = 0; c < in.ch.size(); c++) out.ch[c][m] = (float)(acc[c] / wsum);
    }
    return out;
}

bool wav_write(const std::string &path, const std::vector<std::vector<float>> &planar, uint32_t sampleRate, int bits)
{
    if (planar.empty() || (bits != 16 && bits != 24 && bits != 32)) return false;
    const uint16_t channels = (uint16_t)planar.size();
    const size_t frames = planar[0].size();
    const int bytes = bits / 8;
    const uint32_t dataLen = (uint32_t)(frames * channels * bytes);
    std::vector<unsigned char> out(44 + dataLen);
    auto w32 = [&](size_t o, uint32_t v) { out[o] = v & 255; out[o + 1] = (v >> 8) & 255; out[o + 2] = (v >> 16) & 255; out[o + 3] = (v >> 24) & 255; };
    auto w16 = [&](size_t o, uint16_t v) { out[o] = v & 255; out[o + 1] = (v >> 8) & 255; };
    memcpy(&out[0], "RIFF", 4); w32(4, 36 + dataLen); memcpy(&out[8], "WAVEfmt ", 8); w32(16, 16);
    w16(20, bits == 32 ? 3 : 1); w16(22, channels); w32(24, sampleRate); w32(28, sampleRate * channels * bytes);
    w16(32, (uint16_t)(channels * bytes)); w16(34, (uint16_t)bits); memcpy(&out[36], "data", 4); w32(40, dataLen);
    unsigned char *d = out.data() + 44;
    for (size_t n = 0; n < frames; n++)
        for (int c = 0; c < channels; c++, d += bytes) {
            const float v = planar[c][n];
            if (bits == 32) { uint32_t u; memcpy(&u, &v, 4); d[0] = u & 255; d[1] = (u >> 8) & 255; d[2] = (u >> 16) & 255; d[3] = (u >> 24) & 255; }
            else {
                const double full = bits == 16 ? 32768.0 : 8388608.0;
                long q = lrint((double)v * full);
                if (q > (long)full - 1) q = (long)full - 1;
                if (q < -(long)full) q = -(long)full;
                d[0] = q & 255; d[1] = (q >> 8) & 255;
                if (bits == 24) d[2] = (q >> 16) & 255;
            }
        }
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

void WavFile::upload(const float *left, const float *right, size_t frames)
{
    numFrames = frames;
    if (!frames) return;
    std::vector<float2> host(frames);
    for (size_t n = 0; n < frames; n++) host[n] = make_float2(left[n], right[n]);
    if (cudaMalloc(&buffer, frames * sizeof(float2)) != cudaSuccess ||
        cudaMemcpy(buffer, host.data(), frames * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        error = std::string("CUDA: ") + cudaGetErrorString(cudaGetLastError());
        if (buffer) cudaFree(buffer);
        buffer = nullptr;
        numFrames = 0;
    }
}

WavFile::WavFile(const std::string &path, uint32_t resampleTo) : path(path)
{
    WavData w = wav_read(path, 0.5f);  // half scale like wav.cu
    if (!w.ok()) { error = w.error; Log::error("wav", "%s", error.c_str()); return; }
    if (resampleTo && w.sampleRate && resampleTo != w.sampleRate) {
        Log::info("wav", "resampling %u -> %u Hz: %s", w.sampleRate, resampleTo, path.c_str());
        w = wav_resample(w, resampleTo);
    }
    sampleRate = w.sampleRate;
    Log::info("wav", "IR [%0.2f s] %s", w.sampleRate ? (double)w.frames / w.sampleRate : 0.0, path.c_str());
    const std::vector<float> &L = w.ch[0], &R = w.channels > 1 ? w.ch[1] : w.ch[0];
    upload(L.data(), R.data(), w.frames);
}

WavFile::WavFile(const float *left, const float *right, size_t frames, uint32_t sampleRate) : path("<memory>"), sampleRate(sampleRate)
{
    upload(left, right ? right : left, frames);
}

WavFile::~WavFile()
{
    if (buffer) cudaFree(buffer);
}
