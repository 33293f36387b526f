// rawmidi.cpp -- raw MIDI device reader + stream parser (see rawmidi.h).
#include "rawmidi.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <poll.h>
#include <unistd.h>

#include "logger.h"

int MidiParser::dataBytes(uint8_t status)
{
    switch (status & 0xF0) {
    case 0x80: case 0x90: case 0xA0: case 0xB0: case 0xE0: return 2;
    case 0xC0: case 0xD0: return 1;
    default: break;
    }
    switch (status) {
    case 0xF1: case 0xF3: return 1;  // MTC quarter frame, song select
    case 0xF2: return 2;             // song position
    default: return 0;               // F6 tune request, F4/F5 undefined
    }
}

void MidiParser::feed(const uint8_t *bytes, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        const uint8_t b = bytes[i];
        if (b >= 0xF8) {  // system real-time: may appear anywhere, even inside another message
            _sink(&b, 1);
            continue;
        }
        if (b & 0x80) {   // status byte: ends whatever was in progress
            if (_sysex) {
                if (b == 0xF7 && !_overflow && _len < sizeof(_buf)) { _buf[_len++] = b; _sink(_buf, _len); }
                _sysex = false; _overflow = false; _len = 0; _need = 0;
                if (b == 0xF7) continue;
            }
            if (b == 0xF7) continue;  // stray end-of-exclusive
            _len = 0;
            if (b == 0xF0) { _sysex = true; _overflow = false; _buf[_len++] = b; _running = 0; continue; }
            if (b < 0xF0) _running = b; else _running = 0;  // system common cancels running status
            _buf[_len++] = b;
            _need = (size_t)dataBytes(b);
            if (!_need) { _sink(_buf, 1); _len = 0; }
            continue;
        }
        // data byte
        if (_sysex) {
            if (_len < sizeof(_buf) - 1) _buf[_len++] = b; else _overflow = true;
            continue;
        }
        if (!_len) {
            if (!_running) continue;  // data without status: ignore
            _buf[_len++] = _running;
            _need = (size_t)dataBytes(_running);
        }
        _buf[_len++] = b;
        if (_len == 1 + _need) { _sink(_buf, _len); _len = 0; }
    }
}

std::string RawMidi::Device::devicePath(const std::string &id)
{
    int card = 0, dev = 0, sub = 0;
    if (sscanf(id.c_str(), "hw:%d,%d,%d", &card, &dev, &sub) >= 2 || sscanf(id.c_str(), "hw:%d,%d", &card, &dev) == 2) {
        char path[64];
        snprintf(path, sizeof(path), "/dev/snd/midiC%dD%d", card, dev);
        return path;
    }
    return id;
}

bool RawMidi::Device::start()
{
    if (isOpen) return true;
    const std::string path = devicePath(id);
    _writable = true;
    _fd = open(path.c_str(), O_RDWR | O_NONBLOCK | O_CLOEXEC);
    if (_fd < 0) { _writable = false; _fd = open(path.c_str(), O_RDONLY | O_NONBLOCK | O_CLOEXEC); }
    if (_fd < 0) {
        error = "cannot open MIDI device " + id + " (" + path + "): " + strerror(errno);
        Log::error("midi", "%s", error.c_str());
        return false;
    }
    isOpen = true;
    isRunning = true;
    _thread = std::thread([this] { run(); });
    return true;
}

void RawMidi::Device::run()
{
    MidiParser parser([this](const uint8_t *msg, size_t len) { inject(msg, len); });
    uint8_t buf[256];
    while (isRunning.load(std::memory_order_acquire)) {
        pollfd p{_fd, POLLIN, 0};
        const int rc = poll(&p, 1, 50);  // wake up every 50 ms to notice stop()
        if (rc < 0 && errno != EINTR) break;
        if (rc <= 0) continue;
        if (p.revents & (POLLERR | POLLNVAL)) break;
        const ssize_t n = read(_fd, buf, sizeof(buf));
        if (n > 0) parser.feed(buf, (size_t)n);
        else if (n == 0) { if (p.revents & POLLHUP) usleep(20000); }  // FIFO without a writer: wait for the next one
        else if (errno != EAGAIN && errno != EINTR) break;
    }
    isRunning = false;
}

void RawMidi::Device::stop()
{
    if (!isOpen) return;
    isRunning = false;
    if (_thread.joinable()) _thread.join();
    if (_fd >= 0) close(_fd);
    _fd = -1;
    isOpen = false;
}

bool RawMidi::Device::send(const uint8_t *data, size_t len)
{
    if (!isOpen || !_writable) return false;
    while (len) {
        const ssize_t n = write(_fd, data, len);
        if (n < 0) { if (errno == EINTR) continue; return false; }
        data += n; len -= (size_t)n;
    }
    return true;
}
