// jack_client.h -- JACK client base class with the reference's surface (src/jackclient.h:10-63):
// open a client, register ports, route the real-time process callback to a virtual
// onProcess(nframes).  Written against the JACK C API; link real libjack or the in-process
// headless shim (headless_jack.cpp) that the wav-in/wav-out harness uses.
#pragma once
#include <cassert>
#include <cstddef>
#include <map>
#include <string>

#include <jack/jack.h>

#include "logger.h"

typedef jack_port_t *JackPort;

class JackClient {
public:
    const std::string name;
    jack_client_t *handle = nullptr;
    size_t samplerate = 0;
    std::map<std::string, JackPort> ports;

    explicit JackClient(const std::string &name) : name(name) {}
    // a client that is destroyed while running is closed first: JACK must not call into a dead object
    // (the reference's class has no destructor and keeps the callback registered)
    virtual ~JackClient() { if (handle) jack_client_close(handle); }

    void start();
    void stop();
    bool isRunning() const { return _running; }

protected:
    JackPort addInput(const std::string &port, const std::string &type = JACK_DEFAULT_AUDIO_TYPE, size_t bufferSize = 0) { return addPort(port, type, JackPortIsInput, bufferSize); }
    JackPort addOutput(const std::string &port, const std::string &type = JACK_DEFAULT_AUDIO_TYPE, size_t bufferSize = 0) { return addPort(port, type, JackPortIsOutput, bufferSize); }
    void activate() { jack_activate(handle); Log::info(name, "Activated."); }

    virtual void onStart() {}
    virtual void onStop() {}
    virtual void onProcess(size_t nframes) = 0;
    virtual void onShutdown() {}

private:
    JackPort addPort(const std::string &port, const std::string &type, unsigned long flags, size_t bufferSize);
    static int processTrampoline(jack_nframes_t nframes, void *self);
    static void shutdownTrampoline(void *self);
    bool _running = false;
};
