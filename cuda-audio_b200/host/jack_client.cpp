#include "jack_client.h"

#include <unistd.h>

// jackclient.cu:4-11: the callback returns 0 whatever onProcess did
int JackClient::processTrampoline(jack_nframes_t nframes, void *self)
{
    static_cast<JackClient *>(self)->onProcess(nframes);
    return 0;
}

void JackClient::shutdownTrampoline(void *self)
{
    auto *jc = static_cast<JackClient *>(self);
    Log::warn(jc->name, "JACK is shutting down");
    jc->onShutdown();
}

JackPort JackClient::addPort(const std::string &port, const std::string &type, unsigned long flags, size_t bufferSize)
{
    JackPort p = jack_port_register(handle, port.c_str(), type.c_str(), flags, bufferSize);
    if (!p) { Log::error(name, "cannot register port %s", port.c_str()); return nullptr; }
    ports[port] = p;
    return p;
}

// jackclient.cu:24-44
void JackClient::start()
{
    jack_status_t status = (jack_status_t)0;
    handle = jack_client_open(name.c_str(), JackNoStartServer, &status, nullptr);
    if (!handle || (status & JackNameNotUnique)) { Log::error(name, "cannot open JACK client"); handle = nullptr; return; }
    jack_set_process_callback(handle, processTrampoline, this);
    jack_on_shutdown(handle, shutdownTrampoline, this);
    samplerate = jack_get_sample_rate(handle);
    Log::info(name, "Samplerate: %zu", samplerate);
    _running = true;
    onStart();
}

// jackclient.cu:46-55 (without the reference's 0.5 s sleep)
void JackClient::stop()
{
    if (!handle || !_running) return;
    onStop();
    jack_client_close(handle);
    handle = nullptr;
    _running = false;
}
