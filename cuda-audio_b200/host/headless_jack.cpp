#include "headless_jack.h"

#include <cstdarg>
#include <map>
#include <string>
#include <vector>

struct _jack_port {
    std::string name;
    unsigned long flags = 0;
    std::vector<float> own;
    float *ext = nullptr;
};

struct _jack_client {
    std::string name;
    JackProcessCallback process = nullptr;
    void *processArg = nullptr;
    JackShutdownCallback shutdown = nullptr;
    void *shutdownArg = nullptr;
    std::vector<_jack_port *> ports;
    std::map<std::string, std::string> wires;  // own port name -> peer name
    bool active = false;
};

static unsigned g_rate = 48000;
static unsigned g_buffer = 0;  // 0: unknown until the first cycle

extern "C" {

void hj_set_sample_rate(unsigned rate) { g_rate = rate; }
void hj_set_buffer_size(unsigned frames) { g_buffer = frames; }

int hj_cycle(jack_client_t *c, jack_nframes_t nframes)
{
    if (!c || !c->active || !c->process) return -1;
    return c->process(nframes, c->processArg);
}

void hj_port_set_buffer(jack_port_t *p, float *buffer) { if (p) p->ext = buffer; }

const char *hj_port_peer(jack_client_t *c, const jack_port_t *p)
{
    if (!c || !p) return nullptr;
    auto it = c->wires.find(p->name);
    return it == c->wires.end() ? nullptr : it->second.c_str();
}

void *jack_port_get_buffer(jack_port_t *p, jack_nframes_t nframes)
{
    if (!p) return nullptr;
    if (p->ext) return p->ext;
    if (p->own.size() < nframes) p->own.assign(nframes, 0.f);
    return p->own.data();
}

jack_port_t *jack_port_register(jack_client_t *c, const char *port_name, const char *, unsigned long flags, unsigned long)
{
    if (!c) return nullptr;
    auto *p = new _jack_port();
    p->name = c->name + ":" + port_name;
    p->flags = flags;
    c->ports.push_back(p);
    return p;
}

int jack_activate(jack_client_t *c) { if (!c) return -1; c->active = true; return 0; }

jack_client_t *jack_client_open(const char *client_name, jack_options_t, jack_status_t *status, ...)
{
    if (status) *status = (jack_status_t)0;
    auto *c = new _jack_client();
    c->name = client_name;
    return c;
}

int jack_set_process_callback(jack_client_t *c, JackProcessCallback cb, void *arg) { c->process = cb; c->processArg = arg; return 0; }
void jack_on_shutdown(jack_client_t *c, JackShutdownCallback cb, void *arg) { c->shutdown = cb; c->shutdownArg = arg; }
jack_nframes_t jack_get_sample_rate(jack_client_t *) { return g_rate; }
jack_nframes_t jack_get_buffer_size(jack_client_t *) { return g_buffer; }

int jack_client_close(jack_client_t *c)
{
    if (!c) return -1;
    for (auto *p : c->ports) delete p;
    delete c;
    return 0;
}

int jack_connect(jack_client_t *c, const char *src, const char *dst)
{
    if (!c || !src || !dst) return -1;
    const std::string prefix = c->name + ":";
    if (std::string(src).compare(0, prefix.size(), prefix) == 0) c->wires[src] = dst;
    if (std::string(dst).compare(0, prefix.size(), prefix) == 0) c->wires[dst] = src;
    return 0;
}

const char *jack_port_name(const jack_port_t *p) { return p ? p->name.c_str() : ""; }

}  // extern "C"
