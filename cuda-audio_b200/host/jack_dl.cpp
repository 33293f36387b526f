// jack_dl.cpp -- the JACK client API resolved at run time (dlopen("libjack.so.0")), so the live
// executable builds and links on machines without JACK headers or libraries (this image has
// neither) and talks to a real jackd where one exists.  Every entry point forwards to the symbol
// of the same name; if libjack is missing, jack_client_open() returns NULL with JackFailure and the
// JackClient reports "cannot open JACK client" instead of aborting.
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>

#include <jack/jack.h>

namespace {
void *lib()
{
    static void *h = [] {
        void *p = dlopen("libjack.so.0", RTLD_NOW | RTLD_GLOBAL);
        if (!p) p = dlopen("libjack.so", RTLD_NOW | RTLD_GLOBAL);
        if (!p) fprintf(stderr, "E [jack] libjack not found: %s\n", dlerror());
        return p;
    }();
    return h;
}
template <class F>
F sym(const char *name)
{
    return lib() ? reinterpret_cast<F>(dlsym(lib(), name)) : nullptr;
}
}  // namespace

extern "C" {

void *jack_port_get_buffer(jack_port_t *port, jack_nframes_t nframes)
{
    static auto f = sym<void *(*)(jack_port_t *, jack_nframes_t)>("jack_port_get_buffer");
    return f ? f(port, nframes) : nullptr;
}
jack_port_t *jack_port_register(jack_client_t *c, const char *name, const char *type, unsigned long flags, unsigned long size)
{
    static auto f = sym<jack_port_t *(*)(jack_client_t *, const char *, const char *, unsigned long, unsigned long)>("jack_port_register");
    return f ? f(c, name, type, flags, size) : nullptr;
}
int jack_activate(jack_client_t *c)
{
    static auto f = sym<int (*)(jack_client_t *)>("jack_activate");
    return f ? f(c) : -1;
}
jack_client_t *jack_client_open(const char *name, jack_options_t options, jack_status_t *status, ...)
{
    static auto f = sym<jack_client_t *(*)(const char *, jack_options_t, jack_status_t *, ...)>("jack_client_open");
    if (!f) { if (status) *status = JackFailure; return nullptr; }
    return f(name, options, status);
}
int jack_set_process_callback(jack_client_t *c, JackProcessCallback cb, void *arg)
{
    static auto f = sym<int (*)(jack_client_t *, JackProcessCallback, void *)>("jack_set_process_callback");
    return f ? f(c, cb, arg) : -1;
}
void jack_on_shutdown(jack_client_t *c, JackShutdownCallback cb, void *arg)
{
    static auto f = sym<void (*)(jack_client_t *, JackShutdownCallback, void *)>("jack_on_shutdown");
    if (f) f(c, cb, arg);
}
jack_nframes_t jack_get_sample_rate(jack_client_t *c)
{
    static auto f = sym<jack_nframes_t (*)(jack_client_t *)>("jack_get_sample_rate");
    return f ? f(c) : 0;
}

jack_nframes_t jack_get_buffer_size(jack_client_t *c)
{
    static auto f = sym<jack_nframes_t (*)(jack_client_t *)>("jack_get_buffer_size");
    return f ? f(c) : 0;
}
int jack_client_close(jack_client_t *c)
{
    static auto f = sym<int (*)(jack_client_t *)>("jack_client_close");
    return f ? f(c) : -1;
}
int jack_connect(jack_client_t *c, const char *src, const char *dst)
{
    static auto f = sym<int (*)(jack_client_t *, const char *, const char *)>("jack_connect");
    return f ? f(c, src, dst) : -1;
}
const char *jack_port_name(const jack_port_t *p)
{
    static auto f = sym<const char *(*)(const jack_port_t *)>("jack_port_name");
    return f ? f(p) : "";
}

}  // extern "C"
