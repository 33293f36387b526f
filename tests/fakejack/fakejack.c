/* fakejack.c -- TEST INFRASTRUCTURE: a stand-in for libjack.so.0 + jackd in one shared library, so that the live
 * executable's run-time binding (cuda-audio_b200/host/jack_dl.cpp: dlopen("libjack.so.0")) and its real-time callback
 * path can be exercised on a machine without JACK.  It implements the eleven entry points the reference uses
 * (jackclient.cu:4-55, main.cu:82-89): a client opens, registers ports, connects them to "system:capture_N" /
 * "system:playback_N"; jack_activate() starts the "server" thread, which calls the process callback once per period
 * like jackd's RT thread, feeding the capture ports from FAKEJACK_IN (raw float32, planar [2][frames]) and recording
 * the playback ports into FAKEJACK_OUT; when the input is exhausted it writes the output, creates FAKEJACK_OUT.done
 * and keeps cycling silence.  FAKEJACK_NFRAMES / FAKEJACK_RATE set the period and the sample rate,
 * FAKEJACK_PERIOD_US paces the cycles (0 = back to back).  Per-period callback time statistics go to the .done file. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef unsigned int jack_nframes_t;
typedef int (*JackProcessCallback)(jack_nframes_t, void *);
typedef void (*JackShutdownCallback)(void *);

#define MAX_PORTS 8
typedef struct port {
    char name[160];
    unsigned long flags;
    float *buf;
    int system_channel; /* 0/1 once connected to system:capture_N (inputs) or system:playback_N (outputs), else -1 */
    struct client *owner;
} port;

typedef struct client {
    char name[64];
    JackProcessCallback cb;
    void *cb_arg;
    port ports[MAX_PORTS];
    int n_ports;
    pthread_t thread;
    int active, stop;
} client;

enum { JackPortIsInput = 1, JackPortIsOutput = 2 };

static unsigned env_u(const char *k, unsigned d) { const char *v = getenv(k); return v && *v ? (unsigned)atoi(v) : d; }
static unsigned nframes(void) { return env_u("FAKEJACK_NFRAMES", 256); }

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

static void *server(void *arg)
{
    client *c = (client *)arg;
    const unsigned B = nframes();
    const unsigned pace = env_u("FAKEJACK_PERIOD_US", 0);
    const char *in_path = getenv("FAKEJACK_IN"), *out_path = getenv("FAKEJACK_OUT");
    float *in = NULL, *out = NULL;
    size_t frames = 0;
    if (in_path) {
        FILE *f = fopen(in_path, "rb");
        if (f) {
            fseek(f, 0, SEEK_END);
            const long bytes = ftell(f);
            fseek(f, 0, SEEK_SET);
            frames = (size_t)bytes / (2 * sizeof(float));
            in = (float *)malloc((size_t)bytes);
            if (fread(in, 1, (size_t)bytes, f) != (size_t)bytes) frames = 0;
            fclose(f);
        }
    }
    const size_t periods = frames / B;
    out = (float *)calloc(2 * periods * B + 1, sizeof(float));
    double worst = 0, total = 0;
    size_t t = 0;
    int reported = 0;
    while (!c->stop) {
        for (int i = 0; i < c->n_ports; i++) {
            port *p = &c->ports[i];
            if (p->flags & JackPortIsInput) {
                if (p->system_channel >= 0 && t < periods) memcpy(p->buf, in + (size_t)p->system_channel * frames + t * B, B * sizeof(float));
                else memset(p->buf, 0, B * sizeof(float));
            }
        }
        const double t0 = now_us();
        if (c->cb) c->cb(B, c->cb_arg);
        const double dt = now_us() - t0;
        if (t < periods) {
            total += dt;
            if (dt > worst) worst = dt;
            for (int i = 0; i < c->n_ports; i++) {
                port *p = &c->ports[i];
                if ((p->flags & JackPortIsOutput) && p->system_channel >= 0) memcpy(out + (size_t)p->system_channel * periods * B + t * B, p->buf, B * sizeof(float));
            }
        }
        t++;
        if (t >= periods && !reported) {
            reported = 1;
            if (out_path) {
                FILE *f = fopen(out_path, "wb");
                if (f) { fwrite(out, sizeof(float), 2 * periods * B, f); fclose(f); }
                char done[600];
                snprintf(done, sizeof done, "%s.done", out_path);
                f = fopen(done, "w");
                if (f) { fprintf(f, "{\"periods\": %zu, \"mean_us\": %.2f, \"max_us\": %.2f}\n", periods, periods ? total / periods : 0.0, worst); fclose(f); }
            }
        }
        if (pace) usleep(pace); else if (t >= periods) usleep(2000);
    }
    free(in);
    free(out);
    return NULL;
}

client *jack_client_open(const char *name, int options, int *status, ...)
{
    (void)options;
    if (getenv("FAKEJACK_REFUSE")) { if (status) *status = 1; return NULL; }
    client *c = (client *)calloc(1, sizeof(client));
    snprintf(c->name, sizeof c->name, "%s", name);
    if (status) *status = 0;
    return c;
}

int jack_set_process_callback(client *c, JackProcessCallback cb, void *arg) { c->cb = cb; c->cb_arg = arg; return 0; }
void jack_on_shutdown(client *c, JackShutdownCallback cb, void *arg) { (void)c; (void)cb; (void)arg; }
jack_nframes_t jack_get_sample_rate(client *c) { (void)c; return env_u("FAKEJACK_RATE", 48000); }
jack_nframes_t jack_get_buffer_size(client *c) { (void)c; return nframes(); }

port *jack_port_register(client *c, const char *name, const char *type, unsigned long flags, unsigned long size)
{
    (void)type; (void)size;
    if (c->n_ports >= MAX_PORTS) return NULL;
    port *p = &c->ports[c->n_ports++];
    snprintf(p->name, sizeof p->name, "%s:%s", c->name, name);
    p->flags = flags;
    p->buf = (float *)calloc(8192, sizeof(float));
    p->system_channel = -1;
    p->owner = c;
    return p;
}

void *jack_port_get_buffer(port *p, jack_nframes_t n) { (void)n; return p ? p->buf : NULL; }
const char *jack_port_name(const port *p) { return p ? p->name : ""; }

int jack_activate(client *c)
{
    if (c->active) return 0;
    c->active = 1;
    return pthread_create(&c->thread, NULL, server, c);
}

int jack_connect(client *c, const char *src, const char *dst)
{
    for (int i = 0; i < c->n_ports; i++) {
        port *p = &c->ports[i];
        if ((p->flags & JackPortIsInput) && !strcmp(dst, p->name) && !strncmp(src, "system:capture_", 15)) { p->system_channel = atoi(src + 15) - 1; return 0; }
        if ((p->flags & JackPortIsOutput) && !strcmp(src, p->name) && !strncmp(dst, "system:playback_", 16)) { p->system_channel = atoi(dst + 16) - 1; return 0; }
    }
    return -1;
}

int jack_client_close(client *c)
{
    if (c->active) { c->stop = 1; pthread_join(c->thread, NULL); }
    for (int i = 0; i < c->n_ports; i++) free(c->ports[i].buf);
    free(c);
    return 0;
}
