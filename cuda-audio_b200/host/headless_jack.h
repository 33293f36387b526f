// headless_jack.h -- driver side of the in-process JACK stand-in (headless_jack.cpp).
// The Convolution class talks to the JACK C API exactly like the reference; the headless
// wav-in/wav-out harness plays jackd: it owns the port buffers and calls the registered
// process callback once per period (the RT thread of jackclient.cu:4-11).
#pragma once
#include <jack/jack.h>

#ifdef __cplusplus
extern "C" {
#endif
void hj_set_sample_rate(unsigned rate);
void hj_set_buffer_size(unsigned frames);  /* what jack_get_buffer_size() reports (0 = unknown) */
/* run one period: invokes the client's process callback with nframes */
int hj_cycle(jack_client_t *client, jack_nframes_t nframes);
/* point a port at caller-owned memory for the next cycles (NULL = internal buffer) */
void hj_port_set_buffer(jack_port_t *port, float *buffer);
/* ports connected with jack_connect(): returns the peer's name or NULL */
const char *hj_port_peer(jack_client_t *client, const jack_port_t *port);
#ifdef __cplusplus
}
#endif
