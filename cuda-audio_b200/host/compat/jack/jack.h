/* JACK client API subset used by the Convolution host class (same calls the reference makes,
 * jackclient.cu:24-55, conv.cu:291-294, main.cu:86-89).  libjack headers are not installed in
 * this image; a real deployment includes <jack/jack.h> and links -ljack, the headless harness
 * links headless_jack.cpp instead.  Declarations follow the public JACK API documentation. */
#ifndef CA_COMPAT_JACK_H
#define CA_COMPAT_JACK_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct _jack_port jack_port_t;
typedef struct _jack_client jack_client_t;
typedef uint32_t jack_nframes_t;
typedef float jack_default_audio_sample_t;
typedef enum { JackNullOption = 0x00, JackNoStartServer = 0x01 } jack_options_t;
typedef enum { JackFailure = 0x01, JackNameNotUnique = 0x04, JackServerStarted = 0x08 } jack_status_t;
enum JackPortFlags { JackPortIsInput = 0x1, JackPortIsOutput = 0x2 };
#define JACK_DEFAULT_AUDIO_TYPE "32 bit float mono audio"
typedef int (*JackProcessCallback)(jack_nframes_t nframes, void *arg);
typedef void (*JackShutdownCallback)(void *arg);
void *jack_port_get_buffer(jack_port_t *port, jack_nframes_t nframes);
jack_port_t *jack_port_register(jack_client_t *client, const char *port_name, const char *port_type,
                                unsigned long flags, unsigned long buffer_size);
int jack_activate(jack_client_t *client);
jack_client_t *jack_client_open(const char *client_name, jack_options_t options, jack_status_t *status, ...);
int jack_set_process_callback(jack_client_t *client, JackProcessCallback cb, void *arg);
void jack_on_shutdown(jack_client_t *client, JackShutdownCallback cb, void *arg);
jack_nframes_t jack_get_sample_rate(jack_client_t *client);
jack_nframes_t jack_get_buffer_size(jack_client_t *client);
int jack_client_close(jack_client_t *client);
int jack_connect(jack_client_t *client, const char *source_port, const char *destination_port);
const char *jack_port_name(const jack_port_t *port);
#ifdef __cplusplus
}
#endif
#endif
