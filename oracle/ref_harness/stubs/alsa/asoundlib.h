/* Minimal stand-in for <alsa/asoundlib.h> (ALSA is not installed in this image): only the
 * rawmidi names midi.h / midi.cu mention.  midi.cu is not compiled into the oracle. */
#pragma once
#include <stddef.h>
#include <sys/types.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct _snd_rawmidi snd_rawmidi_t;
ssize_t snd_rawmidi_read(snd_rawmidi_t *rmidi, void *buffer, size_t size);
int snd_rawmidi_open(snd_rawmidi_t **in_rmidi, snd_rawmidi_t **out_rmidi, const char *name, int mode);
int snd_rawmidi_nonblock(snd_rawmidi_t *rmidi, int nonblock);
int snd_rawmidi_close(snd_rawmidi_t *rmidi);
int snd_rawmidi_drain(snd_rawmidi_t *rmidi);
#ifdef __cplusplus
}
#endif
