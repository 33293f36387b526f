/*
 * cuda_audio_b200.h -- C ABI of the B200-native partitioned-convolution reverb engine.
 *
 * This is the drop-in boundary for the hot path of limitz/cuda-audio: everything the
 * reference's `Convolution` class does on the GPU (src/conv.h:30-86, src/conv.cu:142-466)
 * sits behind these entry points.  Plain pointers and sizes only; no C++/torch types.
 * The C++ mirror of the reference class (cuda-audio_b200/host/convolution.h) is a thin
 * wrapper over this ABI; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Reference interface each entry point replaces:
 *   ca_create            Convolution::Convolution(name, fftSize)       conv.cu:142-195
 *   ca_load_ir[_device]  Convolution::prepare(idx, wav, nframes)        conv.cu:207-253
 *                        (+ WavFile PCM decode, wav.cu:4-118, done by the host shim)
 *   ca_set_params        writes to Convolution::cc[i].value             conv.h:40-50
 *                        (what handleCC / main.cu:63-70 do)             conv.cu:255-276
 *   ca_process           Convolution::onProcess(nframes)                conv.cu:287-466
 *                        with JACK's planar float host buffers          conv.cu:291-294
 *   ca_get_stats         Convolution::avgRuntime()                      conv.h:61
 *   ca_destroy           (the reference never frees, conv.h:53-54)
 *
 * One engine = `n_instances` independent convolution instances that share a geometry
 * (period, IR capacity, n_in x n_out) and are processed by ONE set of kernel launches per
 * period.  One reference `Convolution` object == one instance with n_in = n_out = 2
 * ("true stereo": outL = IN1*h1.L + IN2*h2.L, outR = IN1*h1.R + IN2*h2.R, conv.cu:392-401).
 *
 * All functions return 0 (CA_OK) or a negative error code; none aborts.
 * Thread model: ca_process* from one thread per engine; ca_set_params from any thread
 * (lock-free hand-off, applied at the next ca_process*); ca_load_ir* may block.
 * Device: every entry point that touches the GPU makes the engine's device (ca_config.device) the calling
 * thread's current CUDA device and leaves it so; callers with CUDA work of their own on another device
 * re-select theirs afterwards.
 */
#ifndef CUDA_AUDIO_B200_H
#define CUDA_AUDIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CA_API_VERSION 1
#define CA_MAX_TIERS 4
#define CA_MAX_PREDELAY 8192 /* CONV_MAX_PREDELAY, conv.h:26-28 */
#define CA_MAX_SPEED 1024    /* CONV_MAX_SPEED,    conv.h:22-24 */

enum ca_error {
    CA_OK = 0,
    CA_ERR_INVALID = -1,     /* bad argument / config */
    CA_ERR_CUDA = -2,        /* CUDA runtime failure (see ca_last_error_string) */
    CA_ERR_NOMEM = -3,
    CA_ERR_STATE = -4,       /* e.g. select points at an IR slot that was never loaded */
    CA_ERR_UNSUPPORTED = -5,
    CA_ERR_PERIOD = -6       /* nframes != configured period */
};

enum ca_flags {
    CA_FLAG_GRAPH = 1u << 0,      /* replay one CUDA graph per period instead of 3 launches   */
    CA_FLAG_STREAMING = 1u << 1,  /* working set >> L2: evict-first hints on spectra loads     */
    CA_FLAG_L2_PERSIST = 1u << 2, /* pin IR spectra + FDL in L2 (access-policy window)         */
    CA_FLAG_PROFILE = 1u << 3,    /* record CUDA events around every kernel (ca_get_stats)     */
    CA_FLAG_RAW_WET = 1u << 4,    /* output the unclamped wet signal only: for partition-range
                                   * shards whose partial outputs are summed before clamp + dry */
    CA_FLAG_ASYNC_TIERS = 1u << 5,/* non-uniform partitioning, latency schedule: every long tier starts one
                                   * period later in the IR (offset >= block + period; ca_config_auto_tiers
                                   * plans it), so its result is due two periods after its block closes and
                                   * the tier work runs on a low-priority stream BESIDE the next period's
                                   * output path instead of in front of it (flat p99) */
    CA_FLAG_PERSISTENT = 1u << 7, /* latency schedule without launches: ONE resident cooperative kernel polls a
                                   * mailbox in mapped host memory and processes every period of a single,
                                   * uniformly partitioned instance (period <= 256); ca_process only writes the
                                   * block + parameters and spins on the answer.  The kernel leaves by itself after
                                   * ~1 s of silence and is relaunched by the next call.  ca_process only. */
    CA_FLAG_REF_QUIRKS = 1u << 8, /* reference-compatible DC / Nyquist bins: the reference's two-for-one FFT
                                   * split mis-unpacks bin 0 and never writes bin N/2 (conv.cu:47-73); the exact engine
                                   * matches it only for DC/Nyquist-free IRs.  With this flag the engine adds the
                                   * per-block rank-1 terms those two bins produce in an N-point reference
                                   * (N = ca_config.ref_fft_size), so its output equals conv.cu's on ANY impulse
                                   * response.  True stereo only. */
    CA_FLAG_LEGACY_FFT = 1u << 6  /* A/B: transforms on the warp-shuffle / whole-transform-per-CTA FFT kernels
                                   * of round 1 instead of the row-FFT family (same layouts, same results
                                   * to fp32 rounding) */
};

enum ca_schedule {
    CA_SCHED_MAC_PERSISTENT = 1u << 0,   /* FDL MAC of every tier on the persistent work-list schedule (n_split = 1) */
    CA_SCHED_MAC_PER_ITEM = 1u << 1,     /* never: one CTA per (instance, split, bin tile) */
    CA_SCHED_FUSED_TIER0 = 1u << 2,      /* tier 0 as ONE kernel per instance (forward + MAC + inverse) */
    CA_SCHED_NO_FUSED_TIER0 = 1u << 3,
    CA_SCHED_PIPELINED = 1u << 4,        /* two-lane batch schedule: FFT lanes beside the MAC lane (measured slower end to end) */
    CA_SCHED_NO_PDL = 1u << 5,           /* no programmatic dependent launch between the period's kernels */
    CA_SCHED_ROWS8 = 1u << 6             /* A/B: 256-point row FFTs as 8 x 8 x 4, one row per warp (round 2's first row
                                          * family) instead of 16 x 16, two rows per warp */
};

typedef struct ca_engine ca_engine;

/* ca_config grows at the END between versions of this header.  ca_config_init() writes the LIBRARY's size, ca_create()
 * refuses a struct_size it does not know: build host code against the header that ships with the library. */
typedef struct ca_config {
    uint32_t struct_size;   /* sizeof(ca_config) */
    int32_t device;         /* CUDA device ordinal */
    uint32_t period;        /* frames per ca_process call (JACK nframes); power of two, 32..1024 */
    uint32_t n_instances;   /* instances batched in one engine, 1 .. 1048576 */
    uint32_t n_in, n_out;   /* 1 or 2 each (2,2 = the reference's true-stereo instance) */
    uint32_t max_ir_frames; /* IR capacity L; longer IRs are truncated like conv.cu:239 */
    uint32_t n_ir_slots;    /* size of the IR bank shared by the engine's instances */
    uint32_t flags;         /* ca_flags */
    uint32_t mac_split;     /* partition-range split of the MAC per instance; 0 = auto */
    uint32_t max_voices;    /* IRs that may be audible at once per input during a cross-fade (1..4); 0 = 2.
                             * 1 = an IR `select` change switches hard instead of gliding (conv.cu:15-32) */
    /* partition-range shard for IRs split across GPUs: this engine convolves
     * with partitions [part_begin, part_begin + part_count) of the uniform partitioning only;
     * part_count == 0 means "all". */
    uint32_t part_begin, part_count;
    /* non-uniform partitioning: tier j uses block size tier_block[j] (power of two; tier 0 must
     * equal period, higher tiers 256..16384 and increasing) for tier_parts[j] partitions
     * (0 = cover the rest); n_tiers <= 1 means uniform.  Tier j >= 1 must start at an IR offset
     * >= its block size (sum of the previous tiers' parts*block): its result is then first
     * needed one period after its block completes, so it runs off the output's critical path. */
    uint32_t n_tiers;
    uint32_t tier_block[CA_MAX_TIERS];
    uint32_t tier_parts[CA_MAX_TIERS];
    float sample_rate;      /* only for deadline / xrun accounting; 0 = off */
    /* Cross-fade voices shared by all inputs of the engine (each input owns one voice for good; a second one
     * is only needed while an IR switch glides, conv.cu:15-32).  0 = auto: every voice resident for small
     * engines, n_inputs / 8 shared voices for batches.  When the pool runs dry an IR switch of that input
     * degrades to a hard switch. */
    uint32_t voice_pool;
    /* Schedule choices that change which kernels run (ca_schedule bits; 0 = the engine picks).  The CA_*
     * environment variables of DESIGN.md section 9 override them and exist for development sweeps only. */
    uint32_t schedule;
    uint32_t io_chunks;     /* instance chunks of ca_process's H2D | kernels | D2H pipeline for batches; 0 = 3 */
    /* Batches: give the (memory-bound) MAC lane its own `sm_split` SMs and the (latency-bound) FFT lanes the rest
     * (CUDA green contexts), and run the two-lane pipelined schedule on them.  0 = off. */
    uint32_t sm_split;
    uint32_t ref_fft_size;  /* CA_FLAG_REF_QUIRKS: the reference's fftSize (Convolution::Convolution, conv.cu:142) */
} ca_config;

/* Per-input parameter block == Convolution::CC::value (conv.h:40-50). */
typedef struct ca_params {
    uint32_t select;    /* IR bank slot used by this input                      */
    uint32_t predelay;  /* samples, [0, CA_MAX_PREDELAY); input 0's is used     */
    uint32_t speed;     /* cross-fade / glide length in periods, [0, 1024]      */
    int32_t vsteps;     /* >= 0: (re)start the glide countdown; < 0: leave it   */
    float dry, wet;     /* [0,1] */
    float panDry, panWet; /* [-1,1] */
    float level;        /* [0,1] */
} ca_params;

typedef struct ca_stats {
    uint64_t periods;       /* ca_process* calls so far                                 */
    uint64_t xruns;         /* calls whose host wall time exceeded period / sample_rate */
    double mean_us, p50_us, p99_us, max_us; /* host wall time per ca_process* call      */
    /* CA_FLAG_PROFILE: mean device time per period, CUDA events on the engine's stream:
     * forward R2C, FDL MAC, inverse C2R+mix of tier 0, and the deferred long tiers */
    double fwd_us, mac_us, inv_us, tiers_us, total_us;
    double tier_fwd_us, tier_mac_us, tier_inv_us; /* the long tiers' share of tiers_us by kernel */
    uint64_t gpu_launches;  /* kernels launched by the engine so far                    */
    uint64_t mac_bytes;     /* algorithmic bytes tier 0's FDL MAC streams per period    */
    uint64_t mac_bytes_amortized; /* all tiers, per period (tier j fires every block_j/period) */
    uint32_t partitions;    /* P of tier 0                                             */
    uint32_t mac_split;     /* effective split of tier 0                                */
    uint64_t device_bytes;  /* device memory held by the engine                         */
    uint32_t n_tiers;
    uint32_t tier_block[CA_MAX_TIERS], tier_parts[CA_MAX_TIERS], tier_offset[CA_MAX_TIERS];
    uint32_t tier0_fused;   /* 1: tier 0 runs as one fused kernel; its time is reported in mac_us */
} ca_stats;

int ca_api_version(void);
const char *ca_strerror(int code);
const char *ca_last_error_string(void); /* thread-local detail of the last CUDA failure */

void ca_config_init(ca_config *cfg); /* zero + struct_size + reference defaults (2x2, period 256) */
/* Fill n_tiers / tier_block / tier_parts for cfg->period and cfg->max_ir_frames: every tier's block
 * is `growth` (power of two, 0 = 8) times the previous one, up to max_block (0 = 16384).  Set period, max_ir_frames
 * and flags before calling.  (Growth 4 streams fewer bytes for batches at period 256 but runs more transforms:
 * measured 4 % faster device-resident, equal through ca_process) */
int ca_config_auto_tiers(ca_config *cfg, uint32_t growth, uint32_t max_block);

int ca_create(const ca_config *cfg, ca_engine **out);
int ca_destroy(ca_engine *e);

/* IR bank.  left/right: planar fp32 time-domain IR (right may be NULL when n_out == 1).
 * Host pointers; synchronous (the caller may free the buffers on return, main.cu:78-79). */
int ca_load_ir(ca_engine *e, uint32_t slot, const float *left, const float *right, uint32_t frames);
/* same with device pointers (current device = engine's); synchronous */
int ca_load_ir_device(ca_engine *e, uint32_t slot, const float *d_left, const float *d_right, uint32_t frames);
/* device pointer to (L, R) interleaved frames, the layout of WavFile::buffer (wav.h:10) */
int ca_load_ir_interleaved_device(ca_engine *e, uint32_t slot, const float *d_lr, uint32_t frames);

int ca_set_params(ca_engine *e, uint32_t instance, uint32_t input, const ca_params *p);
int ca_get_params(ca_engine *e, uint32_t instance, uint32_t input, ca_params *p);
/* jump the wet glide of (instance, input) to `g` (tests: skip the 1-0.8^k fade-in) */
int ca_set_glide(ca_engine *e, uint32_t instance, uint32_t input, float g);

/* Only the first n instances are processed by ca_process* (default: all).  Not a real-time call.
 * Instances that become active again restart like new ones: voices, delay lines and pending tier
 * output are reset (no replay of pre-deactivation audio); re-apply ca_set_glide to skip their fade-in. */
int ca_set_active(ca_engine *e, uint32_t n);

/* Every active instance restarts like a new one: the history in the delay lines is dropped, the wet glide starts from
 * silence again (the state of a fresh reference object, conv.cu:142-195); IR bank and parameters stay.  Not a
 * real-time call.  The host mirror calls it after its silent warm-up period. */
int ca_reset(ca_engine *e);

/* One period for every active instance.
 *   in : host, planar [instance][input ][nframes] fp32, contiguous
 *   out: host, planar [instance][output][nframes] fp32, contiguous
 * Synchronous: out is complete on return (like conv.cu:455).  Pinned buffers are used
 * in place; pageable ones are staged through the engine's pinned buffers. */
int ca_process(ca_engine *e, const float *in, float *out, uint32_t nframes);

/* Same with device-resident buffers; asynchronous on the engine's stream.
 * ca_sync waits for it. */
int ca_process_device(ca_engine *e, const float *d_in, float *d_out, uint32_t nframes);
int ca_sync(ca_engine *e);

/* The engine's CUDA stream (cudaStream_t) so callers can order their own work / events. */
void *ca_stream(ca_engine *e);

int ca_get_stats(ca_engine *e, ca_stats *s);
int ca_reset_stats(ca_engine *e);

/* Diagnostics: read-only bandwidth (GB/s) of a `bytes`-sized device buffer swept `iters` times --
 * L2-resident for small sizes, HBM for large ones; the roofline denominator of the single-instance MAC. */
int ca_measure_read_gbs(int device, size_t bytes, int iters, double *gbs);

/* ---- One very long IR split by partition range across the GPUs of one node (BASELINE configs[4]).
 * The reference's only long-IR mechanism is a bigger fftSize on one GPU (conv.cu:239, conv.h:63); here GPU g
 * convolves the same input with partitions [begin_g, begin_g + count_g) (delay-line offset by ring
 * indexing) and the n_out x B partial spectra (4 KB at B = 256) are summed on the first device, which runs
 * the one inverse transform, clamp and dry mix.  Exchange variants:
 *   CA_EXCHANGE_P2P   fused: the last CTA of each peer's MAC kernel stores its summed spectrum into the root's
 *                     memory over NVLink peer access and publishes a flag; the root's inverse kernel waits on
 *                     the flags.  No collective launch; one CUDA graph per GPU per period.
 *   CA_EXCHANGE_NCCL  ncclReduce(sum) of the spectra onto the root, then the inverse (libnccl resolved at run
 *                     time with dlopen; CA_ERR_UNSUPPORTED when it is not installed).
 * One process drives all GPUs; same call sequence and semantics as a 1-instance engine. */
enum ca_exchange { CA_EXCHANGE_P2P = 0, CA_EXCHANGE_NCCL = 1 };

typedef struct ca_group ca_group;

typedef struct ca_group_config {
    uint32_t struct_size;   /* sizeof(ca_group_config) */
    uint32_t n_devices;     /* 1..8; devices[0] is the root (input fan-out, inverse, output) */
    int32_t devices[8];
    uint32_t period, n_in, n_out;
    uint32_t max_ir_frames, n_ir_slots;
    uint32_t flags;         /* ca_flags for the member engines (CA_FLAG_L2_PERSIST, ...); graphs are always used with P2P */
    uint32_t exchange;      /* ca_exchange */
    uint32_t max_voices;
    float sample_rate;
} ca_group_config;

typedef struct ca_group_stats {
    uint64_t periods;
    double mean_us, p50_us, p99_us, max_us;   /* host wall time per ca_group_process call */
    uint32_t n_devices, exchange;
    uint32_t part_begin[8], part_count[8];    /* partition range per device */
    uint32_t mac_split[8];
    uint64_t mac_bytes[8];                    /* algorithmic bytes each device's MAC streams per period */
    uint64_t exchange_bytes_per_peer;         /* bytes each peer sends to the root per period */
    uint64_t gpu_launches;
    int32_t peer_timeout;                     /* 1: the root gave up waiting for a peer's flag */
} ca_group_stats;

void ca_group_config_init(ca_group_config *cfg);
int ca_group_create(const ca_group_config *cfg, ca_group **out);
int ca_group_destroy(ca_group *g);
int ca_group_load_ir(ca_group *g, uint32_t slot, const float *left, const float *right, uint32_t frames);
int ca_group_set_params(ca_group *g, uint32_t input, const ca_params *p);
int ca_group_set_glide(ca_group *g, uint32_t input, float glide);
/* ca_reset on every member: history dropped, wet glide from silence again (after a silent warm-up period). */
int ca_group_reset(ca_group *g);
/* One period: in = planar [n_in][nframes], out = planar [n_out][nframes], host memory; synchronous. */
int ca_group_process(ca_group *g, const float *in, float *out, uint32_t nframes);
int ca_group_get_stats(ca_group *g, ca_group_stats *s);
int ca_group_reset_stats(ca_group *g);

/* Diagnostics, CA_FLAG_PERSISTENT: GPU %globaltimer (ns) at the phase boundaries of the last period: input seen,
 * input read, forward done, first CTA's MAC done, all CTAs arrived, partials summed, inverse done, published. */
int ca_persist_stamps(ca_engine *e, uint64_t stamps_ns[8]);

/* Pinned host memory helpers for callers that want zero staging copies. */
int ca_host_alloc(void **p, size_t bytes);
int ca_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* CUDA_AUDIO_B200_H */
