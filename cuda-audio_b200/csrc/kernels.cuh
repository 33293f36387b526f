// kernels.cuh -- sm_100a kernels of the partitioned overlap-save convolution engine.
//
// Per period, per engine (all instances batched):
//   k_forward : predelay ring + wet glide + window [x(t-1) | x(t)] -> R2C (2B) -> FDL slot
//               replaces f_pack2R2C + cufftExecC2C fwd + f_unpackC22R + f_interpolate
//               (conv.cu:15-73, 321-371)
//   k_mac     : Y_o = sum_i pan_io * sum_k X_i[t-k] (.) H_io,k, TMA(bulk)-staged, mbarrier
//               pipelined, split over partition ranges; replaces f_pointwiseMultiplyAndScale
//               over the fftSize-long spectra (conv.cu:102-123, 392-401)
//   k_inverse : sum of partial spectra -> C2R (2B) -> overlap discard -> clamp -> dry mix
//               replaces cufftExecC2C inverse x2 + f_pointwiseAdd + f_addDryInterleaved +
//               residual shift (conv.cu:405-451)
//   k_ir_fft  : IR partitions -> spectra (prepare(), conv.cu:207-253)
//
// Spectra are stored in packed real-FFT format: B complex per partition, bin 0 = (DC, Nyquist).
#pragma once
#include <cstdio>
#include "fft_cta.cuh"

namespace ca {

constexpr uint32_t kMaxPredelay = 8192;  // CONV_MAX_PREDELAY, conv.h:26-28
#ifndef CA_FWD_WARPS
#define CA_FWD_WARPS 4
#endif
constexpr int kFwdWarps = CA_FWD_WARPS;
constexpr int kMaxVoices = 4;

// per (instance, input): written by the host (ca_set_params), read by the kernels
struct InParamDev {
    float wet, dry, level, panWet, panDry;
    uint32_t predelay, select;
    uint32_t vsteps_cmd, vsteps_seq;  // glide countdown command + sequence number
    float glide_cmd;
    uint32_t glide_seq;               // glide-jump command + sequence number
    uint32_t pad;
};

// per (instance, input): device-owned state, double-buffered by period parity.
// The reference glides its LIVE IR spectrum toward wet * (selected IR) once per period
// (f_interpolate, conv.cu:15-32, 339-353): Hlive <- Hlive + (wet*Hsel - Hlive)/(vsteps+5).
// Hlive is therefore always a linear combination sum_j c_j H_j of loaded IRs whose coefficients
// follow the same recurrence, and the whole response of input block t carries Hlive(t).
// A "voice" is one (IR slot, coefficient c, frequency-domain delay line fed with c(t) x(t)):
// exact for time-varying wet and for IR switches, zero extra bytes in steady state.
struct ItemState {
    float c[kMaxVoices];
    uint32_t slot[kMaxVoices];
    unsigned long long start[kMaxVoices];  // period at which the voice's delay lines were (re)started
    uint32_t quiet[kMaxVoices];            // periods since the coefficient reached zero
    uint32_t active;                       // bit v: voice v is audible or still ringing out
    uint32_t fresh;                        // bit v: voice v was (re)allocated in this period
    uint32_t vsteps;
    uint32_t vsteps_seq_seen, glide_seq_seen, pad;
    // storage of voice v (time ring + one delay line per tier): 1 + index into the voice pool, 0 = none.
    // Item i owns pool entry i for good (its "home"); entries >= n_home are shared by all items and only
    // held while a second voice is audible or ringing out (an IR cross-fade), see voice_storage_update().
    uint32_t pool[kMaxVoices];
};

// Shared pool of cross-fade voices.  Outside cross-fades an input needs ONE voice; a resident second delay
// line per input was 22 % of an instance's memory (and HBM capacity, not time, bounds the channel count).
struct VoicePool {
    uint32_t *bitmap;   // bit b of word w: pool entry n_home + 32 w + b is in use
    uint32_t n_words;
    uint32_t n_home;    // = allocated (instance, input) items
    uint32_t n_extra;   // shared entries
};

__device__ __forceinline__ uint32_t voice_pool_alloc(const VoicePool &vp, uint32_t hint)
{
    for (uint32_t i = 0; i < vp.n_words; i++) {
        const uint32_t w = (hint + i) % vp.n_words;
        uint32_t cur = vp.bitmap[w];
        while (~cur) {
            const uint32_t bit = (uint32_t)__ffs((int)~cur) - 1u;
            if (32u * w + bit >= vp.n_extra) break;
            const uint32_t prev = atomicOr(&vp.bitmap[w], 1u << bit);
            if (!(prev & (1u << bit))) return vp.n_home + 32u * w + bit + 1u;
            cur = prev | (1u << bit);
        }
    }
    return 0u;  // pool exhausted
}

__device__ __forceinline__ void voice_pool_free(const VoicePool &vp, uint32_t entry_p1)
{
    if (entry_p1 > vp.n_home) {  // homes are never returned
        const uint32_t e = entry_p1 - 1u - vp.n_home;
        atomicAnd(&vp.bitmap[e >> 5], ~(1u << (e & 31u)));
    }
}

// ONE lane per item, after step_item_state(): release the storage of voices that went silent in this step,
// give every fresh voice storage (the item's home entry when no other voice of the item holds it, else a
// shared entry).  When the shared pool is exhausted the new voice takes over the storage of the item's
// quietest other voice, which is cut off: the IR switch degrades to a hard switch instead of failing.
__device__ __forceinline__ void voice_storage_update(ItemState &s, uint32_t old_active, uint32_t item, int nv, const VoicePool &vp)
{
#pragma unroll
    for (int v = 0; v < kMaxVoices; v++)
        if (v < nv && ((old_active >> v) & 1u) && !((s.active >> v) & 1u)) { voice_pool_free(vp, s.pool[v]); s.pool[v] = 0u; }
#pragma unroll
    for (int v = 0; v < kMaxVoices; v++) {
        if (!(v < nv && ((s.active >> v) & 1u) && s.pool[v] == 0u)) continue;
        bool home_used = false;
#pragma unroll
        for (int u = 0; u < kMaxVoices; u++) home_used |= (u < nv && s.pool[u] == item + 1u);
        uint32_t got = home_used ? voice_pool_alloc(vp, item) : item + 1u;
        if (!got) {
            int q = -1;
            float best = 3.0e38f;
#pragma unroll
            for (int u = 0; u < kMaxVoices; u++)
                if (u < nv && u != v && ((s.active >> u) & 1u) && s.pool[u] && fabsf(s.c[u]) < best) { best = fabsf(s.c[u]); q = u; }
#pragma unroll
            for (int u = 0; u < kMaxVoices; u++)
                if (u == q) { got = s.pool[u]; s.pool[u] = 0u; s.active &= ~(1u << u); }
        }
        s.pool[v] = got;
        s.fresh |= 1u << v;  // whatever the entry held belongs to another voice's past
    }
}

// the whole per-period voice step of one (instance, input) item, warp-wide: every lane computes the pure
// step, lane 0 updates the storage, the result is broadcast and stored for the period's other kernels
__device__ __forceinline__ ItemState item_step_warp(ItemState *st, uint32_t n_items_alloc, uint32_t item, const InParamDev &p, unsigned long long t,
                                                    int nv, uint32_t ring_out, const VoicePool &vp, int lane);
struct Ctl {
    // period counter.  k_forward and k_mac read t; k_forward publishes t_next = t + 1; k_inverse reads
    // only t_next and finally sets t = t_next, so no kernel reads a field another CTA of it writes.
    unsigned long long t, t_next;
    // CA_FLAG_ASYNC_TIERS under a graph: the long tiers of the block that closed at t_end run while the
    // next period advances t, so they read t_end from t_def[t_end & 1] (written together with t)
    unsigned long long t_def[2];
};

// period count for the deferred kernels: host argument (host-driven launches), the parity slot
// (t_sel = 1 + (t_end & 1), asynchronous tiers inside a graph) or the live counter
__device__ __forceinline__ unsigned long long ctl_tend(const Ctl *ctl, unsigned long long tend_host, uint32_t t_sel, uint32_t bias)
{
    if (tend_host) return tend_host;
    return t_sel ? ctl->t_def[t_sel - 1u] : ctl->t + bias;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may become resident while the previous kernel of the stream drains; pdl_wait() blocks until
// that kernel has completed and its writes are visible, so everything after it sees normal stream
// order.  Both are no-ops for ordinary launches.  Hides the launch gap and CTA ramp between the
// period's dependent kernels (2.5-5 us each, `CA_PIPE_TRACE` timeline).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// pan law, conv.cu:386-389 / 418-421
__device__ __forceinline__ float pan_gain(float pan, int o, int n_out)
{
    if (n_out == 1) return 1.0f;
    return o == 0 ? (pan >= 0.f ? 1.f - pan : 1.f) : (pan <= 0.f ? 1.f + pan : 1.f);
}

// One glide step (start of period t) for all voices of one (instance, input).
__device__ __forceinline__ ItemState step_item_state(ItemState s, const InParamDev &p, unsigned long long t, int nv, uint32_t ring_out)
{
    s.fresh = 0;
    if (s.glide_seq_seen != p.glide_seq) {  // jump: one voice at coefficient glide_cmd, history dropped
        s.glide_seq_seen = p.glide_seq;
        s.active = 1u; s.fresh = 1u;
        s.slot[0] = p.select; s.c[0] = p.glide_cmd; s.start[0] = t; s.quiet[0] = 0;
    }
    if (s.vsteps_seq_seen != p.vsteps_seq) { s.vsteps = p.vsteps_cmd; s.vsteps_seq_seen = p.vsteps_seq; }
    int tv = -1;
#pragma unroll
    for (int v = 0; v < kMaxVoices; v++)
        if (v < nv && ((s.active >> v) & 1u) && s.slot[v] == p.select) tv = v;
    if (tv < 0) {  // new target IR: take a free voice, else steal the quietest one
        float best = 3.0e38f;
#pragma unroll
        for (int v = 0; v < kMaxVoices; v++)
            if (v < nv) {
                const float score = ((s.active >> v) & 1u) ? fabsf(s.c[v]) : -1.0f;
                if (score < best) { best = score; tv = v; }
            }
#pragma unroll
        for (int v = 0; v < kMaxVoices; v++)
            if (v == tv) { s.slot[v] = p.select; s.c[v] = 0.f; s.start[v] = t; s.quiet[v] = 0; }
        s.active |= 1u << tv;
        s.fresh |= 1u << tv;
    }
    const float d = (float)(s.vsteps + 5u);
#pragma unroll
    for (int v = 0; v < kMaxVoices; v++)
        if (v < nv && ((s.active >> v) & 1u)) {
            const float target = (v == tv) ? p.wet : 0.f;
            s.c[v] = s.c[v] + (target - s.c[v]) / d;  // conv.cu:27
            if (v != tv) {
                if (fabsf(s.c[v]) < 1e-9f) {  // inaudible (< -180 dB): stop feeding, let the delay line ring out
                    s.c[v] = 0.f;
                    if (++s.quiet[v] > ring_out) s.active &= ~(1u << v);
                } else s.quiet[v] = 0;
            }
        }
    if (s.vsteps > 0) s.vsteps--;  // conv.cu:345,353
    return s;
}

__device__ __forceinline__ ItemState item_step_warp(ItemState *st, uint32_t n_items_alloc, uint32_t item, const InParamDev &p, unsigned long long t,
                                                    int nv, uint32_t ring_out, const VoicePool &vp, int lane)
{
    const ItemState old = st[(t & 1ull) * n_items_alloc + item];
    ItemState s = step_item_state(old, p, t, nv, ring_out);
    if (lane == 0) voice_storage_update(s, old.active, item, nv, vp);
    s.active = __shfl_sync(0xffffffffu, s.active, 0);
    s.fresh = __shfl_sync(0xffffffffu, s.fresh, 0);
#pragma unroll
    for (int v = 0; v < kMaxVoices; v++) s.pool[v] = __shfl_sync(0xffffffffu, s.pool[v], 0);
    if (lane == 0) st[((t + 1ull) & 1ull) * n_items_alloc + item] = s;
    return s;
}

// ca_set_active: items that are about to be reset hand their shared voice entries back
__global__ void k_release_voices(const ItemState *st, uint32_t n, const VoicePool vp)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int v = 0; v < kMaxVoices; v++) voice_pool_free(vp, st[i].pool[v]);
}

// same with the old state already loaded by the caller (so that load can overlap other set-up work)
__device__ __forceinline__ ItemState item_step_warp_from(const ItemState &old, ItemState *st, uint32_t n_items_alloc, uint32_t item, const InParamDev &p,
                                                         unsigned long long t, int nv, uint32_t ring_out, const VoicePool &vp, int lane)
{
    ItemState s = step_item_state(old, p, t, nv, ring_out);
    if (lane == 0) voice_storage_update(s, old.active, item, nv, vp);
    s.active = __shfl_sync(0xffffffffu, s.active, 0);
    s.fresh = __shfl_sync(0xffffffffu, s.fresh, 0);
#pragma unroll
    for (int v = 0; v < kMaxVoices; v++) s.pool[v] = __shfl_sync(0xffffffffu, s.pool[v], 0);
    if (lane == 0) st[((t + 1ull) & 1ull) * n_items_alloc + item] = s;
    return s;
}

// pool entry (0-based) of voice v; only called for voices that are active (they always have storage)
__device__ __forceinline__ uint32_t voice_entry(const ItemState &s, uint32_t v)
{
    uint32_t e = 0;
#pragma unroll
    for (int q = 0; q < kMaxVoices; q++) e = (q == (int)v) ? s.pool[q] : e;
    return e - 1u;
}

// ------------------------------------------------------------------------------------------
// forward (tier 0): one warp per (instance, input, voice)
// ------------------------------------------------------------------------------------------
struct FwdArgs {
    const float *in;   // [inst][n_in][B]
    float *ring;       // [voice pool entry][ring_len]: predelayed, gain-scaled input stream
    float2 *X;         // FDL [voice pool entry][Lring][B]
    const InParamDev *par;
    ItemState *st;     // [2][inst*n_in]
    Ctl *ctl;
    const float2 *twM, *tw2M;
    uint32_t n_items;  // active instances * n_in
    uint32_t n_items_alloc;  // stride between the two state buffers
    uint32_t n_in, nv, Lring, ring_len, ring_out;
    uint32_t item0;    // first (instance, input) item of this launch (chunked host pipeline)
    // host-driven launches (no graph): the period count + 1 comes as an argument, which takes the
    // ctl->t round trip off the head of every warp's chain of dependent loads (0: read ctl->t)
    unsigned long long t_host_p1;
    const float2 *rowtw;  // [W_256^n | W_512^k] for the row-FFT kernels (kernels_rows.cuh)
    VoicePool vp;
};

// One warp per (instance, input); it steps the voice state once and then runs every audible voice
// (one in steady state, two or three during an IR cross-fade).
template <int R>
__global__ void __launch_bounds__(kFwdWarps * 32, 20 / kFwdWarps) k_forward(const FwdArgs a)
{
    pdl_trigger();  // the next kernel of the stream may become resident while this one drains
    pdl_wait();     // before any global access and before any early exit: stream order holds transitively
    constexpr int B = 32 * R;
    const int lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    if (w >= a.n_items) return;
    const uint32_t item = a.item0 + w;
    const unsigned long long t = a.t_host_p1 ? a.t_host_p1 - 1ull : a.ctl->t;
    if (w == 0 && lane == 0) a.ctl->t_next = t + 1ull;

    const InParamDev p = a.par[item];
    const uint32_t pd = a.par[(item / a.n_in) * a.n_in].predelay;  // input 0's, conv.cu:412,415
    const ItemState s = item_step_warp(a.st, a.n_items_alloc, item, p, t, (int)a.nv, a.ring_out, a.vp, lane);

    WarpFft<R> f;
    f.init(a.twM);
    const uint32_t mask = a.ring_len - 1;
    const float *x = a.in + (size_t)item * B;
    const uint32_t base = (uint32_t)((t * (unsigned long long)B) & mask);
    const uint32_t prev = (base - B) & mask;
    const uint32_t off = ((lane < 16) ? prev : base) + 2 * R * (lane & 15);  // time layout: lane a holds floats [2R a, 2R a + 2R)
    const unsigned long long n_fire = t + 1ull;                               // tier 0 fires every period
    const uint32_t slot = (a.Lring - 1u) - (uint32_t)(n_fire % a.Lring);      // the FDL ring runs backwards

#pragma unroll 1
    for (uint32_t v = 0; v < a.nv; v++) {
        if (!((s.active >> v) & 1u)) continue;
        float cv = 0.f;
#pragma unroll
        for (int q = 0; q < kMaxVoices; q++) cv = (q == (int)v) ? s.c[q] : cv;
        const float gain = cv * p.level;
        const uint32_t entry = voice_entry(s, v);
        float *ring = a.ring + (size_t)entry * a.ring_len;
        if ((s.fresh >> v) & 1u) {  // (re)allocated voice: its time-domain history belongs to another IR
            for (uint32_t n = 4 * lane; n < a.ring_len; n += 128) *reinterpret_cast<float4 *>(ring + n) = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < R; j++)  // clear the block that becomes reachable by the predelay scatter in this period
            ring[(base + kMaxPredelay + lane + 32 * j) & mask] = 0.f;

        float2 z[R];
        if (pd == 0) {
            // no predelay: the block lands exactly on the current ring block.  One round trip: read the
            // window (previous block | what earlier, delayed blocks left in the current one) and the
            // input in window layout, add, keep the sum in registers and write the current block back.
            if constexpr (R == 1) {
                float2 q = *reinterpret_cast<const float2 *>(ring + off);
                if (lane >= 16) {
                    const float2 xi = *reinterpret_cast<const float2 *>(x + 2 * (lane & 15));
                    q.x = fmaf(gain, xi.x, q.x); q.y = fmaf(gain, xi.y, q.y);
                    *reinterpret_cast<float2 *>(ring + off) = q;
                }
                z[0] = q;
            } else {
#pragma unroll
                for (int j = 0; j < R / 2; j++) {
                    float4 q = *reinterpret_cast<const float4 *>(ring + off + 4 * j);
                    if (lane >= 16) {
                        const float4 xi = *reinterpret_cast<const float4 *>(x + 2 * R * (lane & 15) + 4 * j);
                        q.x = fmaf(gain, xi.x, q.x); q.y = fmaf(gain, xi.y, q.y);
                        q.z = fmaf(gain, xi.z, q.z); q.w = fmaf(gain, xi.w, q.w);
                        *reinterpret_cast<float4 *>(ring + off + 4 * j) = q;
                    }
                    z[2 * j] = make_float2(q.x, q.y);
                    z[2 * j + 1] = make_float2(q.z, q.w);
                }
            }
        } else {
            // predelay ring: the whole response of this block is delayed by pd samples (conv.cu:97):
            // scatter-add the scaled block at +pd, then read the window back
            __syncwarp();
#pragma unroll
            for (int j = 0; j < R; j++) {
                const int n = lane + 32 * j;
                const uint32_t idx = (base + pd + n) & mask;
                ring[idx] += gain * __ldg(&x[n]);
            }
            __syncwarp();
            if constexpr (R == 1) {
                z[0] = *reinterpret_cast<const float2 *>(ring + off);
            } else {
#pragma unroll
                for (int j = 0; j < R / 2; j++) {
                    const float4 q = *reinterpret_cast<const float4 *>(ring + off + 4 * j);
                    z[2 * j] = make_float2(q.x, q.y);
                    z[2 * j + 1] = make_float2(q.z, q.w);
                }
            }
        }
        f.forward(z);
        f.split_r2c(z, a.tw2M);
        float2 *dst = a.X + ((size_t)entry * a.Lring + slot) * B;
#pragma unroll
        for (int d = 0; d < R; d++) dst[f.c + 32 * d] = z[d];
    }
}

// ------------------------------------------------------------------------------------------
// IR partitions -> spectra (one warp per (output channel, partition))
// ------------------------------------------------------------------------------------------
struct IrArgs {
    const float *h[2];  // time-domain IR per output channel (device)
    float2 *H;          // this slot's spectra [n_out][P][B]
    const float2 *twM, *tw2M;
    uint32_t frames, P, n_out, frame_off;
    uint32_t stride;    // 1 = planar, 2 = (L, R) interleaved like WavFile::buffer (wav.h:10)
    float scale;        // 1/(2B): both FFT normalisations live in H
};

template <int R>
__global__ void __launch_bounds__(kFwdWarps * 32) k_ir_fft(const IrArgs a)
{
    constexpr int B = 32 * R;
    const int lane = threadIdx.x & 31;
    const uint32_t item = blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    if (item >= a.n_out * a.P) return;
    const uint32_t o = item / a.P, k = item % a.P;
    WarpFft<R> f;
    f.init(a.twM);
    const float *h = o == 0 ? a.h[0] : a.h[1];
    float2 v[R];
#pragma unroll
    for (int b = 0; b < R; b++) {
        float re = 0.f, im = 0.f;
        if (lane < 16) {  // [h_k | 0]: the IR block sits in the first half of the 2B window
            const size_t n0 = (size_t)a.frame_off + (size_t)k * B + 2 * (R * lane + b);
            if (n0 < a.frames) re = __ldg(&h[n0 * a.stride]) * a.scale;
            if (n0 + 1 < a.frames) im = __ldg(&h[(n0 + 1) * a.stride]) * a.scale;
        }
        v[b] = make_float2(re, im);
    }
    f.forward(v);
    f.split_r2c(v, a.tw2M);
    float2 *dst = a.H + ((size_t)o * a.P + k) * B;
#pragma unroll
    for (int d = 0; d < R; d++) dst[f.c + 32 * d] = v[d];
}

// ------------------------------------------------------------------------------------------
// FDL complex multiply-accumulate
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
#ifdef CA_MBAR_DEBUG
__device__ int g_mbar_abort = 0;
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
#ifdef CA_MBAR_DEBUG  // bounded spin: report the first barrier that never completes instead of hanging the GPU
    for (unsigned n = 0; !mbar_try_wait(bar, parity); n++) {
        if (*(volatile int *)&g_mbar_abort) return;
        if (n > (1u << 14)) {
            if (atomicExch(&g_mbar_abort, 1) == 0)
                printf("mbar timeout: block (%d,%d,%d) thread %d bar smem 0x%x parity %u\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, smem_u32(bar), parity);
            return;
        }
    }
#else
    while (!mbar_try_wait(bar, parity)) {}
#endif
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// One launch = one tier of one period for every active instance.  The work of an instance is the
// flattened list of (stream, partition) rows, stream = (input, voice); a row is the FDL spectrum
// X_s[n - k] plus the n_out IR spectra H_s,o[k] it multiplies.  Rows are cut into n_split
// contiguous ranges (one CTA each) so few instances still cover all SMs; bins are cut into tiles
// of BT complex.
struct MacArgs {
    const float2 *X;   // FDL [voice pool entry][Lring][S]
    const float2 *H;   // IR bank [slot][n_out][P][S]
    float2 *Ypart;     // [inst][n_split][n_out][S]
    const InParamDev *par;
    const ItemState *st;  // [2][n_items_alloc]
    const Ctl *ctl;
    uint32_t n_items_alloc;
    uint32_t n_in, nv;
    uint32_t Lring, P, S;
    uint32_t k_off;    // FDL delay of this engine's first partition (partition-range shards)
    uint32_t m;        // tier block / period: the tier fires when (t_end / period) % m == 0
    uint32_t t_bias;   // 1 for tier 0 (runs before k_inverse advances ctl->t), 0 for deferred tiers
    uint32_t n_split;
    uint32_t stream_hint;
    // instances of this launch: inst0 + blockIdx.z * inst_stride.  Long tiers are phase-staggered:
    // instance s closes its tier-j block when (t_end + s mod m) % m == 0, so every period only 1/m of
    // the instances run the tier and the load per period is flat.
    uint32_t inst0, inst_stride;
    uint32_t yp_local;  // 1: Ypart is indexed by the launch's instance index (long tiers: only 1/m of the instances fire)
    // pipelined batch schedule: the launch may run while k_inverse advances ctl->t, so the host passes
    // the period count itself (0: read ctl->t + t_bias)
    unsigned long long tend_host;
    uint32_t t_sel;  // see ctl_tend()
    // One IR split by partition range across GPUs (ca_group): the last CTA of the launch to finish
    // sums the launch's n_split partial spectra in fixed order and stores the ONE resulting spectrum
    // (n_out x S complex, 4 KB at B = 256) into `gather` -- this device's slot of the root GPU's gather
    // buffer, i.e. a peer store over NVLink when this GPU is not the root -- then publishes the period
    // count in *gflag (system-scope release).  The root's inverse kernel waits on the flags: compute and
    // exchange are one kernel, no collective launch.  nullptr: off.
    float2 *gather;
    unsigned long long *gflag;
    uint32_t *gcount;
    // persistent schedule: [0] next work item to hand out (beyond the first gridDim.x, which are static),
    // [1] CTAs that have finished; the last one to finish zeroes both for the next launch of this tier
    uint32_t *work_ctr;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// root side of the fused exchange: wait until every peer has published period count `epoch`.  Bounded
// (~2 s at 2 GHz): a peer that died must not hang this GPU; *gerr reports it.
__device__ __forceinline__ void group_wait(const unsigned long long *flags, uint32_t n_wait, unsigned long long epoch, int *gerr, int idx)
{
    if ((uint32_t)idx < n_wait) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + idx) < epoch) {
            if (clock64() - t0 > (4ll << 30)) { if (gerr) *gerr = 1; break; }
        }
    }
}

constexpr int kMacConsumers = 256;
constexpr int kMacThreads = kMacConsumers + 32;
constexpr int kMaxStreams = 2 * kMaxVoices;

template <int BT, int NOUT, int KC, int NSTAGE>
struct MacCfg {
    static constexpr int NARR = 1 + NOUT;                   // arrays per row: X, H_0 .. H_{NOUT-1}
    static constexpr int LR = BT / 2;                       // float4 lanes per array
    static constexpr int G = kMacConsumers / LR;            // rows processed concurrently
    static constexpr uint32_t ARR_BYTES = BT * 8;
    static constexpr uint32_t STAGE_BYTES = KC * NARR * ARR_BYTES;
    static constexpr uint32_t SMEM_BYTES = NSTAGE * STAGE_BYTES + 2 * NSTAGE * 8 + 16;
    static_assert(KC % G == 0, "stage rows must be a multiple of the row groups");
    static_assert(G * NOUT * LR * 16 + G * NOUT * 8 <= NSTAGE * STAGE_BYTES, "reduction scratch must fit");
};

template <int BT, int NOUT, int KC, int NSTAGE>
__global__ void __launch_bounds__(kMacThreads) k_mac(const MacArgs a)
{
    pdl_trigger();  // the next kernel of the stream may become resident while this one drains
    pdl_wait();     // before any global access and before any early exit: stream order holds transitively
    using Cfg = MacCfg<BT, NOUT, KC, NSTAGE>;
    constexpr int NARR = Cfg::NARR, LR = Cfg::LR, G = Cfg::G;
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *stage = reinterpret_cast<float4 *>(smem);  // [NSTAGE][KC][NARR][LR] float4
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NSTAGE * Cfg::STAGE_BYTES);
    uint64_t *empty = full + NSTAGE;
    __shared__ uint32_t s_rowstart[kMaxStreams + 1];  // prefix sum of rows per stream
    __shared__ uint32_t s_slot[kMaxStreams];
    __shared__ uint32_t s_entry[kMaxStreams];  // voice pool entry of every stream
    __shared__ float s_pan[2][NOUT];

    const uint32_t split = blockIdx.x, tile = blockIdx.y, inst = a.inst0 + blockIdx.z * a.inst_stride;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ns = a.n_in * a.nv;
    const unsigned long long tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, a.t_bias);  // periods completed at the end of this tier block
    const uint32_t phase = inst % a.m;
    const unsigned long long n_fire = (tend + phase) / a.m;

    // --- set-up, spread over three warps so the dependent global loads overlap ---
    if (tid == 32) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMacConsumers / 32); }
        mbar_fence_init();
    } else if (warp == 0) {
        // lane s: rows of stream s = (input s / nv, voice s % nv); exclusive prefix sum by shuffles
        uint32_t nk = 0, slot = 0, entry = 0;
        if ((uint32_t)lane < ns) {
            const uint32_t i = lane / a.nv, v = lane % a.nv;
            const ItemState &st = a.st[(tend & 1ull) * a.n_items_alloc + inst * a.n_in + i];
            if ((st.active >> v) & 1u) {
                entry = st.pool[v] - 1u;
                // partition k reads the block fired at index n_fire - k_off - k; it is valid when that
                // block was built after the voice's (re)start: fire * m - 1 >= start
                const long long first_fire = (long long)((st.start[v] + phase + a.m) / a.m);  // ceil((start + 1 + phase) / m)
                const long long cnt = (long long)n_fire - first_fire + 1 - (long long)a.k_off;
                nk = (uint32_t)max(0ll, min((long long)a.P, cnt));
                slot = st.slot[v];
            }
        }
        uint32_t incl = nk;
#pragma unroll
        for (int d = 1; d < kMaxStreams; d <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += up;
        }
        const uint32_t total_rows = __shfl_sync(kFull, incl, kMaxStreams - 1);  // all 32 lanes take part
        if (lane < kMaxStreams) { s_rowstart[lane] = incl - nk; s_slot[lane] = slot; s_entry[lane] = entry; }
        else if (lane == kMaxStreams) s_rowstart[kMaxStreams] = total_rows;
    } else if (warp == 2 && lane < 2 * NOUT) {
        const uint32_t i = lane / NOUT;
        const int o = lane % NOUT;
        s_pan[i][o] = pan_gain(a.par[inst * a.n_in + min(i, a.n_in - 1)].panWet, o, NOUT);
    }
    __syncthreads();

    const uint32_t total = s_rowstart[kMaxStreams];
    const uint32_t boundary = s_rowstart[a.nv];  // first row of input 1 (== total when n_in == 1)
    uint32_t rps = (total + a.n_split - 1) / a.n_split;
    rps = ((rps + KC - 1) / KC) * KC;
    const uint32_t r_begin = min(total, split * rps), r_end = min(total, r_begin + rps);
    const int nrows = (int)(r_end - r_begin);
    const int n_iter = (nrows + KC - 1) / KC;

    float4 acc[NOUT], y[NOUT];
    float2 e0[NOUT], y0[NOUT];  // bin 0 = (DC, Nyquist): two real products, not a complex one
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        acc[o] = y[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        e0[o] = y0[o] = make_float2(0.f, 0.f);
    }

    if (warp == kMacConsumers / 32) {
        // ===== producer warp: every lane issues its own bulk copies =====
        const uint32_t head = (a.Lring - 1u) - (uint32_t)(n_fire % a.Lring);
        const uint64_t pol = a.stream_hint ? l2_policy_evict_first() : l2_policy_evict_last();
        for (int it = 0; it < n_iter; it++) {
            const int st = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            if (it >= NSTAGE) mbar_wait(&empty[st], ph ^ 1u);
            const int rows = min(KC, nrows - it * KC);
            if (lane == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)rows * NARR * Cfg::ARR_BYTES);
            __syncwarp();
            for (int cidx = lane; cidx < rows * NARR; cidx += 32) {
                const int r = cidx / NARR, w = cidx % NARR;
                const uint32_t rho = r_begin + it * KC + r;
                uint32_t s = 0;
#pragma unroll
                for (int q = 1; q < kMaxStreams; q++) s += (rho >= s_rowstart[q]) ? 1u : 0u;
                const uint32_t k = rho - s_rowstart[s];
                const float2 *src;
                if (w == 0) {
                    const uint32_t slot = (head + a.k_off + k) % a.Lring;
                    src = a.X + ((size_t)s_entry[s] * a.Lring + slot) * a.S + tile * BT;
                } else {
                    src = a.H + (((size_t)s_slot[s] * NOUT + (w - 1)) * a.P + k) * a.S + tile * BT;
                }
                float4 *dst = stage + ((size_t)(st * KC + r) * NARR + w) * LR;
                tma_load_1d(dst, src, Cfg::ARR_BYTES, &full[st], pol);
            }
        }
    } else {
        // ===== consumers: thread (g, q) owns bins (2q, 2q+1) of every G-th row =====
        const int q = tid % LR, g = tid / LR;
        const bool bin0 = (q == 0) && (tile == 0);
        bool second = false;  // rows of input 1 reached: input 0's sum has been folded into y with its pan
        for (int it = 0; it < n_iter; it++) {
            const int st = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            const int rows = min(KC, nrows - it * KC);
            mbar_wait(&full[st], ph);
#pragma unroll
            for (int rr = 0; rr < KC / G; rr++) {
                const int r = g + rr * G;
                if (r < rows) {
                    if (!second && r_begin + it * KC + r >= boundary) {
                        second = true;
#pragma unroll
                        for (int o = 0; o < NOUT; o++) {
                            const float pan = s_pan[0][o];
                            y[o].x = pan * acc[o].x; y[o].y = pan * acc[o].y; y[o].z = pan * acc[o].z; y[o].w = pan * acc[o].w;
                            y0[o].x = pan * e0[o].x; y0[o].y = pan * e0[o].y;
                            acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                            e0[o] = make_float2(0.f, 0.f);
                        }
                    }
                    const float4 *row = stage + ((size_t)(st * KC + r) * NARR) * LR + q;
                    const float4 x = row[0];
#pragma unroll
                    for (int o = 0; o < NOUT; o++) {
                        const float4 h = row[(1 + o) * LR];
                        acc[o].x = fmaf(x.x, h.x, fmaf(-x.y, h.y, acc[o].x));
                        acc[o].y = fmaf(x.x, h.y, fmaf(x.y, h.x, acc[o].y));
                        acc[o].z = fmaf(x.z, h.z, fmaf(-x.w, h.w, acc[o].z));
                        acc[o].w = fmaf(x.z, h.w, fmaf(x.w, h.z, acc[o].w));
                        if (bin0) {
                            e0[o].x = fmaf(x.x, h.x, e0[o].x);
                            e0[o].y = fmaf(x.y, h.y, e0[o].y);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
#pragma unroll
        for (int o = 0; o < NOUT; o++) {  // fold the last input's sum with its pan (conv.cu:392-401)
            const float pan = s_pan[second ? 1 : 0][o];
            y[o].x = fmaf(pan, acc[o].x, y[o].x); y[o].y = fmaf(pan, acc[o].y, y[o].y);
            y[o].z = fmaf(pan, acc[o].z, y[o].z); y[o].w = fmaf(pan, acc[o].w, y[o].w);
            y0[o].x = fmaf(pan, e0[o].x, y0[o].x); y0[o].y = fmaf(pan, e0[o].y, y0[o].y);
        }
    }

    // ===== cross-group reduction (fixed order => deterministic), store the partial spectrum =====
    __syncthreads();  // every TMA write has landed and been consumed: stage memory is free
    float4 *red = reinterpret_cast<float4 *>(smem);                  // [G][NOUT][LR]
    float2 *red0 = reinterpret_cast<float2 *>(red + G * NOUT * LR);  // [G][NOUT]
    if (tid < kMacConsumers) {
        const int q = tid % LR, g = tid / LR;
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            red[(g * NOUT + o) * LR + q] = y[o];
            if (q == 0) red0[g * NOUT + o] = y0[o];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NOUT * LR; idx += kMacThreads) {
        const int o = idx / LR, q = idx % LR;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 s0 = make_float2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < G; g++) {
            const float4 v = red[(g * NOUT + o) * LR + q];
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            const float2 u = red0[g * NOUT + o];
            s0.x += u.x; s0.y += u.y;
        }
        if (q == 0 && tile == 0) { sum.x = s0.x; sum.y = s0.y; }
        float2 *dst = a.Ypart + (((size_t)(a.yp_local ? blockIdx.z : inst) * a.n_split + split) * NOUT + o) * a.S + tile * BT + 2 * q;
        *reinterpret_cast<float4 *>(dst) = sum;
    }
    if (a.gather) {  // ca_group: one instance; see MacArgs
        __shared__ uint32_t s_last;
        __threadfence();  // this CTA's partial is visible device-wide before it is counted
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(a.gcount, 1u) + 1u == gridDim.x * gridDim.y * gridDim.z) ? 1u : 0u;
        __syncthreads();
        if (s_last) {
            __threadfence();
            const float4 *src = reinterpret_cast<const float4 *>(a.Ypart);  // [n_split][NOUT][S]
            const uint32_t n4 = NOUT * a.S / 2;
            for (uint32_t idx = tid; idx < n4; idx += kMacThreads) {
                float4 sum = __ldcg(src + idx);
                for (uint32_t sp = 1; sp < a.n_split; sp++) {
                    const float4 v = __ldcg(src + (size_t)sp * n4 + idx);
                    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                }
                reinterpret_cast<float4 *>(a.gather)[idx] = sum;
            }
            __threadfence_system();  // the spectrum has landed (in the root's memory) before the flag does
            __syncthreads();
            if (tid == 0) { *a.gcount = 0u; st_release_sys(a.gflag, tend); }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Persistent schedule of the same MAC for batches (n_split == 1): a CTA walks the work list
// (instance, bin tile) with stride gridDim.x and its producer warp keeps the TMA ring full ACROSS
// work items, so the pipeline never drains between items (a one-item CTA streams only 96..264 KB:
// barrier set-up, first-byte latency and the final drain cost ~15 % of the HBM rate, ncu r01).
// Stage metadata travels with the stage (written before the full barrier is armed).
// ------------------------------------------------------------------------------------------
struct MacStageMeta {
    uint32_t rows;      // rows staged (0: the item has nothing to read yet)
    uint32_t rho0;      // index of the first staged row in the item's row list
    uint32_t boundary;  // first row of input 1
    uint32_t last;      // 1: last stage of the item
    float pan[4];       // [input][output] wet pan gains
    uint32_t item;      // work item (instance * tiles + tile) this stage belongs to; kMacNoItem ends the CTA's list
    uint32_t pad[3];
};
constexpr uint32_t kMacNoItem = 0xffffffffu;

template <int BT, int NOUT, int KC, int NSTAGE>
struct MacPCfg {
    using Base = MacCfg<BT, NOUT, KC, NSTAGE>;
    static constexpr int G = Base::G, LR = Base::LR;
    static constexpr uint32_t RING_BYTES = NSTAGE * Base::STAGE_BYTES;
    static constexpr uint32_t BAR_BYTES = 2 * NSTAGE * 8;
    static constexpr uint32_t META_BYTES = NSTAGE * sizeof(MacStageMeta);
    static constexpr uint32_t RED_BYTES = (G - 1) * NOUT * LR * 16 + G * NOUT * 8;
    static constexpr uint32_t SMEM_BYTES = RING_BYTES + BAR_BYTES + META_BYTES + RED_BYTES + 16;
};

template <int BT, int NOUT, int KC, int NSTAGE>
__global__ void __launch_bounds__(kMacThreads, 4) k_mac_p(const MacArgs a, const uint32_t n_work, const uint32_t n_tiles)
{
    pdl_trigger();  // the next kernel of the stream may become resident while this one drains
    pdl_wait();     // before any global access and before any early exit: stream order holds transitively
    using Cfg = MacCfg<BT, NOUT, KC, NSTAGE>;
    using PCfg = MacPCfg<BT, NOUT, KC, NSTAGE>;
    constexpr int NARR = Cfg::NARR, LR = Cfg::LR, G = Cfg::G;
    static_assert(NOUT <= 2, "pan table holds two outputs");
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *stage = reinterpret_cast<float4 *>(smem);  // [NSTAGE][KC][NARR][LR] float4
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + PCfg::RING_BYTES);
    uint64_t *empty = full + NSTAGE;
    MacStageMeta *meta = reinterpret_cast<MacStageMeta *>(smem + PCfg::RING_BYTES + PCfg::BAR_BYTES);
    float4 *red = reinterpret_cast<float4 *>(smem + PCfg::RING_BYTES + PCfg::BAR_BYTES + PCfg::META_BYTES);  // [G-1][NOUT][LR]
    float2 *red0 = reinterpret_cast<float2 *>(red + (G - 1) * NOUT * LR);                                     // [G][NOUT]
    __shared__ uint32_t s_rs[kMaxStreams], s_sl[kMaxStreams], s_en[kMaxStreams];  // producer-private: first row / IR slot / voice pool entry of every stream

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ns = a.n_in * a.nv;
    const unsigned long long tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, a.t_bias);
    // m (tier block / period) and n_tiles (S / BT) are powers of two: shifts and masks instead of the
    // software 64-bit divisions, which sat on the producer's critical path at every work item (ncu r01)
    const uint32_t m_log = 31 - __clz((int)a.m), m_mask = a.m - 1u;
    const uint32_t t_log = 31 - __clz((int)n_tiles), t_mask = n_tiles - 1u;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMacConsumers / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kMacConsumers / 32) {
        // ===== producer warp =====
        const uint64_t pol = a.stream_hint ? l2_policy_evict_first() : l2_policy_evict_last();
        // raw descriptor of a work item, one stream (lane < ns) / one pan gain (lane < 2 NOUT) per lane;
        // the loads of item j + 1 are issued before item j is streamed so their latency is hidden
        struct Raw { uint32_t active, slot, entry; unsigned long long start; float pan; };
        auto load_raw = [&](uint32_t j) {
            Raw r{0u, 0u, 0u, 0ull, 0.f};
            const uint32_t inst = a.inst0 + (j >> t_log) * a.inst_stride;
            if ((uint32_t)lane < ns) {
                const uint32_t i = lane / a.nv, v = lane % a.nv;
                const ItemState &st = a.st[(tend & 1ull) * a.n_items_alloc + inst * a.n_in + i];
                r.active = (st.active >> v) & 1u;
                r.start = st.start[v];
                r.slot = st.slot[v];
                r.entry = st.pool[v] - 1u;
            }
            if (lane < 2 * NOUT) {
                const uint32_t i = lane / NOUT;
                r.pan = pan_gain(a.par[inst * a.n_in + min(i, a.n_in - 1)].panWet, lane % NOUT, NOUT);
            }
            return r;
        };
        // Work items are handed out dynamically (an atomic counter; the first gridDim.x statically): CTAs on SMs
        // that happen to stream faster take more items, so the launch has no tail of a few late CTAs
        // (ncu r02: 7-10 % of the SM cycles of a MAC launch were idle with the static stride).
        // The atomic that hands out the item AFTER the next one is issued at the top of an item and its result is
        // first touched when the item has been streamed, so its round trip to L2 never sits on the critical path.
        // With few items per CTA (work_ctr == nullptr) the static stride j, j + gridDim.x, ... is kept: every CTA
        // gets the same count and the launch is ~3 % faster than with the counter (measured r02, 7 items per CTA).
        const bool dyn = a.work_ctr != nullptr;
        uint32_t it = 0;
        uint32_t j = blockIdx.x;
        uint32_t grabbed = j + gridDim.x;
        if (dyn && lane == 0) grabbed = gridDim.x + atomicAdd(a.work_ctr, 1u);
        Raw cur = load_raw(min(j, n_work - 1u)), nxt = cur;
        uint32_t jn = __shfl_sync(kFull, grabbed, 0);
        while (j < n_work) {
            if (jn < n_work) {
                nxt = load_raw(jn);
                grabbed = jn + gridDim.x;
                if (dyn && lane == 0) grabbed = gridDim.x + atomicAdd(a.work_ctr, 1u);
            }
            const uint32_t tile = j & t_mask, inst = a.inst0 + (j >> t_log) * a.inst_stride;
            const uint32_t phase = inst & m_mask;
            const unsigned long long n_fire = (tend + phase) >> m_log;
            uint32_t nk = 0;
            if (cur.active) {  // see k_mac: partitions whose FDL block was built after the voice's (re)start
                const long long first_fire = (long long)((cur.start + phase + a.m) >> m_log);
                const long long cnt = (long long)n_fire - first_fire + 1 - (long long)a.k_off;
                nk = (uint32_t)max(0ll, min((long long)a.P, cnt));
            }
            uint32_t incl = nk;
#pragma unroll
            for (int d = 1; d < kMaxStreams; d <<= 1) {
                const uint32_t up = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += up;
            }
            const uint32_t excl = incl - nk;
            __syncwarp();  // every lane has issued the previous item's copies: the row table may change
            if (lane < kMaxStreams) { s_rs[lane] = excl; s_sl[lane] = cur.slot; s_en[lane] = cur.entry; }
            __syncwarp();
            const uint32_t total = __shfl_sync(kFull, incl, kMaxStreams - 1);
            const uint32_t boundary = a.nv < kMaxStreams ? __shfl_sync(kFull, excl, a.nv) : total;
            float pan[4];
#pragma unroll
            for (int q = 0; q < 4; q++) pan[q] = __shfl_sync(kFull, cur.pan, q < 2 * NOUT ? q : 0);
            // ring position of partition 0: (Lring - 1 - n_fire % Lring + k_off) % Lring; 32-bit fast path
            const uint32_t nf_mod = (n_fire >> 32) ? (uint32_t)(n_fire % a.Lring) : (uint32_t)n_fire % a.Lring;
            const uint32_t head0 = ((a.Lring - 1u) - nf_mod + a.k_off % a.Lring) % a.Lring;
            const uint32_t n_iter = max(1u, (total + KC - 1) / KC);
            for (uint32_t i = 0; i < n_iter; i++, it++) {
                const uint32_t st = it % NSTAGE;
                const uint32_t ph = (it / NSTAGE) & 1u;
                if (it >= NSTAGE) mbar_wait(&empty[st], ph ^ 1u);
                const uint32_t rows = total > i * KC ? min((uint32_t)KC, total - i * KC) : 0u;
                if (lane == 0) {
                    MacStageMeta &m = meta[st];
                    m.rows = rows; m.rho0 = i * KC; m.boundary = boundary; m.last = (i + 1 == n_iter) ? 1u : 0u;
                    m.pan[0] = pan[0]; m.pan[1] = pan[1]; m.pan[2] = pan[2]; m.pan[3] = pan[3];
                    m.item = j;
                    if (rows) mbar_arrive_expect_tx(&full[st], rows * NARR * Cfg::ARR_BYTES);
                    else mbar_arrive(&full[st]);
                }
                __syncwarp();
                for (uint32_t cidx = lane; cidx < rows * NARR; cidx += 32) {
                    const uint32_t r = cidx / NARR, w = cidx % NARR;
                    const uint32_t rho = i * KC + r;
                    uint32_t s = 0;
#pragma unroll
                    for (int q = 1; q < kMaxStreams; q++) s += (rho >= s_rs[q]) ? 1u : 0u;
                    const uint32_t k = rho - s_rs[s], slot = s_sl[s];
                    const float2 *src;
                    if (w == 0) {
                        uint32_t pos = head0 + k;  // k < P <= Lring
                        if (pos >= a.Lring) pos -= a.Lring;
                        src = a.X + ((size_t)s_en[s] * a.Lring + pos) * a.S + tile * BT;
                    } else {
                        src = a.H + (((size_t)slot * NOUT + (w - 1)) * a.P + k) * a.S + tile * BT;
                    }
                    float4 *dst = stage + ((size_t)(st * KC + r) * NARR + w) * LR;
                    tma_load_1d(dst, src, Cfg::ARR_BYTES, &full[st], pol);
                }
            }
            cur = nxt;
            j = jn;
            jn = (jn < n_work) ? __shfl_sync(kFull, grabbed, 0) : kMacNoItem;
        }
        {   // end of this CTA's list: one empty stage tells the consumers
            const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1u;
            if (it >= NSTAGE) mbar_wait(&empty[st], ph ^ 1u);
            if (lane == 0) {
                MacStageMeta &m = meta[st];
                m.rows = 0u; m.rho0 = 0u; m.boundary = 0u; m.last = 1u; m.item = kMacNoItem;
                mbar_arrive(&full[st]);
            }
        }
    } else {
        // ===== consumers: thread (g, q) owns bins (2q, 2q+1).  Row groups g take every G-th row and their
        // sums are added through shared memory at the end of the item -- except when there are as many
        // groups as outputs (BT = 256, true stereo): then group g takes EVERY row for output g, and the
        // item ends without a barrier or a reduction (OSPLIT). =====
        constexpr bool OSPLIT = (G == NOUT) && (G > 1);
        constexpr int NO = OSPLIT ? 1 : NOUT;     // outputs per thread
        constexpr int RSTEP = OSPLIT ? 1 : G;     // row stride of a thread
        const int q = tid % LR, g = tid / LR;
        const int r_first = OSPLIT ? 0 : g, o_base = OSPLIT ? g : 0;
        uint32_t it = 0;
        for (;;) {
            uint32_t tile = 0, z = 0;
            bool bin0 = false, first = true, done = false;
            float4 acc[NO], y[NO];
            float2 e0[NO], y0[NO];
#pragma unroll
            for (int o = 0; o < NO; o++) {
                acc[o] = y[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                e0[o] = y0[o] = make_float2(0.f, 0.f);
            }
            bool second = false;
            uint32_t last;
            float pan0[NO], pan1[NO];
            do {
                const uint32_t st = it % NSTAGE;
                const uint32_t ph = (it / NSTAGE) & 1u;
                mbar_wait(&full[st], ph);
                const MacStageMeta &m = meta[st];
                const uint32_t rows = m.rows, rho0 = m.rho0, boundary = m.boundary;
                last = m.last;
                if (first) {  // first stage of an item: which one it is
                    first = false;
                    const uint32_t j = m.item;
                    done = j == kMacNoItem;
                    tile = j & t_mask; z = j >> t_log;
                    bin0 = (q == 0) && (tile == 0);
                }
#pragma unroll
                for (int o = 0; o < NO; o++) { pan0[o] = m.pan[o_base + o]; pan1[o] = m.pan[NOUT + o_base + o]; }
#pragma unroll
                for (int rr = 0; rr < KC / RSTEP; rr++) {
                    const uint32_t r = r_first + rr * RSTEP;
                    if (r < rows) {
                        if (!second && rho0 + r >= boundary) {
                            second = true;
#pragma unroll
                            for (int o = 0; o < NO; o++) {
                                const float pan = pan0[o];
                                y[o].x = pan * acc[o].x; y[o].y = pan * acc[o].y; y[o].z = pan * acc[o].z; y[o].w = pan * acc[o].w;
                                y0[o].x = pan * e0[o].x; y0[o].y = pan * e0[o].y;
                                acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                                e0[o] = make_float2(0.f, 0.f);
                            }
                        }
                        const float4 *row = stage + ((size_t)(st * KC + r) * NARR) * LR + q;
                        const float4 x = row[0];
#pragma unroll
                        for (int o = 0; o < NO; o++) {
                            const float4 h = row[(1 + o_base + o) * LR];
                            acc[o].x = fmaf(x.x, h.x, fmaf(-x.y, h.y, acc[o].x));
                            acc[o].y = fmaf(x.x, h.y, fmaf(x.y, h.x, acc[o].y));
                            acc[o].z = fmaf(x.z, h.z, fmaf(-x.w, h.w, acc[o].z));
                            acc[o].w = fmaf(x.z, h.w, fmaf(x.w, h.z, acc[o].w));
                            if (bin0) {
                                e0[o].x = fmaf(x.x, h.x, e0[o].x);
                                e0[o].y = fmaf(x.y, h.y, e0[o].y);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
                it++;
            } while (!last);
            if (done) break;
#pragma unroll
            for (int o = 0; o < NO; o++) {  // fold the last input's sum with its pan (conv.cu:392-401)
                const float pan = second ? pan1[o] : pan0[o];
                y[o].x = fmaf(pan, acc[o].x, y[o].x); y[o].y = fmaf(pan, acc[o].y, y[o].y);
                y[o].z = fmaf(pan, acc[o].z, y[o].z); y[o].w = fmaf(pan, acc[o].w, y[o].w);
                y0[o].x = fmaf(pan, e0[o].x, y0[o].x); y0[o].y = fmaf(pan, e0[o].y, y0[o].y);
            }
            float2 *ybase = a.Ypart + ((size_t)(a.yp_local ? z : a.inst0 + z * a.inst_stride) * NOUT) * a.S + tile * BT + 2 * q;
            if constexpr (G == 1 || OSPLIT) {
#pragma unroll
                for (int o = 0; o < NO; o++) {
                    if (bin0) { y[o].x = y0[o].x; y[o].y = y0[o].y; }
                    *reinterpret_cast<float4 *>(ybase + (size_t)(o_base + o) * a.S) = y[o];
                }
            } else {
                // cross-group sum in fixed order (deterministic); consumer-only named barrier
                if (g > 0) {
#pragma unroll
                    for (int o = 0; o < NOUT; o++) red[((g - 1) * NOUT + o) * LR + q] = y[o];
                }
                if (q == 0) {
#pragma unroll
                    for (int o = 0; o < NOUT; o++) red0[g * NOUT + o] = y0[o];
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kMacConsumers) : "memory");
                if (g == 0) {
#pragma unroll
                    for (int o = 0; o < NOUT; o++) {
                        float4 sum = y[o];
#pragma unroll
                        for (int gg = 1; gg < G; gg++) {
                            const float4 v = red[((gg - 1) * NOUT + o) * LR + q];
                            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                        }
                        if (bin0) {
                            float2 s0 = make_float2(0.f, 0.f);
#pragma unroll
                            for (int gg = 0; gg < G; gg++) { s0.x += red0[gg * NOUT + o].x; s0.y += red0[gg * NOUT + o].y; }
                            sum.x = s0.x; sum.y = s0.y;
                        }
                        *reinterpret_cast<float4 *>(ybase + (size_t)o * a.S) = sum;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kMacConsumers) : "memory");  // scratch free for the next item
            }
        }
    }
    __syncthreads();  // the producer warp stays resident until every copy it issued has been consumed
    if (a.work_ctr && tid == 0 && atomicAdd(a.work_ctr + 1, 1u) + 1u == gridDim.x) { a.work_ctr[0] = 0u; a.work_ctr[1] = 0u; }  // nobody hands out items any more
}

// ------------------------------------------------------------------------------------------
// inverse (tier 0): one CTA per (instance, output)
// ------------------------------------------------------------------------------------------
struct InvArgs {
    const float2 *Ypart;  // [inst][n_split][n_out][B]
    const float *in;      // [inst][n_in][B]   (dry path)
    float *out;           // [inst][n_out][B]
    float *accring;       // [inst*n_out + o][acc_len]: output of the deferred tiers, or nullptr
    const InParamDev *par;
    Ctl *ctl;
    const float2 *twM, *tw2M;
    uint32_t n_split, n_in, n_out, acc_len;
    uint32_t n_items;  // (instance, output) items of this launch
    uint32_t item0;    // first item of this launch
    uint32_t advance;  // 1: this launch completes the period (advances ctl->t)
    uint32_t raw_wet;  // 1: store the unclamped wet block only (partition-range shards: clamp + dry after the reduce)
    unsigned long long t_host_p1;  // see FwdArgs
    const float2 *rowtw;           // see FwdArgs
    // ca_group root: Ypart is the gather buffer [n_wait + 1][n_out][B]; slots 1.. are written by the peers'
    // MAC kernels over NVLink: wait for their flags to reach this period's count first (see MacArgs)
    const unsigned long long *wait_flags;
    uint32_t n_wait;
    int *gerr;
};

#ifndef CA_INV_THREADS
#define CA_INV_THREADS 128
#endif
constexpr int kInvThreads = CA_INV_THREADS;

// PACKED: one warp per (instance, output), the warp sums the (few) partial spectra itself -- the
// throughput schedule.  !PACKED: one CTA per item, all 128 threads sum the n_split partials through
// shared memory before warp 0 transforms -- the latency schedule (n_split up to 32).
template <int R, bool PACKED>
__global__ void __launch_bounds__(kInvThreads) k_inverse(const InvArgs a)
{
    pdl_trigger();  // the next kernel of the stream may become resident while this one drains
    pdl_wait();     // before any global access and before any early exit: stream order holds transitively
    constexpr int B = 32 * R;
    __shared__ __align__(16) float2 Ys[PACKED ? 2 : B];
    const int tid = threadIdx.x;
    const uint32_t local = PACKED ? blockIdx.x * (kInvThreads / 32) + (tid >> 5) : blockIdx.x;
    if (PACKED && local >= a.n_items) return;
    const uint32_t item = a.item0 + local;
    const uint32_t inst = item / a.n_out, o = item % a.n_out;
    const unsigned long long t = a.t_host_p1 ? a.t_host_p1 - 1ull : a.ctl->t_next - 1ull;
    if (a.n_wait) {
        group_wait(a.wait_flags, a.n_wait, t + 1ull, a.gerr, PACKED ? (tid & 31) : tid);
        if constexpr (PACKED) __syncwarp(); else __syncthreads();
    }

    if constexpr (!PACKED) {
        // --- sum the partial spectra of the MAC splits (fixed order) ---
        for (int f4 = tid; f4 < B / 2; f4 += kInvThreads) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 *src = reinterpret_cast<const float4 *>(a.Ypart + (((size_t)inst * a.n_split) * a.n_out + o) * B) + f4;
            const size_t stride = (size_t)a.n_out * B / 2;
#pragma unroll 4
            for (uint32_t sp = 0; sp < a.n_split; sp++) {
                const float4 v = src[sp * stride];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            reinterpret_cast<float4 *>(Ys)[f4] = s;
        }
        __syncthreads();
    }

    if (PACKED || tid < 32) {
        const int lane = tid & 31;
        WarpFft<R> f;
        f.init(a.twM);
        float2 v[R];
        if constexpr (PACKED) {
            const float2 *src = a.Ypart + (((size_t)inst * a.n_split) * a.n_out + o) * B;
#pragma unroll
            for (int d = 0; d < R; d++) v[d] = src[f.c + 32 * d];
            for (uint32_t sp = 1; sp < a.n_split; sp++) {
                src += (size_t)a.n_out * B;
#pragma unroll
                for (int d = 0; d < R; d++) { const float2 q = src[f.c + 32 * d]; v[d].x += q.x; v[d].y += q.y; }
            }
        } else {
#pragma unroll
            for (int d = 0; d < R; d++) v[d] = Ys[f.c + 32 * d];
        }
        // overlap discard: only time samples [B, 2B) are kept = lanes 16..31; lane a holds output samples
        // [2R (a-16), 2R (a-16) + 2R).  Everything the epilogue needs is fetched BEFORE the transform so
        // the loads overlap the FFT's shuffle chain.
        constexpr int NV4 = R >= 2 ? R / 2 : 1;
        float4 accv[NV4], xa[NV4], xb[NV4];
        float dg[2] = {0.f, 0.f};
        const int n0 = 2 * R * ((lane & 15));
        float *dst = a.out + ((size_t)inst * a.n_out + o) * B + n0;
        float *accp = nullptr;
        const bool raw = a.raw_wet != 0;
        if (lane >= 16) {
            if (!raw) {  // dry gain per input: dry * panDry * level, conv.cu:418-427
                const InParamDev p0 = a.par[inst * a.n_in];
                const InParamDev p1 = a.par[inst * a.n_in + (a.n_in - 1)];
                dg[0] = p0.dry * pan_gain(p0.panDry, (int)o, (int)a.n_out) * p0.level;
                dg[1] = a.n_in > 1 ? p1.dry * pan_gain(p1.panDry, (int)o, (int)a.n_out) * p1.level : 0.f;
            }
            const float *x0 = a.in + ((size_t)inst * a.n_in) * B + n0;
            const float *x1 = x0 + (a.n_in > 1 ? B : 0);
            if (a.accring) accp = a.accring + (size_t)item * a.acc_len + (uint32_t)((t * (unsigned long long)B) & (a.acc_len - 1)) + n0;
            if constexpr (R == 1) {
                const float2 q0 = *reinterpret_cast<const float2 *>(x0), q1 = *reinterpret_cast<const float2 *>(x1);
                xa[0] = make_float4(q0.x, q0.y, 0.f, 0.f);
                xb[0] = make_float4(q1.x, q1.y, 0.f, 0.f);
                accv[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (accp) { const float2 q = *reinterpret_cast<const float2 *>(accp); accv[0].x = q.x; accv[0].y = q.y; }
            } else {
#pragma unroll
                for (int j = 0; j < NV4; j++) {
                    xa[j] = *reinterpret_cast<const float4 *>(x0 + 4 * j);
                    xb[j] = *reinterpret_cast<const float4 *>(x1 + 4 * j);
                    accv[j] = accp ? *reinterpret_cast<const float4 *>(accp + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        f.split_c2r(v, a.tw2M);
        f.inverse(v);
        if (lane >= 16) {
            auto clampf = [raw](float w) { return raw ? w : fminf(fmaxf(w, -1.0f), 1.0f); };  // conv.cu:98
            if constexpr (R == 1) {
                float2 y = make_float2(clampf(v[0].x + accv[0].x), clampf(v[0].y + accv[0].y));
                y.x = fmaf(dg[0], xa[0].x, fmaf(dg[1], xb[0].x, y.x));
                y.y = fmaf(dg[0], xa[0].y, fmaf(dg[1], xb[0].y, y.y));
                *reinterpret_cast<float2 *>(dst) = y;
                if (accp) *reinterpret_cast<float2 *>(accp) = make_float2(0.f, 0.f);
            } else {
#pragma unroll
                for (int j = 0; j < NV4; j++) {
                    // wet = this period's tier-0 block + what the long (deferred) tiers left for it
                    float4 y = make_float4(clampf(v[2 * j].x + accv[j].x), clampf(v[2 * j].y + accv[j].y),
                                           clampf(v[2 * j + 1].x + accv[j].z), clampf(v[2 * j + 1].y + accv[j].w));
                    y.x = fmaf(dg[0], xa[j].x, fmaf(dg[1], xb[j].x, y.x));
                    y.y = fmaf(dg[0], xa[j].y, fmaf(dg[1], xb[j].y, y.y));
                    y.z = fmaf(dg[0], xa[j].z, fmaf(dg[1], xb[j].z, y.z));
                    y.w = fmaf(dg[0], xa[j].w, fmaf(dg[1], xb[j].w, y.w));
                    *reinterpret_cast<float4 *>(dst + 4 * j) = y;
                    if (accp) *reinterpret_cast<float4 *>(accp + 4 * j) = make_float4(0.f, 0.f, 0.f, 0.f);  // consumed
                }
            }
        }
    }
    if (a.advance && local == 0 && tid == 0) { a.ctl->t = t + 1ull; a.ctl->t_def[(t + 1ull) & 1ull] = t + 1ull; }  // forward + MAC of this period are done (nobody here reads ctl->t)
}

// ------------------------------------------------------------------------------------------
// fused tier 0 (throughput schedule of the tiered engine): one CTA per instance does
//   forward R2C (FFT warps)  ||  TMA streaming of the IR + FDL rows (producer warp)
//   -> MAC (8 consumer warps) -> inverse C2R + ring + clamp + dry (FFT warps)
// With P_0 = 8 partitions the whole working set of an instance (16 rows x 6 KB) is in flight at once,
// the newest spectrum goes from the FFT warps to the consumers through shared memory (rows are
// ordered newest-last), and the partial-spectrum round trip through global memory disappears.
// ------------------------------------------------------------------------------------------
struct FusedArgs {
    const float *in;      // [inst][n_in][B]
    float *out;           // [inst][n_out][B]
    float *ring;          // time rings
    float2 *X;            // tier-0 FDL
    const float2 *H;      // tier-0 IR bank [slot][n_out][P][B]
    float *accring;       // output ring of the long tiers (may be nullptr)
    const InParamDev *par;
    ItemState *st;
    Ctl *ctl;
    const float2 *twM, *tw2M;
    uint32_t n_items_alloc, n_in, nv, Lring, P, ring_len, ring_out, acc_len;
    uint32_t inst0, stream_hint, raw_wet;
    VoicePool vp;
};

constexpr int kFusedThreads = kMacConsumers + 32 + 64;  // 8 consumer warps, 1 producer warp, 2 FFT warps
constexpr int kFusedStages = 7;

template <int R, int NOUT>
struct FusedCfg {
    static constexpr int B = 32 * R, LR = B / 2, G = kMacConsumers / LR, KC = G, NARR = 1 + NOUT;
    static constexpr uint32_t ARR_BYTES = B * 8;
    static constexpr uint32_t STAGE_BYTES = KC * NARR * ARR_BYTES;                 // 12 KB at NOUT = 2
    static constexpr uint32_t XNEW_BYTES = 4 * ARR_BYTES;                          // newest spectrum per (input, voice), <= 4 streams
    static constexpr uint32_t YS_BYTES = NOUT * ARR_BYTES;
    static constexpr uint32_t SMEM_BYTES = kFusedStages * STAGE_BYTES + XNEW_BYTES + YS_BYTES + (2 * kFusedStages + 1) * 8 + 16;
};

__global__ void k_tick(Ctl *ctl) { ctl->t = ctl->t_next; ctl->t_def[ctl->t_next & 1ull] = ctl->t_next; }  // the fused kernel reads ctl->t in every CTA: advance afterwards

template <int R, int NOUT>
__global__ void __launch_bounds__(kFusedThreads, 2) k_fused0(const FusedArgs a)
{
    using Cfg = FusedCfg<R, NOUT>;
    constexpr int B = Cfg::B, LR = Cfg::LR, G = Cfg::G, KC = Cfg::KC, NARR = Cfg::NARR, NSTAGE = kFusedStages;
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *stage = reinterpret_cast<float4 *>(smem);
    float2 *xnew = reinterpret_cast<float2 *>(smem + NSTAGE * Cfg::STAGE_BYTES);   // [stream][B]
    float2 *Ys = xnew + 4 * B;                                                     // [NOUT][B]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NSTAGE * Cfg::STAGE_BYTES + Cfg::XNEW_BYTES + Cfg::YS_BYTES);
    uint64_t *empty = full + NSTAGE;
    uint64_t *xready = empty + NSTAGE;
    __shared__ uint32_t s_nk[4], s_slot[4], s_entry[4], s_rowstart[5];
    __shared__ float s_pan[2][NOUT];

    const uint32_t inst = a.inst0 + blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fw = warp - (kMacConsumers / 32 + 1);  // FFT warp index (0, 1) or negative
    const uint32_t ns = a.n_in * a.nv;
    const unsigned long long t = a.ctl->t;
    const uint32_t slot_new = (a.Lring - 1u) - (uint32_t)((t + 1ull) % a.Lring);

    ItemState s{};
    InParamDev p{};
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NSTAGE; q++) { mbar_init(&full[q], 1); mbar_init(&empty[q], kMacConsumers / 32); }
        mbar_init(xready, 2);
        mbar_fence_init();
        if (blockIdx.x == 0) a.ctl->t_next = t + 1ull;
    }
    if (fw >= 0) {
        // --- voice state of item (inst, fw); rows each voice contributes ---
        if ((uint32_t)fw < a.n_in) {
            const uint32_t item = inst * a.n_in + fw;
            p = a.par[item];
            s = item_step_warp(a.st, a.n_items_alloc, item, p, t, (int)a.nv, a.ring_out, a.vp, lane);
            if ((uint32_t)lane < a.nv) {
                uint32_t nk = 0, sl = 0;
                if ((s.active >> lane) & 1u) {
                    unsigned long long st0 = 0;
                    s_entry[fw * a.nv + lane] = voice_entry(s, (uint32_t)lane);
#pragma unroll
                    for (int q = 0; q < kMaxVoices; q++) { if (q == lane) { st0 = s.start[q]; sl = s.slot[q]; } }
                    const long long cnt = (long long)(t + 1ull) - (long long)(st0 + 1ull) + 1;
                    nk = (uint32_t)max(0ll, min((long long)a.P, cnt));
                }
                s_nk[fw * a.nv + lane] = nk;
                s_slot[fw * a.nv + lane] = sl;
            }
            if (lane < NOUT) s_pan[fw][lane] = pan_gain(p.panWet, lane, NOUT);
        }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t q = 0; q < 4; q++) { s_rowstart[q] = acc; acc += q < ns ? s_nk[q] : 0u; }
        s_rowstart[4] = acc;
    }
    __syncthreads();
    const uint32_t total = s_rowstart[4];
    const uint32_t boundary = a.n_in > 1 ? s_rowstart[a.nv] : total;
    const int n_iter = (int)((total + KC - 1) / KC);

    if (fw >= 0) {
        // ===== FFT warps: forward transform of every audible voice of input fw =====
        if ((uint32_t)fw < a.n_in) {
            const uint32_t item = inst * a.n_in + fw;
            const uint32_t pd = a.par[inst * a.n_in].predelay;  // input 0's, conv.cu:412,415
            WarpFft<R> f;
            f.init(a.twM);
            const uint32_t mask = a.ring_len - 1;
            const float *x = a.in + (size_t)item * B;
            const uint32_t base = (uint32_t)((t * (unsigned long long)B) & mask);
            const uint32_t prev = (base - B) & mask;
            const uint32_t off = ((lane < 16) ? prev : base) + 2 * R * (lane & 15);
#pragma unroll 1
            for (uint32_t v = 0; v < a.nv; v++) {
                if (!((s.active >> v) & 1u)) continue;
                float cv = 0.f;
#pragma unroll
                for (int q = 0; q < kMaxVoices; q++) cv = (q == (int)v) ? s.c[q] : cv;
                const float gain = cv * p.level;
                const uint32_t entry = voice_entry(s, v);
                float *ring = a.ring + (size_t)entry * a.ring_len;
                if ((s.fresh >> v) & 1u) {
                    for (uint32_t n = 4 * lane; n < a.ring_len; n += 128) *reinterpret_cast<float4 *>(ring + n) = make_float4(0.f, 0.f, 0.f, 0.f);
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < R; j++) ring[(base + kMaxPredelay + lane + 32 * j) & mask] = 0.f;
                float2 z[R];
                if (pd == 0) {
                    if constexpr (R == 1) {
                        float2 q = *reinterpret_cast<const float2 *>(ring + off);
                        if (lane >= 16) {
                            const float2 xi = *reinterpret_cast<const float2 *>(x + 2 * (lane & 15));
                            q.x = fmaf(gain, xi.x, q.x); q.y = fmaf(gain, xi.y, q.y);
                            *reinterpret_cast<float2 *>(ring + off) = q;
                        }
                        z[0] = q;
                    } else {
#pragma unroll
                        for (int j = 0; j < R / 2; j++) {
                            float4 q = *reinterpret_cast<const float4 *>(ring + off + 4 * j);
                            if (lane >= 16) {
                                const float4 xi = *reinterpret_cast<const float4 *>(x + 2 * R * (lane & 15) + 4 * j);
                                q.x = fmaf(gain, xi.x, q.x); q.y = fmaf(gain, xi.y, q.y);
                                q.z = fmaf(gain, xi.z, q.z); q.w = fmaf(gain, xi.w, q.w);
                                *reinterpret_cast<float4 *>(ring + off + 4 * j) = q;
                            }
                            z[2 * j] = make_float2(q.x, q.y);
                            z[2 * j + 1] = make_float2(q.z, q.w);
                        }
                    }
                } else {
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < R; j++) {
                        const int n = lane + 32 * j;
                        ring[(base + pd + n) & mask] += gain * __ldg(&x[n]);
                    }
                    __syncwarp();
                    if constexpr (R == 1) {
                        z[0] = *reinterpret_cast<const float2 *>(ring + off);
                    } else {
#pragma unroll
                        for (int j = 0; j < R / 2; j++) {
                            const float4 q = *reinterpret_cast<const float4 *>(ring + off + 4 * j);
                            z[2 * j] = make_float2(q.x, q.y);
                            z[2 * j + 1] = make_float2(q.z, q.w);
                        }
                    }
                }
                f.forward(z);
                f.split_r2c(z, a.tw2M);
                float2 *dst = a.X + ((size_t)entry * a.Lring + slot_new) * B;
                float2 *xs = xnew + (size_t)(fw * a.nv + v) * B;
#pragma unroll
                for (int d = 0; d < R; d++) { dst[f.c + 32 * d] = z[d]; xs[f.c + 32 * d] = z[d]; }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(xready);  // both FFT warps arrive, with or without work
    } else if (warp == kMacConsumers / 32) {
        // ===== producer warp: rows newest-last; the newest row's X comes from the FFT warps =====
        const uint64_t pol = a.stream_hint ? l2_policy_evict_first() : l2_policy_evict_last();
        for (int it = 0; it < n_iter; it++) {
            const int st = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            if (it >= NSTAGE) mbar_wait(&empty[st], ph ^ 1u);
            const int rows = min(KC, (int)total - it * KC);
            // every lane owns at most one copy per round; count the bytes first
            uint32_t my_bytes = 0;
            for (int cidx = lane; cidx < rows * NARR; cidx += 32) {
                const int r = cidx / NARR, w = cidx % NARR;
                const uint32_t rho = it * KC + r;
                uint32_t sidx = 0;
#pragma unroll
                for (int q = 1; q < 4; q++) sidx += (rho >= s_rowstart[q]) ? 1u : 0u;
                const uint32_t k = s_nk[sidx] - 1u - (rho - s_rowstart[sidx]);
                if (!(w == 0 && k == 0)) my_bytes += Cfg::ARR_BYTES;
            }
            const uint32_t bytes = __reduce_add_sync(kFull, my_bytes);
            if (lane == 0) mbar_arrive_expect_tx(&full[st], bytes);
            __syncwarp();
            for (int cidx = lane; cidx < rows * NARR; cidx += 32) {
                const int r = cidx / NARR, w = cidx % NARR;
                const uint32_t rho = it * KC + r;
                uint32_t sidx = 0;
#pragma unroll
                for (int q = 1; q < 4; q++) sidx += (rho >= s_rowstart[q]) ? 1u : 0u;
                const uint32_t k = s_nk[sidx] - 1u - (rho - s_rowstart[sidx]);
                const float2 *src;
                if (w == 0) {
                    if (k == 0) continue;
                    src = a.X + ((size_t)s_entry[sidx] * a.Lring + (slot_new + k) % a.Lring) * B;
                } else {
                    src = a.H + (((size_t)s_slot[sidx] * NOUT + (w - 1)) * a.P + k) * B;
                }
                tma_load_1d(stage + ((size_t)(st * KC + r) * NARR + w) * LR, src, Cfg::ARR_BYTES, &full[st], pol);
            }
        }
    }

    float4 acc[NOUT], y[NOUT];
    float2 e0[NOUT], y0[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; o++) { acc[o] = y[o] = make_float4(0.f, 0.f, 0.f, 0.f); e0[o] = y0[o] = make_float2(0.f, 0.f); }
    if (tid < kMacConsumers) {
        const int q = tid % LR, g = tid / LR;
        const bool bin0 = q == 0;
        bool second = false, xwaited = false;
        for (int it = 0; it < n_iter; it++) {
            const int st = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            const int rows = min(KC, (int)total - it * KC);
            mbar_wait(&full[st], ph);
            if (g < rows) {
                const uint32_t rho = it * KC + g;
                uint32_t sidx = 0;
#pragma unroll
                for (int qq = 1; qq < 4; qq++) sidx += (rho >= s_rowstart[qq]) ? 1u : 0u;
                const uint32_t k = s_nk[sidx] - 1u - (rho - s_rowstart[sidx]);
                if (!second && rho >= boundary) {
                    second = true;
#pragma unroll
                    for (int o = 0; o < NOUT; o++) {
                        const float pan = s_pan[0][o];
                        y[o].x = pan * acc[o].x; y[o].y = pan * acc[o].y; y[o].z = pan * acc[o].z; y[o].w = pan * acc[o].w;
                        y0[o].x = pan * e0[o].x; y0[o].y = pan * e0[o].y;
                        acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                        e0[o] = make_float2(0.f, 0.f);
                    }
                }
                const float4 *row = stage + ((size_t)(st * KC + g) * NARR) * LR + q;
                float4 x;
                if (k == 0) {
                    if (!xwaited) { mbar_wait(xready, 0); xwaited = true; }
                    x = reinterpret_cast<const float4 *>(xnew + (size_t)sidx * B)[q];
                } else x = row[0];
#pragma unroll
                for (int o = 0; o < NOUT; o++) {
                    const float4 h = row[(1 + o) * LR];
                    acc[o].x = fmaf(x.x, h.x, fmaf(-x.y, h.y, acc[o].x));
                    acc[o].y = fmaf(x.x, h.y, fmaf(x.y, h.x, acc[o].y));
                    acc[o].z = fmaf(x.z, h.z, fmaf(-x.w, h.w, acc[o].z));
                    acc[o].w = fmaf(x.z, h.w, fmaf(x.w, h.z, acc[o].w));
                    if (bin0) { e0[o].x = fmaf(x.x, h.x, e0[o].x); e0[o].y = fmaf(x.y, h.y, e0[o].y); }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            const float pan = s_pan[second ? 1 : 0][o];
            y[o].x = fmaf(pan, acc[o].x, y[o].x); y[o].y = fmaf(pan, acc[o].y, y[o].y);
            y[o].z = fmaf(pan, acc[o].z, y[o].z); y[o].w = fmaf(pan, acc[o].w, y[o].w);
            y0[o].x = fmaf(pan, e0[o].x, y0[o].x); y0[o].y = fmaf(pan, e0[o].y, y0[o].y);
        }
    }

    // ===== cross-group reduction into Ys (fixed order) =====
    __syncthreads();  // every TMA write has landed and been consumed: stage memory is free
    float4 *red = reinterpret_cast<float4 *>(smem);                  // [G][NOUT][LR]
    float2 *red0 = reinterpret_cast<float2 *>(red + G * NOUT * LR);  // [G][NOUT]
    if (tid < kMacConsumers) {
        const int q = tid % LR, g = tid / LR;
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            red[(g * NOUT + o) * LR + q] = y[o];
            if (q == 0) red0[g * NOUT + o] = y0[o];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NOUT * LR; idx += kFusedThreads) {
        const int o = idx / LR, q = idx % LR;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 s0 = make_float2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < G; g++) {
            const float4 v = red[(g * NOUT + o) * LR + q];
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            const float2 u = red0[g * NOUT + o];
            s0.x += u.x; s0.y += u.y;
        }
        if (q == 0) { sum.x = s0.x; sum.y = s0.y; }
        reinterpret_cast<float4 *>(Ys + (size_t)o * B)[q] = sum;
    }
    __syncthreads();

    // ===== FFT warps: inverse transform + output ring + clamp + dry mix of output fw =====
    if (fw >= 0 && fw < NOUT) {
        const int o = fw;
        const uint32_t item = inst * NOUT + o;
        WarpFft<R> f;
        f.init(a.twM);
        float2 v[R];
#pragma unroll
        for (int d = 0; d < R; d++) v[d] = Ys[(size_t)o * B + f.c + 32 * d];
        constexpr int NV4 = R >= 2 ? R / 2 : 1;
        float4 accv[NV4], xa[NV4], xb[NV4];
        float dg[2] = {0.f, 0.f};
        const int n0 = 2 * R * (lane & 15);
        float *dst = a.out + (size_t)item * B + n0;
        float *accp = nullptr;
        const bool raw = a.raw_wet != 0;
        if (lane >= 16) {
            if (!raw) {
                const InParamDev p0 = a.par[inst * a.n_in];
                const InParamDev p1 = a.par[inst * a.n_in + (a.n_in - 1)];
                dg[0] = p0.dry * pan_gain(p0.panDry, o, NOUT) * p0.level;
                dg[1] = a.n_in > 1 ? p1.dry * pan_gain(p1.panDry, o, NOUT) * p1.level : 0.f;
            }
            const float *x0 = a.in + ((size_t)inst * a.n_in) * B + n0;
            const float *x1 = x0 + (a.n_in > 1 ? B : 0);
            if (a.accring) accp = a.accring + (size_t)item * a.acc_len + (uint32_t)((t * (unsigned long long)B) & (a.acc_len - 1)) + n0;
            if constexpr (R == 1) {
                const float2 q0 = *reinterpret_cast<const float2 *>(x0), q1 = *reinterpret_cast<const float2 *>(x1);
                xa[0] = make_float4(q0.x, q0.y, 0.f, 0.f);
                xb[0] = make_float4(q1.x, q1.y, 0.f, 0.f);
                accv[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (accp) { const float2 q = *reinterpret_cast<const float2 *>(accp); accv[0].x = q.x; accv[0].y = q.y; }
            } else {
#pragma unroll
                for (int j = 0; j < NV4; j++) {
                    xa[j] = *reinterpret_cast<const float4 *>(x0 + 4 * j);
                    xb[j] = *reinterpret_cast<const float4 *>(x1 + 4 * j);
                    accv[j] = accp ? *reinterpret_cast<const float4 *>(accp + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        f.split_c2r(v, a.tw2M);
        f.inverse(v);
        if (lane >= 16) {
            auto clampf = [raw](float w) { return raw ? w : fminf(fmaxf(w, -1.0f), 1.0f); };
            if constexpr (R == 1) {
                float2 yy = make_float2(clampf(v[0].x + accv[0].x), clampf(v[0].y + accv[0].y));
                yy.x = fmaf(dg[0], xa[0].x, fmaf(dg[1], xb[0].x, yy.x));
                yy.y = fmaf(dg[0], xa[0].y, fmaf(dg[1], xb[0].y, yy.y));
                *reinterpret_cast<float2 *>(dst) = yy;
                if (accp) *reinterpret_cast<float2 *>(accp) = make_float2(0.f, 0.f);
            } else {
#pragma unroll
                for (int j = 0; j < NV4; j++) {
                    float4 yy = make_float4(clampf(v[2 * j].x + accv[j].x), clampf(v[2 * j].y + accv[j].y),
                                            clampf(v[2 * j + 1].x + accv[j].z), clampf(v[2 * j + 1].y + accv[j].w));
                    yy.x = fmaf(dg[0], xa[j].x, fmaf(dg[1], xb[j].x, yy.x));
                    yy.y = fmaf(dg[0], xa[j].y, fmaf(dg[1], xb[j].y, yy.y));
                    yy.z = fmaf(dg[0], xa[j].z, fmaf(dg[1], xb[j].z, yy.z));
                    yy.w = fmaf(dg[0], xa[j].w, fmaf(dg[1], xb[j].w, yy.w));
                    *reinterpret_cast<float4 *>(dst + 4 * j) = yy;
                    if (accp) *reinterpret_cast<float4 *>(accp + 4 * j) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// long tiers of the non-uniform partitioning (block S = 256 * 2^s_log, fired every m periods,
// AFTER the period's output has been produced: k_inverse has already advanced ctl->t)
// ------------------------------------------------------------------------------------------
constexpr int kTierThreads = 512;  // launch bound; the host launches S/16 threads clamped to [128, 512] (engine.cu: tier_div)

struct TierFwdArgs {
    const float *ring;    // [voice pool entry][ring_len]
    float2 *X;            // FDL of this tier [voice pool entry][Lring][S]
    const ItemState *st;  // [2][n_items_alloc]
    const Ctl *ctl;
    const float2 *twM, *tw2M;
    uint32_t n_items_alloc, n_in, nv, Lring, ring_len, S, s_log, m, B;
    uint32_t inst0, inst_stride;  // firing instances: inst0 + i * inst_stride
    unsigned long long tend_host;  // see MacArgs
    uint32_t t_sel;
    const float2 *rowtw;           // see FwdArgs
};

// one CTA per (firing instance, input, voice): window of the last 2S samples -> R2C -> FDL slot
__global__ void __launch_bounds__(kTierThreads, 2) k_tier_forward(const TierFwdArgs a)
{
    pdl_trigger();  // the next kernel of the stream may become resident while this one drains
    pdl_wait();     // before any global access and before any early exit: stream order holds transitively
    extern __shared__ __align__(16) float2 sm[];
    __shared__ CtaTw tw;
    // grid (voice, input, firing instance): no integer divisions in the prologue
    const uint32_t v = blockIdx.x;
    const uint32_t inst = a.inst0 + blockIdx.z * a.inst_stride;
    const uint32_t item = inst * a.n_in + blockIdx.y;
    const unsigned long long tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, 0u);
    const ItemState &st = a.st[(tend & 1ull) * a.n_items_alloc + item];
    if (!((st.active >> v) & 1u)) return;
    const uint32_t w = st.pool[v] - 1u;  // voice pool entry
    __shared__ uint32_t s_slot;
    if (threadIdx.x == 0) {  // m is a power of two; the 64-bit modulo runs once per CTA
        const unsigned long long n_fire = (tend + (inst & (a.m - 1u))) >> (31 - __clz((int)a.m));
        s_slot = (a.Lring - 1u) - (uint32_t)(n_fire % a.Lring);
    }
    const uint32_t mask = a.ring_len - 1;
    const float *ring = a.ring + (size_t)w * a.ring_len;
    const uint32_t start = (uint32_t)((tend * (unsigned long long)a.B - 2ull * a.S) & mask);
    // window load, 8 independent float4 loads in flight per thread (the loop is latency-bound otherwise)
    for (uint32_t n0 = threadIdx.x; n0 < a.S / 2; n0 += 8 * blockDim.x) {
        float4 r[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t n = n0 + u * blockDim.x;
            if (n < a.S / 2) r[u] = *reinterpret_cast<const float4 *>(ring + ((start + 4 * n) & mask));
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t n = n0 + u * blockDim.x;
            if (n < a.S / 2) *reinterpret_cast<float4 *>(sm + swz(2 * (int)n)) = r[u];
        }
    }
    cta_tw_init(tw, (int)a.S, a.twM, a.tw2M);
    cta_fft_forward(sm, (int)a.S, (int)a.s_log, tw, a.twM);
    cta_split_r2c(sm, (int)a.S, (int)a.s_log, tw);
    float2 *dst = a.X + ((size_t)w * a.Lring + s_slot) * a.S;  // written before the transform's barriers
    for (uint32_t p = threadIdx.x; p < a.S / 2; p += blockDim.x)  // position order (fft_cta.cuh), 16-byte stores
        reinterpret_cast<float4 *>(dst)[p] = *reinterpret_cast<const float4 *>(sm + swz(2 * (int)p));
}

struct TierInvArgs {
    const float2 *Ypart;  // [firing instance of this launch][n_split][n_out][S]
    float *accring;       // [inst*n_out + o][acc_len]
    const Ctl *ctl;
    const float2 *twM, *tw2M;
    uint32_t n_split, n_out, S, s_log, B, off, acc_len;
    uint32_t inst0, inst_stride;
    unsigned long long tend_host;  // see MacArgs
    uint32_t t_sel;
    const float2 *rowtw;           // see FwdArgs
};

// one CTA per (instance, output): partial-sum -> C2R -> overlap discard -> += output ring at +off
__global__ void __launch_bounds__(kTierThreads, 2) k_tier_inverse(const TierInvArgs a)
{
    pdl_trigger();  // the next kernel of the stream may become resident while this one drains
    pdl_wait();     // before any global access and before any early exit: stream order holds transitively
    extern __shared__ __align__(16) float2 sm[];
    __shared__ CtaTw tw;
    const uint32_t z = blockIdx.y, o = blockIdx.x;  // grid (output, firing instance)
    const uint32_t inst = a.inst0 + z * a.inst_stride;
    const uint32_t item = inst * a.n_out + o;
    const unsigned long long tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, 0u);
    // sum of the partial spectra (position order), 8 independent float4 loads in flight per thread
    {
        const float4 *src = reinterpret_cast<const float4 *>(a.Ypart + (((size_t)z * a.n_split) * a.n_out + o) * a.S);
        const size_t stride4 = (size_t)a.n_out * a.S / 2;
        for (uint32_t n0 = threadIdx.x; n0 < a.S / 2; n0 += 8 * blockDim.x) {
            float4 r[8];
#pragma unroll
            for (int u = 0; u < 8; u++) r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t sp = 0; sp < a.n_split; sp++) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint32_t n = n0 + u * blockDim.x;
                    if (n < a.S / 2) { const float4 q = src[sp * stride4 + n]; r[u].x += q.x; r[u].y += q.y; r[u].z += q.z; r[u].w += q.w; }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t n = n0 + u * blockDim.x;
                if (n < a.S / 2) *reinterpret_cast<float4 *>(sm + swz(2 * (int)n)) = r[u];
            }
        }
    }
    cta_tw_init(tw, (int)a.S, a.twM, a.tw2M);
    cta_split_c2r(sm, (int)a.S, (int)a.s_log, tw);
    cta_fft_inverse(sm, (int)a.S, (int)a.s_log, tw, a.twM);
    // z[n] = y[2n] + j y[2n+1]; keep y[S, 2S) = z[S/2, S): they belong to output times
    // [t_end*B - S + off, t_end*B + off)
    const uint32_t amask = a.acc_len - 1;
    const uint32_t pos0 = (uint32_t)((tend * (unsigned long long)a.B - a.S + a.off) & amask);
    float *acc = a.accring + (size_t)item * a.acc_len;
    // read-modify-write of the output ring in float4s (S/4 of them), 8 loads in flight per thread
    for (uint32_t n0 = threadIdx.x; n0 < a.S / 4; n0 += 8 * blockDim.x) {
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t n = n0 + u * blockDim.x;
            if (n < a.S / 4) q[u] = *reinterpret_cast<const float4 *>(acc + ((pos0 + 4 * n) & amask));
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t n = n0 + u * blockDim.x;
            if (n < a.S / 4) {
                const float4 z = *reinterpret_cast<const float4 *>(sm + swz((int)(a.S / 2 + 2 * n)));
                q[u].x += z.x; q[u].y += z.y; q[u].z += z.z; q[u].w += z.w;
                *reinterpret_cast<float4 *>(acc + ((pos0 + 4 * n) & amask)) = q[u];
            }
        }
    }
}

struct TierIrArgs {
    const float *h[2];
    float2 *H;  // this slot's spectra of this tier [n_out][P][S]
    const float2 *twM, *tw2M;
    uint32_t frames, P, n_out, frame_off, stride, S, s_log;
    float scale;
};

// one CTA per (output channel, partition): [h_k | 0] -> spectrum
__global__ void __launch_bounds__(kTierThreads) k_tier_ir(const TierIrArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    __shared__ CtaTw tw;
    const uint32_t o = blockIdx.x / a.P, k = blockIdx.x % a.P;
    const float *h = o == 0 ? a.h[0] : a.h[1];
    for (uint32_t n = threadIdx.x; n < a.S; n += blockDim.x) {
        float re = 0.f, im = 0.f;
        if (n < a.S / 2) {
            const size_t n0 = (size_t)a.frame_off + (size_t)k * a.S + 2 * n;
            if (n0 < a.frames) re = __ldg(&h[n0 * a.stride]) * a.scale;
            if (n0 + 1 < a.frames) im = __ldg(&h[(n0 + 1) * a.stride]) * a.scale;
        }
        sm[swz((int)n)] = make_float2(re, im);
    }
    cta_tw_init(tw, (int)a.S, a.twM, a.tw2M);
    cta_fft_forward(sm, (int)a.S, (int)a.s_log, tw, a.twM);
    cta_split_r2c(sm, (int)a.S, (int)a.s_log, tw);
    float2 *dst = a.H + ((size_t)o * a.P + k) * a.S;
    for (uint32_t p = threadIdx.x; p < a.S; p += blockDim.x) dst[p] = sm[swz((int)p)];
}

}  // namespace ca
