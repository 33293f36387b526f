// engine.cu -- host runtime + C ABI (include/cuda_audio_b200.h) of the B200 convolution engine.
//
// Owns the device memory (IR spectra bank, frequency-domain delay lines, time rings, parameter
// blocks), the CUDA stream / graphs of the per-period pipeline and the statistics.
// No cuFFT, no CPU fallback: if CUDA is not usable every entry point returns CA_ERR_CUDA.
#include "../../include/cuda_audio_b200.h"
#include "kernels.cuh"
#include "kernels_rows.cuh"
#include "kernels_rows16.cuh"
#include "kernels_quirks.cuh"
#include "kernels_persist.cuh"
#include "param_queue.h"

#include <cuda.h>  // types + prototypes only: driver entry points are resolved through the runtime (no libcuda link dependency)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace ca;

namespace {

thread_local std::string g_last_error;

#define CA_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _rc = (expr);                                                                  \
        if (_rc != cudaSuccess) {                                                                  \
            char _buf[512];                                                                        \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_rc)); \
            g_last_error = _buf;                                                                   \
            return _rc == cudaErrorMemoryAllocation ? CA_ERR_NOMEM : CA_ERR_CUDA;                  \
        }                                                                                          \
    } while (0)

// Streams, events and graphs belong to the engine's GPU: a launch into them fails when the calling thread's
// current device is another one (a JACK callback thread starts on device 0 whatever GPU its engine was built
// on: engine.gpus).  One thread-local read when the device is already current.
#define CA_BIND(e)                                                                                 \
    do {                                                                                           \
        int _cur = -1;                                                                             \
        if (cudaGetDevice(&_cur) != cudaSuccess || _cur != (e)->device) CA_CUDA(cudaSetDevice((e)->device)); \
    } while (0)

typedef void (*fwd_fn)(const FwdArgs);
typedef void (*ir_fn)(const IrArgs);
typedef void (*mac_fn)(const MacArgs);
typedef void (*macp_fn)(const MacArgs, const uint32_t, const uint32_t);
typedef void (*inv_fn)(const InvArgs);

// Launch with (pdl) or without the programmatic-stream-serialization attribute, see pdl_wait() in kernels.cuh.
template <class... P, class... A>
inline void launch_k(bool pdl, void (*fn)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const A &...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, fn, P(args)...);
}

// complex points per thread of the CTA-level tier FFT: threads = clamp(S / div, min, 512).  Measured r01
// (K = 4096, tier forward + inverse per period): div 4: 149 us, 8: 134, 16: 123 (128 threads for the 2 K
// tier: twice the CTAs per SM, one wave instead of 1.7), 32 with min 64: 128; 12 288 instances: +3.9 %.
inline uint32_t tier_div()
{
    static const uint32_t d = [] { const char *v = getenv("CA_TIER_DIV"); const int x = v ? atoi(v) : 16; return (uint32_t)(x >= 2 && x <= 64 ? x : 16); }();
    return d;
}

inline uint32_t tier_min()
{
    static const uint32_t d = [] { const char *v = getenv("CA_TIER_MIN"); const int x = v ? atoi(v) : 128; return (uint32_t)(x >= 32 && x <= 512 ? x : 128); }();
    return d;
}

struct MacVariant { mac_fn fn; uint32_t smem; int kc; macp_fn pfn; uint32_t psmem; };

template <int BT, int NOUT, int MULT, int NSTAGE>
MacVariant mac_variant()
{
    constexpr int G = kMacConsumers / (BT / 2);
    using Cfg = MacCfg<BT, NOUT, G * MULT, NSTAGE>;
    return MacVariant{k_mac<BT, NOUT, G * MULT, NSTAGE>, Cfg::SMEM_BYTES, G * MULT,
                      k_mac_p<BT, NOUT, G * MULT, NSTAGE>, MacPCfg<BT, NOUT, G * MULT, NSTAGE>::SMEM_BYTES};
}

// stage = G*MULT rows of (1 + NOUT) arrays of BT complex; BT = 256, NOUT = 2: 12 KB * MULT
template <int BT, int NOUT>
MacVariant mac_pick_v(int variant)
{
    switch (variant) {
    case 0: return mac_variant<BT, NOUT, 4, 4>();   // 192 KB: 1 CTA / SM
    case 2: return mac_variant<BT, NOUT, 2, 3>();   //  72 KB: 3 CTAs / SM
    case 3: return mac_variant<BT, NOUT, 4, 2>();   //  96 KB, longer stages
    case 4: return mac_variant<BT, NOUT, 1, 4>();   //  48 KB: 4 CTAs / SM
    case 5: return mac_variant<BT, NOUT, 2, 6>();   // 144 KB: 1 CTA / SM, deep
    case 6: return mac_variant<BT, NOUT, 1, 3>();   //  36 KB: 6 CTAs / SM
    case 7: return mac_variant<BT, NOUT, 1, 2>();   //  24 KB: 9 CTAs / SM
    case 12: return mac_variant<BT, NOUT, 2, 2>();  //  48 KB: two stages of twice the rows
    case 13: return mac_variant<BT, NOUT, 4, 1>();  //  48 KB: ONE stage
    case 14: return mac_variant<BT, NOUT, 2, 1>();  //  24 KB: ONE stage
    case 15: return mac_variant<BT, NOUT, 8, 1>();  //  96 KB: ONE stage, 2 CTAs / SM
    case 16: return mac_variant<BT, NOUT, 6, 1>();  //  72 KB: ONE stage, 3 CTAs / SM
    case 17: return mac_variant<BT, NOUT, 3, 1>();  //  36 KB: ONE stage
    default: return mac_variant<BT, NOUT, 2, 4>();  //  96 KB: 2 CTAs / SM (measured best)
    }
}

MacVariant mac_pick(int bt, int n_out, int variant)
{
    switch (bt) {
    case 32: return n_out == 1 ? mac_pick_v<32, 1>(variant) : mac_pick_v<32, 2>(variant);
    case 64: return n_out == 1 ? mac_pick_v<64, 1>(variant) : mac_pick_v<64, 2>(variant);
    case 128: return n_out == 1 ? mac_pick_v<128, 1>(variant) : mac_pick_v<128, 2>(variant);
    case 512: return n_out == 1 ? mac_pick_v<512, 1>(variant) : mac_pick_v<512, 2>(variant);
    default: return n_out == 1 ? mac_pick_v<256, 1>(variant) : mac_pick_v<256, 2>(variant);
    }
}

typedef void (*fused_fn)(const FusedArgs);
struct FftFns { fwd_fn fwd; ir_fn ir; inv_fn inv, inv_packed; fused_fn fused1, fused2; };
FftFns fft_pick(int R)
{
    switch (R) {
    case 1: return {k_forward<1>, k_ir_fft<1>, k_inverse<1, false>, k_inverse<1, true>, k_fused0<1, 1>, k_fused0<1, 2>};
    case 2: return {k_forward<2>, k_ir_fft<2>, k_inverse<2, false>, k_inverse<2, true>, k_fused0<2, 1>, k_fused0<2, 2>};
    case 4: return {k_forward<4>, k_ir_fft<4>, k_inverse<4, false>, k_inverse<4, true>, k_fused0<4, 1>, k_fused0<4, 2>};
    case 8: return {k_forward<8>, k_ir_fft<8>, k_inverse<8, false>, k_inverse<8, true>, k_fused0<8, 1>, k_fused0<8, 2>};
    case 16: return {k_forward<16>, k_ir_fft<16>, k_inverse<16, false>, k_inverse<16, true>, nullptr, nullptr};
    case 32: return {k_forward<32>, k_ir_fft<32>, k_inverse<32, false>, k_inverse<32, true>, nullptr, nullptr};
    default: return {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    }
}

double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }
uint32_t ilog2(uint32_t v) { uint32_t l = 0; while ((1u << l) < v) l++; return l; }

// One tier of the (non-)uniform partitioning.  Tier 0 has block S = period and runs inside every
// period; tier j >= 1 has block S_j = m * period, covers IR frames [off, off + P*S) and runs after
// the output of every m-th period has been produced (its result is first needed one period later
// because off >= S).

constexpr uint32_t kIoChunks = 16;  // upper bound; the count used is io_chunks (default 3, env CA_IO_CHUNKS)

struct Tier {
    uint32_t S = 0, m = 1, P = 0, off = 0, s_log = 0, bt = 0, tiles = 1, n_split = 1, Lring = 0;
    float2 *H = nullptr, *X = nullptr, *Ypart = nullptr, *Ypart2 = nullptr, *tw = nullptr;  // Ypart2: odd periods of the pipelined schedule
    size_t h_bytes = 0, x_bytes = 0;
    MacVariant mac{};
    uint32_t p_slots = 0;  // resident CTA slots of the persistent MAC (0: schedule disabled)
    uint32_t *workctr = nullptr;  // persistent MAC: [next work item, finished CTAs]
    bool p_force = false;
    // FFT family of this tier's transforms: 0 = fft_warp / fft_cta (general), 1 = row FFTs, one CTA per
    // transform (M1 = S / 256 <= 16), 2 = row FFTs, columns and rows as two launches (M1 = 32, 64)
    int rows_mode = 0;
    uint32_t M1 = 0;
};

}  // namespace

struct ca_engine {
    ca_config cfg{};
    int device = 0;
    uint32_t B = 0, R = 0, k_off = 0;
    uint32_t n_inst = 0, n_active = 0, n_in = 0, n_out = 0, nv = 2, ring_len = 16384, ring_out = 0, acc_len = 0;
    std::vector<Tier> tiers;
    cudaStream_t stream = nullptr;
    // device memory
    unsigned char *d_arena = nullptr;  // every tier's [H | X], contiguous (one L2 access-policy window)
    size_t arena_bytes = 0;
    float *d_ring = nullptr, *d_acc = nullptr, *d_in = nullptr, *d_out = nullptr;
    InParamDev *d_par = nullptr;
    ItemState *d_st = nullptr;
    Ctl *d_ctl = nullptr;
    uint64_t device_bytes = 0;
    // pinned host staging
    float *h_in = nullptr, *h_out = nullptr;
    InParamDev *h_upload[2] = {nullptr, nullptr};
    cudaEvent_t upload_done[2] = {nullptr, nullptr};
    cudaEvent_t out_ready = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the chunked host pipeline
    cudaStream_t s_tier[CA_MAX_TIERS] = {};         // side streams: forward FFT + MAC of concurrently firing tiers
    cudaEvent_t fork_ev = nullptr, join_ev[CA_MAX_TIERS] = {};
    cudaEvent_t io_ev[2][kIoChunks + 1] = {};
    // pipelined batch schedule (run_pipelined): lane C = stream (FFT kernels), lane M = s_mac (MAC kernels)
    cudaStream_t s_mac = nullptr;
    cudaEvent_t pf_ev[kIoChunks + 1] = {}, pm_ev[kIoChunks + 1] = {};  // forward done / tier-0 MAC done, per chunk
    cudaEvent_t ptf_ev[CA_MAX_TIERS] = {}, ptm_ev[CA_MAX_TIERS][2] = {};  // tier forward done / tier MAC done (by period parity)
    cudaEvent_t pm_tail = nullptr;   // last launch on lane M
    cudaEvent_t ptinv_ev[CA_MAX_TIERS] = {};  // tier inverse done
    bool tinv_pending[CA_MAX_TIERS] = {};
    uint64_t pipe_prev_tend = 0;     // period whose long-tier inverse transforms are still to be launched (0: none)
    bool pipe_used = false;
    // CA_FLAG_ASYNC_TIERS: deferred tiers on their own low-priority stream, two periods of slack
    bool async_tiers = false, async_used = false;
    cudaStream_t s_def = nullptr;
    cudaEvent_t ev_period = nullptr, ev_def[2] = {};
    bool pdl = false;                // programmatic dependent launch between the period's kernels (CA_PDL; default: batches without a graph)
    int pipe_mode = 0;               // CA_PIPELINE=1 enables it (measured r01: +5 % device-resident, -8 % end to end: off by default)
    // CA_PIPE_TRACE=n: print the device timeline (CUDA events around every launch) of periods n, n+1
    struct TraceEv { const char *name; int lane; cudaEvent_t a, b; };
    std::vector<TraceEv> trace;
    uint64_t trace_at = 0;
    bool tracing = false;
    uint32_t io_chunks = 3;  // measured at 16 128 instances (16 x 16 row kernels): 2 chunks 1.189 ms, 3: 1.159, 4: 1.172 (device-resident 1.095)
    int upload_idx = 0;
    // parameters: `par` (host shadow of the device blocks) belongs to the processing thread; setters reach it
    // through the lock-free command ring; `user` (what ca_get_params returns) is seqlock-guarded per item
    ParamQueue par_queue;
    std::vector<InParamDev> par;
    ParamMirror user;
    bool par_dirty = true;  // processing thread only
    std::unique_ptr<std::atomic<uint8_t>[]> ir_loaded;  // set by ca_load_ir*, read by ca_set_params on any thread
    FftFns fft{};
    // member of a ca_group (one IR split by partition range across GPUs, csrc/group.cuh)
    struct Link {
        float2 *gather = nullptr;             // MAC epilogue: this device's slot of the root's gather buffer
        unsigned long long *gflag = nullptr;  // ... and its flag (root memory)
        uint32_t *gcount = nullptr;           // local CTA counter of the MAC launch
        bool skip_inverse = false;            // peers stop after the MAC (the root transforms the sum)
        bool defer_inverse = false;           // NCCL variant, root: the inverse is launched after the reduce (launch_inverse_phase)
        const float2 *inv_src = nullptr;      // root: inverse reads the gather buffer ...
        uint32_t inv_split = 0;               // ... of this many spectra
        const unsigned long long *wait_flags = nullptr;
        uint32_t n_wait = 0;
        int *gerr = nullptr;
    } link;
    float2 *d_rowtw = nullptr;  // [W_256^n | W_512^k]: twiddles of the 256-point row FFT
    // CA_FLAG_REF_QUIRKS (kernels_quirks.cuh)
    bool quirks = false;
    double *d_irsum = nullptr, *d_qdelta = nullptr, *d_qrun = nullptr;
    float *d_qring = nullptr;
    uint32_t q_kr = 0, q_len = 0;
    // CA_FLAG_PERSISTENT (kernels_persist.cuh)
    bool persistent = false, p_running = false;
    PersistBox *pbox = nullptr;            // mapped pinned host memory
    unsigned long long *d_go = nullptr;
    unsigned int *d_arrive = nullptr;
    InParamDev *d_ppar = nullptr;
    float2 *d_pY = nullptr, *d_pYsum = nullptr;
    bool p_stamp = false;
    uint32_t p_ctas = 0, p_gen = 0;
    // SM partitioning (green contexts): MAC lane on sm_mac SMs, FFT lanes on the rest
    CUgreenCtx g_mac = nullptr, g_fft = nullptr;
    uint32_t sm_mac = 0, sm_fft = 0;
    uint32_t *d_vpool = nullptr;  // bitmap of the shared cross-fade voice entries
    uint32_t n_voices = 0, n_extra = 0;  // voice pool: n_items homes + n_extra shared entries
    bool rows0 = false;         // tier 0 (period 256) on the row-FFT kernels
    bool rows16 = true;         // row FFTs as 16 x 16, two rows per warp (kernels_rows16.cuh); false: 8 x 8 x 4 (CA_SCHED_ROWS8)
    bool fused = false;       // tier 0 runs as one fused kernel (k_fused0)
    uint32_t fused_smem = 0;
    // graphs: [0] = the period pipeline (tier 0), [mask] = the deferred tiers that fire together
    std::map<uint64_t, cudaGraphExec_t> graphs;
    const float *g_in = nullptr;
    float *g_out = nullptr;
    uint32_t g_active = 0;
    // profiling
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t tev[CA_MAX_TIERS][4] = {};
    bool tev_used[CA_MAX_TIERS] = {};
    double prof_us[4] = {0, 0, 0, 0};
    double tier_prof_us[3] = {0, 0, 0};  // long tiers: forward FFT, MAC, inverse FFT
    uint64_t prof_n = 0;
    // stats
    std::vector<float> wall;  // ring of host wall times (us)
    uint64_t periods = 0, xruns = 0, launches = 0, t_host = 0;
    double wall_sum = 0, wall_max = 0, deadline_us = 0;
    // pinned-pointer cache
    const void *pin_ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    bool pin_val[4] = {false, false, false, false};
    int pin_next = 0;
};

namespace {

bool is_pinned(ca_engine *e, const void *p)
{
    for (int i = 0; i < 4; i++)
        if (e->pin_ptr[i] == p) return e->pin_val[i];
    cudaPointerAttributes at{};
    bool pinned = false;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess) pinned = (at.type == cudaMemoryTypeHost);
    else (void)cudaGetLastError();
    e->pin_ptr[e->pin_next] = p;
    e->pin_val[e->pin_next] = pinned;
    e->pin_next = (e->pin_next + 1) & 3;
    return pinned;
}

void fill_dev_param(InParamDev &d, const ca_params &u)
{
    d.wet = u.wet; d.dry = u.dry; d.level = u.level; d.panWet = u.panWet; d.panDry = u.panDry;
    d.predelay = u.predelay; d.select = u.select;
}

// processing thread: drain the command ring into the host shadow; true when anything changed
bool drain_params(ca_engine *e)
{
    ParamCmd c;
    while (e->par_queue.pop(&c)) {
        InParamDev &d = e->par[c.item];
        if (c.kind == 0) {
            fill_dev_param(d, c.p);
            if (c.p.vsteps >= 0) { d.vsteps_cmd = (uint32_t)c.p.vsteps; d.vsteps_seq++; }
        } else {
            d.glide_cmd = c.glide;
            d.glide_seq++;
        }
        e->par_dirty = true;
    }
    return e->par_dirty;
}

int flush_params(ca_engine *e)
{
    if (!drain_params(e)) return CA_OK;
    const int b = e->upload_idx;
    e->upload_idx ^= 1;
    CA_CUDA(cudaEventSynchronize(e->upload_done[b]));  // the upload two changes ago: long finished
    const size_t bytes = e->par.size() * sizeof(InParamDev);
    memcpy(e->h_upload[b], e->par.data(), bytes);
    e->par_dirty = false;
    CA_CUDA(cudaMemcpyAsync(e->d_par, e->h_upload[b], bytes, cudaMemcpyHostToDevice, e->stream));
    CA_CUDA(cudaEventRecord(e->upload_done[b], e->stream));
    return CA_OK;
}

VoicePool voice_pool(const ca_engine *e)
{
    return VoicePool{e->d_vpool, (e->n_extra + 31u) / 32u, e->n_inst * e->n_in, e->n_extra};
}

MacArgs mac_args(ca_engine *e, const Tier &t, uint32_t t_bias)
{
    return MacArgs{t.X, t.H, t.Ypart, e->d_par, e->d_st, e->d_ctl, e->n_inst * e->n_in, e->n_in, e->nv, t.Lring, t.P, t.S,
                   e->k_off, t.m, t_bias, t.n_split, (e->cfg.flags & CA_FLAG_STREAMING) ? 1u : 0u, 0u, 1u, t.m > 1 ? 1u : 0u};
}

MacArgs with_workctr(MacArgs ma, const Tier &t, uint32_t set) { ma.work_ctr = t.workctr + 2u * set; return ma; }  // two sets: launches of alternating chunks may overlap
constexpr uint32_t kMacDynamicItems = 10;  // work items per CTA from which the persistent MAC hands them out dynamically

// One MAC launch over `count` instances of tier t.  Batches (n_split == 1 and more work items than
// resident CTA slots) take the persistent schedule: every CTA gets the same number of work items.
void launch_mac(const Tier &t, const MacArgs &ma, uint32_t count, cudaStream_t st, bool pdl = false, uint32_t ctr_set = 0)
{
    const uint32_t n_work = count * t.tiles;
    if (t.p_slots && t.n_split == 1 && (t.p_force || n_work > t.p_slots)) {
        const uint32_t per = (n_work + t.p_slots - 1) / t.p_slots;
        // many items per CTA: hand them out with a counter (no tail of late CTAs: -4 % per period at 16 128 instances);
        // few: static stride (the counter costs ~3 % when every CTA has the same 7 items anyway)
        const bool dynamic = per >= kMacDynamicItems;
        launch_k(pdl, t.mac.pfn, dim3((n_work + per - 1) / per), dim3(kMacThreads), t.mac.psmem, st, dynamic ? with_workctr(ma, t, ctr_set) : ma, n_work, t.tiles);
    } else {
        launch_k(pdl, t.mac.fn, dim3(t.n_split, t.tiles, count), dim3(kMacThreads), t.mac.smem, st, ma);
    }
}

// ---- FFT launches: row-FFT family (kernels_rows.cuh) where the size allows, else the general kernels ----
void launch_fwd0(ca_engine *e, bool pdl, FwdArgs fa, cudaStream_t st)
{
    fa.rowtw = e->d_rowtw;
    fa.vp = voice_pool(e);
    if (e->rows0 && e->rows16) launch_k(pdl, k_fwd0_x2, dim3((fa.n_items + kX2Rows - 1) / kX2Rows), dim3(kX2Threads), kX2Smem, st, fa);
    else if (e->rows0) launch_k(pdl, k_fwd0_rows, dim3((fa.n_items + kRowsWarps - 1) / kRowsWarps), dim3(kRowsThreads), kRowsSmem, st, fa);
    else launch_k(pdl, e->fft.fwd, dim3((fa.n_items + kFwdWarps - 1) / kFwdWarps), dim3(kFwdWarps * 32), 0, st, fa);
}

void launch_inv0(ca_engine *e, bool pdl, InvArgs ia, cudaStream_t st)
{
    ia.rowtw = e->d_rowtw;
    if (ia.n_split <= 4) {
        if (e->rows0 && e->rows16) launch_k(pdl, k_inv0_x2, dim3((ia.n_items + kX2Rows - 1) / kX2Rows), dim3(kX2Threads), kX2Smem, st, ia);
        else if (e->rows0) launch_k(pdl, k_inv0_rows, dim3((ia.n_items + kRowsWarps - 1) / kRowsWarps), dim3(kRowsThreads), kRowsSmem, st, ia);
        else launch_k(pdl, e->fft.inv_packed, dim3((ia.n_items + kInvThreads / 32 - 1) / (kInvThreads / 32)), dim3(kInvThreads), 0, st, ia);
    } else {
        launch_k(pdl, e->fft.inv, dim3(ia.n_items), dim3(kInvThreads), 0, st, ia);  // latency schedule: one CTA sums up to 256 partials
    }
}

uint32_t tier_threads(const Tier &t) { return std::min<uint32_t>(kTierThreads, std::max<uint32_t>(tier_min(), t.S / tier_div())); }

template <int M1>
void launch_tfwd_rows(const Tier &t, bool x2, bool pdl, const TierFwdArgs &fa, uint32_t nv, uint32_t n_in, uint32_t count, cudaStream_t st)
{
    if constexpr (M1 <= 16) {
        if constexpr (M1 >= 2) {
            if (x2) { launch_k(pdl, k_tfwd_x2<M1>, dim3(nv, n_in, count), dim3(M1 * 16), M1 * kR16Slots * sizeof(float2), st, fa); return; }
        }
        launch_k(pdl, k_tfwd_fused<M1>, dim3(nv, n_in, count), dim3(M1 * 32), M1 * kRowSlots * sizeof(float2), st, fa);
    } else {
        launch_k(pdl, k_tcols_fwd<M1>, dim3(nv * 8, n_in, count), dim3(256), 0, st, fa);
        if (x2) launch_k(pdl, k_trows_fwd_x2<M1>, dim3(nv * (M1 / 16), n_in, count), dim3(kX2Threads), kX2Smem, st, fa);
        else launch_k(pdl, k_trows_fwd<M1>, dim3(nv * (M1 / 8), n_in, count), dim3(kRowsThreads), kRowsSmem, st, fa);
    }
}

template <int M1>
void launch_tinv_rows(const Tier &t, bool x2, bool pdl, const TierInvArgs &ia, uint32_t n_out, uint32_t count, cudaStream_t st)
{
    if constexpr (M1 <= 16) {
        if constexpr (M1 >= 2) {
            if (x2) { launch_k(pdl, k_tinv_x2<M1>, dim3(n_out, count), dim3(M1 * 16), M1 * kR16Slots * sizeof(float2), st, ia); return; }
        }
        launch_k(pdl, k_tinv_fused<M1>, dim3(n_out, count), dim3(M1 * 32), M1 * kRowSlots * sizeof(float2), st, ia);
    } else {
        if (x2) launch_k(pdl, k_trows_inv_x2<M1>, dim3(n_out * (M1 / 16), count), dim3(kX2Threads), kX2Smem, st, ia);
        else launch_k(pdl, k_trows_inv<M1>, dim3(n_out * (M1 / 8), count), dim3(kRowsThreads), kRowsSmem, st, ia);
        launch_k(pdl, k_tcols_inv<M1>, dim3(n_out * 8, count), dim3(256), 0, st, ia);
    }
}

// forward transforms of tier t for `count` firing instances; returns the number of kernels launched
uint32_t launch_tier_fwd(ca_engine *e, const Tier &t, bool pdl, TierFwdArgs fa, uint32_t count, cudaStream_t st)
{
    fa.rowtw = e->d_rowtw;
    switch (t.rows_mode ? t.M1 : 0u) {
    case 1: launch_tfwd_rows<1>(t, e->rows16, pdl, fa, e->nv, e->n_in, count, st); return 1;
    case 2: launch_tfwd_rows<2>(t, e->rows16, pdl, fa, e->nv, e->n_in, count, st); return 1;
    case 4: launch_tfwd_rows<4>(t, e->rows16, pdl, fa, e->nv, e->n_in, count, st); return 1;
    case 8: launch_tfwd_rows<8>(t, e->rows16, pdl, fa, e->nv, e->n_in, count, st); return 1;
    case 16: launch_tfwd_rows<16>(t, e->rows16, pdl, fa, e->nv, e->n_in, count, st); return 1;
    case 32: launch_tfwd_rows<32>(t, e->rows16, pdl, fa, e->nv, e->n_in, count, st); return 2;
    case 64: launch_tfwd_rows<64>(t, e->rows16, pdl, fa, e->nv, e->n_in, count, st); return 2;
    default: launch_k(pdl, k_tier_forward, dim3(e->nv, e->n_in, count), dim3(tier_threads(t)), t.S * sizeof(float2), st, fa); return 1;
    }
}

uint32_t launch_tier_inv(ca_engine *e, const Tier &t, bool pdl, TierInvArgs ia, uint32_t count, cudaStream_t st)
{
    ia.rowtw = e->d_rowtw;
    switch (t.rows_mode ? t.M1 : 0u) {
    case 1: launch_tinv_rows<1>(t, e->rows16, pdl, ia, e->n_out, count, st); return 1;
    case 2: launch_tinv_rows<2>(t, e->rows16, pdl, ia, e->n_out, count, st); return 1;
    case 4: launch_tinv_rows<4>(t, e->rows16, pdl, ia, e->n_out, count, st); return 1;
    case 8: launch_tinv_rows<8>(t, e->rows16, pdl, ia, e->n_out, count, st); return 1;
    case 16: launch_tinv_rows<16>(t, e->rows16, pdl, ia, e->n_out, count, st); return 1;
    case 32: launch_tinv_rows<32>(t, e->rows16, pdl, ia, e->n_out, count, st); return 2;
    case 64: launch_tinv_rows<64>(t, e->rows16, pdl, ia, e->n_out, count, st); return 2;
    default: launch_k(pdl, k_tier_inverse, dim3(e->n_out, count), dim3(tier_threads(t)), t.S * sizeof(float2), st, ia); return 1;
    }
}

// kernels of one firing of tier t (forward + MAC + inverse)
uint32_t tier_launch_count(const Tier &t) { return t.rows_mode == 2 ? 5u : 3u; }

// instances whose tier-j block closes at the end of period t_end - 1: s = r + i*m, r = (-t_end) mod m
uint32_t tier_residue(const Tier &t, uint64_t tend) { return (uint32_t)((t.m - tend % t.m) % t.m); }
uint32_t tier_count(const ca_engine *e, const Tier &t, uint64_t tend)
{
    const uint32_t r = tier_residue(t, tend);
    return e->n_active > r ? (e->n_active - r + t.m - 1) / t.m : 0u;
}

// tier 0: the period pipeline for instances [i0, i1).  After the launch that covers the last
// instance (last == true) the output block is complete and the device period counter advances.
int launch_period(ca_engine *e, const float *d_in, float *d_out, bool profile, uint32_t i0, uint32_t i1, bool last, cudaStream_t st = nullptr, uint32_t ctr_set = 0)
{
    if (!st) st = e->stream;
    const Tier &t0 = e->tiers[0];
    const uint32_t n_items = (i1 - i0) * e->n_in;
    const uint32_t n_alloc = e->n_inst * e->n_in;
    // without a graph the host passes the period count itself (e->t_host == ctl->t between periods)
    const unsigned long long tp1 = (e->cfg.flags & CA_FLAG_GRAPH) ? 0ull : e->t_host + 1ull;
    FwdArgs fa{d_in, e->d_ring, t0.X, e->d_par, e->d_st, e->d_ctl, t0.tw, t0.tw + e->B,
               n_items, n_alloc, e->n_in, e->nv, t0.Lring, e->ring_len, e->ring_out, i0 * e->n_in, tp1};
    MacArgs ma = mac_args(e, t0, 1u);
    ma.inst0 = i0; ma.tend_host = tp1;
    ma.gather = e->link.gather; ma.gflag = e->link.gflag; ma.gcount = e->link.gcount;
    InvArgs ia{t0.Ypart, d_in, d_out, e->tiers.size() > 1 ? e->d_acc : nullptr, e->d_par, e->d_ctl, t0.tw, t0.tw + e->B,
               t0.n_split, e->n_in, e->n_out, e->acc_len, (i1 - i0) * e->n_out, i0 * e->n_out, last ? 1u : 0u, ((e->cfg.flags & CA_FLAG_RAW_WET) || e->quirks) ? 1u : 0u, tp1};
    if (e->link.inv_src) { ia.Ypart = e->link.inv_src; ia.n_split = e->link.inv_split; ia.wait_flags = e->link.wait_flags; ia.n_wait = e->link.n_wait; ia.gerr = e->link.gerr; }
    if (e->fused) {
        // tiered throughput schedule: forward + MAC + inverse of tier 0 in one CTA per instance
        fused_fn fn = e->n_out == 1 ? e->fft.fused1 : e->fft.fused2;
        FusedArgs ga{d_in, d_out, e->d_ring, t0.X, t0.H, e->tiers.size() > 1 ? e->d_acc : nullptr, e->d_par, e->d_st, e->d_ctl, t0.tw, t0.tw + e->B,
                     n_alloc, e->n_in, e->nv, t0.Lring, t0.P, e->ring_len, e->ring_out, e->acc_len, i0,
                     (e->cfg.flags & CA_FLAG_STREAMING) ? 1u : 0u, (e->cfg.flags & CA_FLAG_RAW_WET) ? 1u : 0u, voice_pool(e)};
        if (profile) { CA_CUDA(cudaEventRecord(e->ev[0], st)); CA_CUDA(cudaEventRecord(e->ev[1], st)); }
        fn<<<i1 - i0, kFusedThreads, e->fused_smem, st>>>(ga);
        if (last) k_tick<<<1, 1, 0, st>>>(e->d_ctl);
        if (profile) { CA_CUDA(cudaEventRecord(e->ev[2], st)); CA_CUDA(cudaEventRecord(e->ev[3], st)); }
        CA_CUDA(cudaGetLastError());
        return CA_OK;
    }
    if (profile) CA_CUDA(cudaEventRecord(e->ev[0], st));
    const bool pdl = e->pdl && !profile;
    launch_fwd0(e, pdl, fa, st);
    if (profile) CA_CUDA(cudaEventRecord(e->ev[1], st));
    launch_mac(t0, ma, i1 - i0, st, pdl, ctr_set);
    if (profile) CA_CUDA(cudaEventRecord(e->ev[2], st));
    if (e->link.skip_inverse) { if (last) k_tick<<<1, 1, 0, st>>>(e->d_ctl); }  // group peer: only the period counter advances
    else if (!e->link.defer_inverse) launch_inv0(e, pdl, ia, st);
    if (e->quirks) {  // reference-compatible DC / Nyquist terms, then clamp + dry mix (the inverse above stored the raw wet block)
        QuirkArgs qa{d_in, d_out, e->d_par, e->d_st, e->d_ctl, e->d_irsum, e->d_qdelta, e->d_qrun, e->d_qring,
                     n_alloc, e->nv, e->B, e->cfg.ref_fft_size, e->q_kr, e->q_len, i0, i1 - i0, tp1};
        k_ref_quirks<<<(i1 - i0 + kQuirkWarps - 1) / kQuirkWarps, kQuirkWarps * 32, 0, st>>>(qa);
    }
    if (profile) CA_CUDA(cudaEventRecord(e->ev[3], st));
    CA_CUDA(cudaGetLastError());
    return CA_OK;
}

// ca_group, NCCL variant, root: the inverse transform of the period whose forward + MAC were launched by
// launch_period (link.defer_inverse), after the reduce has delivered the summed spectrum into link.inv_src
int launch_inverse_phase(ca_engine *e, const float *d_in, float *d_out)
{
    const Tier &t0 = e->tiers[0];
    const unsigned long long tp1 = e->t_host + 1ull;
    InvArgs ia{e->link.inv_src, d_in, d_out, nullptr, e->d_par, e->d_ctl, t0.tw, t0.tw + e->B,
               e->link.inv_split, e->n_in, e->n_out, e->acc_len, e->n_active * e->n_out, 0u, 1u, (e->cfg.flags & CA_FLAG_RAW_WET) ? 1u : 0u, tp1};
    launch_inv0(e, false, ia, e->stream);
    e->launches += 1;
    CA_CUDA(cudaGetLastError());
    return CA_OK;
}

// deferred tiers: for every tier the phase-staggered subset of instances whose block just closed
int launch_tiers(ca_engine *e, uint64_t tend, bool profile, cudaStream_t base = nullptr)
{
    if (!base) base = e->stream;
    const uint32_t n_alloc = e->n_inst * e->n_in;
    uint32_t firing = 0;
    for (size_t j = 1; j < e->tiers.size(); j++) firing += tier_count(e, e->tiers[j], tend) ? 1u : 0u;
    // Different tiers work on disjoint buffers until their inverse adds into the shared output ring:
    // with >= 2 tiers firing, each tier's forward FFT + MAC runs on its own side stream (the 16 K tier
    // launches too few CTAs to fill the GPU alone), then the inverse kernels run in tier order on the
    // main stream so the output-ring accumulation stays race-free and deterministic.
    const bool fork = firing >= 2 && !profile && !e->async_tiers;
    const unsigned long long th = (e->cfg.flags & CA_FLAG_GRAPH) ? 0ull : tend;  // graphs replay: the kernels read ctl->t ...
    const uint32_t tsel = e->async_tiers ? 1u + (uint32_t)(tend & 1) : 0u;       // ... or its parity slot (asynchronous tiers)
    if (fork) CA_CUDA(cudaEventRecord(e->fork_ev, base));
    for (size_t j = 1; j < e->tiers.size(); j++) {
        const Tier &t = e->tiers[j];
        const uint32_t count = tier_count(e, t, tend), r = tier_residue(t, tend);
        e->tev_used[j] = profile && count;
        if (!count) continue;
        cudaStream_t st = fork ? e->s_tier[j] : base;
        if (fork) CA_CUDA(cudaStreamWaitEvent(st, e->fork_ev, 0));
        TierFwdArgs fa{e->d_ring, t.X, e->d_st, e->d_ctl, t.tw, t.tw + t.S, n_alloc, e->n_in, e->nv, t.Lring, e->ring_len, t.S, t.s_log, t.m, e->B, r, t.m, th, tsel};
        if (profile) CA_CUDA(cudaEventRecord(e->tev[j][0], st));
        const bool pdl = e->pdl && !profile;
        launch_tier_fwd(e, t, pdl, fa, count, st);
        if (profile) CA_CUDA(cudaEventRecord(e->tev[j][1], st));
        MacArgs ma = mac_args(e, t, 0u);
        ma.inst0 = r; ma.inst_stride = t.m; ma.tend_host = th; ma.t_sel = tsel;
        launch_mac(t, ma, count, st, pdl);
        if (profile) CA_CUDA(cudaEventRecord(e->tev[j][2], st));
        if (fork) CA_CUDA(cudaEventRecord(e->join_ev[j], st));
        else {
            TierInvArgs ia{t.Ypart, e->d_acc, e->d_ctl, t.tw, t.tw + t.S, t.n_split, e->n_out, t.S, t.s_log, e->B, t.off, e->acc_len, r, t.m, th, tsel};
            launch_tier_inv(e, t, pdl, ia, count, st);
            if (profile) CA_CUDA(cudaEventRecord(e->tev[j][3], st));
        }
    }
    if (fork)
        for (size_t j = 1; j < e->tiers.size(); j++) {
            const Tier &t = e->tiers[j];
            const uint32_t count = tier_count(e, t, tend), r = tier_residue(t, tend);
            if (!count) continue;
            CA_CUDA(cudaStreamWaitEvent(base, e->join_ev[j], 0));
            TierInvArgs ia{t.Ypart, e->d_acc, e->d_ctl, t.tw, t.tw + t.S, t.n_split, e->n_out, t.S, t.s_log, e->B, t.off, e->acc_len, r, t.m, th, tsel};
            launch_tier_inv(e, t, e->pdl, ia, count, base);
        }
    CA_CUDA(cudaGetLastError());
    return CA_OK;
}

uint32_t tiers_firing(const ca_engine *e, uint64_t tend)
{
    uint32_t n = 0;
    for (size_t j = 1; j < e->tiers.size(); j++) n += tier_count(e, e->tiers[j], tend) ? 1u : 0u;
    return n;
}

// graph key of the deferred launch pattern at t_end: which tiers fire and for which residue class
// (key 0 is the period pipeline)
uint64_t tiers_key(const ca_engine *e, uint64_t tend)
{
    uint64_t key = 1;
    for (size_t j = 1; j < e->tiers.size(); j++) {
        const Tier &t = e->tiers[j];
        key = key * (t.m + 1) + (tier_count(e, t, tend) ? tier_residue(t, tend) + 1 : 0);
    }
    return key;
}

template <class F>
int capture_graph(ca_engine *e, cudaGraphExec_t *out, F body)
{
    cudaGraph_t g = nullptr;
    CA_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    const cudaError_t erc = cudaStreamEndCapture(e->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    CA_CUDA(erc);
    CA_CUDA(cudaGraphInstantiate(out, g, 0));
    cudaGraphDestroy(g);
    return CA_OK;
}

void drop_graphs(ca_engine *e)
{
    for (auto &kv : e->graphs) cudaGraphExecDestroy(kv.second);
    e->graphs.clear();
}

int pipe_drain(ca_engine *e);

// phase 1 of a period: everything the output block depends on
int run_period(ca_engine *e, const float *d_in, float *d_out)
{
    int rc = pipe_drain(e);
    if (rc) return rc;
    if (e->async_used) {
        // asynchronous tiers: this period (t_end = t_host + 1) needs the tiers that fired at t_end - 2 (their
        // results start in this period's output block, and their state buffer is the one fwd0 rewrites now);
        // the tiers of t_end - 1 keep running beside this period
        if (drain_params(e)) CA_CUDA(cudaStreamWaitEvent(e->stream, e->ev_def[e->t_host & 1], 0));  // they read the parameters
        CA_CUDA(cudaStreamWaitEvent(e->stream, e->ev_def[(e->t_host + 1) & 1], 0));
    }
    rc = flush_params(e);
    if (rc) return rc;
    const bool profile = (e->cfg.flags & CA_FLAG_PROFILE) != 0;
    if ((e->cfg.flags & CA_FLAG_GRAPH) && !profile) {
        if (e->g_in != d_in || e->g_out != d_out || e->g_active != e->n_active) {
            drop_graphs(e);
            e->g_in = d_in; e->g_out = d_out; e->g_active = e->n_active;
        }
        auto it = e->graphs.find((uint64_t)0);
        if (it == e->graphs.end()) {
            cudaGraphExec_t ge = nullptr;
            rc = capture_graph(e, &ge, [&] { return launch_period(e, d_in, d_out, false, 0, e->n_active, true); });
            if (rc) return rc;
            it = e->graphs.emplace((uint64_t)0, ge).first;
        }
        CA_CUDA(cudaGraphLaunch(it->second, e->stream));
    } else {
        rc = launch_period(e, d_in, d_out, profile, 0, e->n_active, true);
        if (rc) return rc;
    }
    e->launches += (e->fused ? 2 : 3) + (e->quirks ? 1 : 0);
    return CA_OK;
}

// phase 2: the long tiers, off the output's critical path
int run_deferred(ca_engine *e)
{
    const uint64_t tend = ++e->t_host;  // == ctl->t once k_inverse of this period has run
    const bool profile = (e->cfg.flags & CA_FLAG_PROFILE) != 0;
    const uint32_t firing = tiers_firing(e, tend);
    int rc = CA_OK;
    cudaStream_t ts = e->stream;
    if (firing && e->async_tiers && !profile) {
        ts = e->s_def;
        CA_CUDA(cudaEventRecord(e->ev_period, e->stream));  // forward transforms (time ring, voice state) of this period
        CA_CUDA(cudaStreamWaitEvent(ts, e->ev_period, 0));
        e->async_used = true;
    }
    if (firing) {
        if ((e->cfg.flags & CA_FLAG_GRAPH) && !profile) {
            const uint64_t key = tiers_key(e, tend);
            auto it = e->graphs.find(key);
            if (it == e->graphs.end()) {
                cudaGraphExec_t ge = nullptr;
                rc = capture_graph(e, &ge, [&] { return launch_tiers(e, tend, false); });
                if (rc) return rc;
                it = e->graphs.emplace(key, ge).first;
            }
            CA_CUDA(cudaGraphLaunch(it->second, ts));
        } else {
            rc = launch_tiers(e, tend, profile, ts);
            if (rc) return rc;
        }
        if (ts != e->stream) CA_CUDA(cudaEventRecord(e->ev_def[tend & 1], ts));
        for (size_t j = 1; j < e->tiers.size(); j++)
            if (tier_count(e, e->tiers[j], tend)) e->launches += tier_launch_count(e->tiers[j]);
    }
    if (profile) {
        CA_CUDA(cudaEventRecord(e->ev[4], e->stream));
        CA_CUDA(cudaEventSynchronize(e->ev[4]));
        for (int i = 0; i < 4; i++) {
            float ms = 0;
            CA_CUDA(cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
            e->prof_us[i] += 1e3 * ms;
        }
        for (size_t j = 1; j < e->tiers.size(); j++)
            if (e->tev_used[j])
                for (int i = 0; i < 3; i++) {
                    float ms = 0;
                    CA_CUDA(cudaEventElapsedTime(&ms, e->tev[j][i], e->tev[j][i + 1]));
                    e->tier_prof_us[i] += 1e3 * ms;
                }
        e->prof_n++;
    }
    return CA_OK;
}

// ------------------------------------------------------------------------------------------
// Pipelined batch schedule.  The period's kernels fall into two classes: the MACs stream HBM and
// barely issue instructions, the FFT kernels (forward, inverse, long-tier transforms) are
// latency/issue-bound and barely touch HBM.  Run back to back they leave each resource idle half of
// the time, so batches run them on two lanes that overlap:
//   lane C (e->stream): fwd0(p) | tier fwd_j(p) | tier inv_j(p-1) | inv0(p)
//   lane M (e->s_mac) : mac0(p) | tier mac_j(p)
// fwd0 -> mac0 -> inv0 and tier fwd_j -> mac_j -> inv_j are ordered by events; the long tiers' inverse
// of period p runs in period p + 1 (its result is first read by inv0(p + 1)), which takes the tier
// MACs off lane C's critical path.  Lane M kernels get the period count from the host (they run
// while inv0 advances ctl->t) and the long tiers' partial sums are double-buffered by period parity.
// Host-driven launches, no graph: a period of a batch is hundreds of microseconds.
// ------------------------------------------------------------------------------------------
void trace_mark(ca_engine *e, const char *name, int lane, bool begin)
{
    if (!e->tracing) return;
    cudaStream_t st = lane == 0 ? e->stream : lane == 1 ? e->s_mac : e->s_tier[lane - 1];
    if (begin) {
        ca_engine::TraceEv t{name, lane, nullptr, nullptr};
        cudaEventCreate(&t.a); cudaEventCreate(&t.b);
        cudaEventRecord(t.a, st);
        e->trace.push_back(t);
    } else {
        cudaEventRecord(e->trace.back().b, st);
    }
}

void trace_dump(ca_engine *e)
{
    cudaStreamSynchronize(e->s_mac); cudaStreamSynchronize(e->stream);
    for (auto &st : e->s_tier) cudaStreamSynchronize(st);
    for (auto &t : e->trace) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e->trace.front().a, t.a);
        cudaEventElapsedTime(&b, e->trace.front().a, t.b);
        fprintf(stderr, "trace lane %c %-10s %8.1f .. %8.1f us (%6.1f)\n", "CM23456"[t.lane], t.name, 1e3 * a, 1e3 * b, 1e3 * (b - a));
    }
    for (auto &t : e->trace) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    e->trace.clear();
}

bool use_pipeline(const ca_engine *e)
{
    if (e->tiers.size() < 2 || e->fused || !e->s_mac || e->async_tiers || (e->cfg.flags & (CA_FLAG_PROFILE | CA_FLAG_GRAPH))) return false;
    return e->pipe_mode == 1;
}

// long-tier lanes: inverse transforms of the tiers that fired at `pipe_prev_tend`, in tier order (they
// accumulate into the same output ring).  e->stream joins them now (drain) or right before inv0.
int pipe_join_tinv(ca_engine *e)
{
    for (size_t j = 1; j < e->tiers.size(); j++)
        if (e->tinv_pending[j]) {
            CA_CUDA(cudaStreamWaitEvent(e->stream, e->ptinv_ev[j], 0));
            e->tinv_pending[j] = false;
        }
    return CA_OK;
}

int pipe_finish_prev(ca_engine *e, bool join_now)
{
    const uint64_t tend = e->pipe_prev_tend;
    e->pipe_prev_tend = 0;
    size_t prev = 0;
    for (size_t j = 1; tend && j < e->tiers.size(); j++) {
        const Tier &t = e->tiers[j];
        const uint32_t count = tier_count(e, t, tend), r = tier_residue(t, tend);
        if (!count) continue;
        cudaStream_t st = e->s_tier[j];
        CA_CUDA(cudaStreamWaitEvent(st, e->ptm_ev[j][tend & 1], 0));
        if (prev) CA_CUDA(cudaStreamWaitEvent(st, e->ptinv_ev[prev], 0));
        TierInvArgs ia{(tend & 1) ? t.Ypart2 : t.Ypart, e->d_acc, e->d_ctl, t.tw, t.tw + t.S, t.n_split, e->n_out, t.S, t.s_log, e->B, t.off, e->acc_len, r, t.m, tend};
        trace_mark(e, j == 1 ? "tinv1'" : "tinv2+'", (int)(1 + j), true);
        e->launches += launch_tier_inv(e, t, false, ia, count, st);
        trace_mark(e, "", (int)(1 + j), false);
        CA_CUDA(cudaEventRecord(e->ptinv_ev[j], st));
        e->tinv_pending[j] = true;
        prev = j;
    }
    CA_CUDA(cudaGetLastError());
    return join_now ? pipe_join_tinv(e) : CA_OK;
}

// everything in flight on lane M becomes a dependency of e->stream (callers then order against / sync e->stream only)
int pipe_drain(ca_engine *e)
{
    if (!e->pipe_used) return CA_OK;
    const int rc = pipe_finish_prev(e, true);
    if (rc) return rc;
    for (size_t j = 1; j < e->tiers.size(); j++) CA_CUDA(cudaStreamWaitEvent(e->stream, e->ptf_ev[j], 0));  // tier lanes: last forward
    CA_CUDA(cudaStreamWaitEvent(e->stream, e->pm_tail, 0));
    e->pipe_used = false;
    return CA_OK;
}

// quiescence for callers that touch shared state (IR bank, active count, sync): the pipelined lanes and
// the asynchronous tier stream become dependencies of e->stream
int drain_all(ca_engine *e)
{
    if (e->async_used) {
        CA_CUDA(cudaStreamWaitEvent(e->stream, e->ev_def[0], 0));
        CA_CUDA(cudaStreamWaitEvent(e->stream, e->ev_def[1], 0));
    }
    return pipe_drain(e);
}

// one period of a batch; h_src / h_dst: pinned host buffers (ca_process) or nullptr (device-resident)
int run_pipelined(ca_engine *e, const float *d_in, float *d_out, uint32_t chunks, const float *h_src, float *h_dst)
{
    if (drain_params(e)) {  // parameters change: lane M still reads the old ones
        const int rc = pipe_drain(e);
        if (rc) return rc;
    }
    int rc = flush_params(e);
    if (rc) return rc;
    const uint64_t tend = e->t_host + 1;
    const Tier &t0 = e->tiers[0];
    const uint32_t n_alloc = e->n_inst * e->n_in;
    const size_t in_stride = (size_t)e->n_in * e->B, out_stride = (size_t)e->n_out * e->B;
    auto cut = [&](uint32_t c) { return (uint32_t)((uint64_t)e->n_active * c / chunks); };
    e->tracing = e->trace_at && (tend == e->trace_at || tend == e->trace_at + 1);
    rc = pipe_finish_prev(e, false);  // tier lanes: the previous period's long-tier results (needed by this period's inv0)
    if (rc) return rc;
    // the previous period's tier forwards read the time ring this period's fwd0 writes (ahead of them, but
    // within ring_len of a maximal predelay): order them, they finished long ago
    for (size_t j = 1; j < e->tiers.size(); j++) CA_CUDA(cudaStreamWaitEvent(e->stream, e->ptf_ev[j], 0));
    if (h_src)
        for (uint32_t c = 0; c < chunks; c++) {
            const uint32_t i0 = cut(c), i1 = cut(c + 1);
            CA_CUDA(cudaMemcpyAsync(const_cast<float *>(d_in) + i0 * in_stride, h_src + i0 * in_stride, (i1 - i0) * in_stride * sizeof(float), cudaMemcpyHostToDevice, e->s_in));
            CA_CUDA(cudaEventRecord(e->io_ev[0][c], e->s_in));
        }
    for (uint32_t c = 0; c < chunks; c++) {
        const uint32_t i0 = cut(c), i1 = cut(c + 1);
        const bool last = c + 1 == chunks;
        if (h_src) CA_CUDA(cudaStreamWaitEvent(e->stream, e->io_ev[0][c], 0));
        // lane C: forward transforms of the chunk
        const uint32_t n_items = (i1 - i0) * e->n_in;
        FwdArgs fa{d_in, e->d_ring, t0.X, e->d_par, e->d_st, e->d_ctl, t0.tw, t0.tw + e->B,
                   n_items, n_alloc, e->n_in, e->nv, t0.Lring, e->ring_len, e->ring_out, i0 * e->n_in};
        trace_mark(e, "fwd0", 0, true);
        launch_fwd0(e, false, fa, e->stream);
        trace_mark(e, "", 0, false);
        CA_CUDA(cudaEventRecord(e->pf_ev[c], e->stream));
        // lane M: tier-0 MAC of the chunk
        CA_CUDA(cudaStreamWaitEvent(e->s_mac, e->pf_ev[c], 0));
        MacArgs ma = mac_args(e, t0, 1u);
        ma.inst0 = i0; ma.tend_host = tend;
        trace_mark(e, "mac0", 1, true);
        launch_mac(t0, ma, i1 - i0, e->s_mac);
        trace_mark(e, "", 1, false);
        CA_CUDA(cudaEventRecord(e->pm_ev[c], e->s_mac));
        e->launches += 2;
        if (last)  // every chunk's input is in the time ring: long tiers whose block closes with this period
            for (size_t j = 1; j < e->tiers.size(); j++) {
                const Tier &t = e->tiers[j];
                const uint32_t count = tier_count(e, t, tend), r = tier_residue(t, tend);
                if (!count) continue;
                TierFwdArgs tf{e->d_ring, t.X, e->d_st, e->d_ctl, t.tw, t.tw + t.S, n_alloc, e->n_in, e->nv, t.Lring, e->ring_len, t.S, t.s_log, t.m, e->B, r, t.m, tend};
                CA_CUDA(cudaStreamWaitEvent(e->s_tier[j], e->pf_ev[c], 0));
                trace_mark(e, j == 1 ? "tfwd1" : "tfwd2+", (int)(1 + j), true);
                e->launches += launch_tier_fwd(e, t, false, tf, count, e->s_tier[j]) - 1;
                trace_mark(e, "", (int)(1 + j), false);
                CA_CUDA(cudaEventRecord(e->ptf_ev[j], e->s_tier[j]));
                CA_CUDA(cudaStreamWaitEvent(e->s_mac, e->ptf_ev[j], 0));
                MacArgs tm = mac_args(e, t, 0u);
                tm.inst0 = r; tm.inst_stride = t.m; tm.tend_host = tend;
                if (tend & 1) tm.Ypart = t.Ypart2;
                trace_mark(e, j == 1 ? "tmac1" : "tmac2+", 1, true);
                launch_mac(t, tm, count, e->s_mac);
                trace_mark(e, "", 1, false);
                CA_CUDA(cudaEventRecord(e->ptm_ev[j][tend & 1], e->s_mac));
                e->launches += 2;
            }
        if (c == 0) {  // the previous period's long-tier results land in the output ring before any inv0 reads it
            rc = pipe_join_tinv(e);
            if (rc) return rc;
        }
        // lane C: inverse transform + mix of the chunk
        CA_CUDA(cudaStreamWaitEvent(e->stream, e->pm_ev[c], 0));
        InvArgs ia{t0.Ypart, d_in, d_out, e->d_acc, e->d_par, e->d_ctl, t0.tw, t0.tw + e->B,
                   t0.n_split, e->n_in, e->n_out, e->acc_len, (i1 - i0) * e->n_out, i0 * e->n_out, last ? 1u : 0u, (e->cfg.flags & CA_FLAG_RAW_WET) ? 1u : 0u};
        trace_mark(e, "inv0", 0, true);
        launch_inv0(e, false, ia, e->stream);
        trace_mark(e, "", 0, false);
        e->launches += 1;
        if (h_dst) {
            CA_CUDA(cudaEventRecord(e->io_ev[1][c], e->stream));
            CA_CUDA(cudaStreamWaitEvent(e->s_out, e->io_ev[1][c], 0));
            CA_CUDA(cudaMemcpyAsync(h_dst + i0 * out_stride, d_out + i0 * out_stride, (i1 - i0) * out_stride * sizeof(float), cudaMemcpyDeviceToHost, e->s_out));
        }
    }
    CA_CUDA(cudaEventRecord(e->pm_tail, e->s_mac));
    if (h_dst) CA_CUDA(cudaEventRecord(e->out_ready, e->s_out));
    CA_CUDA(cudaGetLastError());
    e->pipe_used = true;
    e->pipe_prev_tend = tiers_firing(e, tend) ? tend : 0;
    e->t_host = tend;
    if (e->tracing && tend == e->trace_at + 1) trace_dump(e);
    e->tracing = false;
    return CA_OK;
}

// Capture + instantiate every graph ca_process() will need (period pipeline over the engine's own
// staging buffers, and one graph per deferred-tier launch pattern) so that no capture or
// instantiation ever happens inside a real-time period.
int prewarm_graphs(ca_engine *e)
{
    if (!(e->cfg.flags & CA_FLAG_GRAPH) || (e->cfg.flags & CA_FLAG_PROFILE)) return CA_OK;
    drop_graphs(e);
    e->g_in = e->d_in; e->g_out = e->d_out; e->g_active = e->n_active;
    cudaGraphExec_t ge = nullptr;
    int rc = capture_graph(e, &ge, [&] { return launch_period(e, e->d_in, e->d_out, false, 0, e->n_active, true); });
    if (rc) return rc;
    e->graphs.emplace((uint64_t)0, ge);
    const uint64_t cycle = e->tiers.back().m;
    for (uint64_t tend = 1; tend <= cycle && e->tiers.size() > 1; tend++) {
        if (!tiers_firing(e, tend)) continue;
        const uint64_t key = tiers_key(e, tend);
        if (e->graphs.count(key)) continue;
        ge = nullptr;
        rc = capture_graph(e, &ge, [&] { return launch_tiers(e, tend, false); });
        if (rc) return rc;
        e->graphs.emplace(key, ge);
    }
    for (auto &kv : e->graphs) cudaGraphUpload(kv.second, e->stream);
    CA_CUDA(cudaStreamSynchronize(e->stream));
    return CA_OK;
}

// Split the device's SMs into two green contexts: `n_mac` SMs (rounded to the hardware's granularity) for the
// memory-bound MAC lane, the rest for the latency-bound FFT lanes of the pipelined batch schedule.  Two kernels
// sharing an SM starve each other (the persistent MAC holds every register); on disjoint SMs both run at their
// own speed and the MAC still saturates HBM from a subset of the SMs.
template <class F>
bool drv(const char *name, F *fn)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) { (void)cudaGetLastError(); return false; }
    *fn = reinterpret_cast<F>(p);
    return true;
}

int setup_sm_split(ca_engine *e, uint32_t n_mac, int prio_lo, int prio_hi)
{
    decltype(&cuDeviceGet) pDeviceGet; decltype(&cuDeviceGetDevResource) pGetRes; decltype(&cuDevSmResourceSplitByCount) pSplit;
    decltype(&cuDevResourceGenerateDesc) pDesc; decltype(&cuGreenCtxCreate) pCreate; decltype(&cuGreenCtxStreamCreate) pStream;
    if (!drv("cuDeviceGet", &pDeviceGet) || !drv("cuDeviceGetDevResource", &pGetRes) || !drv("cuDevSmResourceSplitByCount", &pSplit) ||
        !drv("cuDevResourceGenerateDesc", &pDesc) || !drv("cuGreenCtxCreate", &pCreate) || !drv("cuGreenCtxStreamCreate", &pStream)) {
        g_last_error = "green contexts are not available in this driver";
        return CA_ERR_UNSUPPORTED;
    }
    CA_CUDA(cudaFree(nullptr));  // primary context
    CUdevice dev;
    CUdevResource all, grp, rem;
    unsigned int nb = 1;
    CUdevResourceDesc d1, d2;
    auto ok = [](CUresult r, const char *what) { if (r != CUDA_SUCCESS) { g_last_error = std::string(what) + " failed: CUresult " + std::to_string((int)r); return false; } return true; };
    if (!ok(pDeviceGet(&dev, e->device), "cuDeviceGet") || !ok(pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM), "cuDeviceGetDevResource") ||
        !ok(pSplit(&grp, &nb, &all, &rem, 0, n_mac), "cuDevSmResourceSplitByCount") || nb != 1 || rem.sm.smCount == 0 ||
        !ok(pDesc(&d1, &grp, 1), "cuDevResourceGenerateDesc") || !ok(pDesc(&d2, &rem, 1), "cuDevResourceGenerateDesc") ||
        !ok(pCreate(&e->g_mac, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate") || !ok(pCreate(&e->g_fft, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate"))
        return CA_ERR_UNSUPPORTED;
    e->sm_mac = grp.sm.smCount; e->sm_fft = rem.sm.smCount;
    CUstream st;
    if (!ok(pStream(&st, e->g_fft, CU_STREAM_NON_BLOCKING, prio_hi), "cuGreenCtxStreamCreate")) return CA_ERR_UNSUPPORTED;
    e->stream = st;
    if (!ok(pStream(&st, e->g_mac, CU_STREAM_NON_BLOCKING, prio_lo), "cuGreenCtxStreamCreate")) return CA_ERR_UNSUPPORTED;
    e->s_mac = st;
    for (auto &ts : e->s_tier) { if (!ok(pStream(&st, e->g_fft, CU_STREAM_NON_BLOCKING, prio_hi), "cuGreenCtxStreamCreate")) return CA_ERR_UNSUPPORTED; ts = st; }
    return CA_OK;
}

// ---- CA_FLAG_PERSISTENT: the resident kernel and its mailbox ----
typedef void (*persist_fn)(const PersistArgs);
persist_fn persist_pick(uint32_t R, uint32_t n_out)
{
    switch (R) {
    case 1: return n_out == 1 ? k_persist<1, 1> : k_persist<1, 2>;
    case 2: return n_out == 1 ? k_persist<2, 1> : k_persist<2, 2>;
    case 4: return n_out == 1 ? k_persist<4, 1> : k_persist<4, 2>;
    case 8: return n_out == 1 ? k_persist<8, 1> : k_persist<8, 2>;
    default: return nullptr;
    }
}

int persist_stop(ca_engine *e)
{
    if (!e->persistent || !e->p_running) return CA_OK;
    e->pbox->seq_in = kPersistExit;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    CA_CUDA(cudaStreamSynchronize(e->stream));
    e->p_running = false;
    e->pbox->seq_in = e->t_host;
    return CA_OK;
}

int persist_launch(ca_engine *e)
{
    const Tier &t0 = e->tiers[0];
    e->p_gen++;
    e->pbox->exited = 0;
    e->pbox->seq_in = e->t_host;       // nothing to do yet
    e->pbox->seq_out = e->t_host;
    e->pbox->pad[0] = (unsigned int)(e->t_host & 0xffffffffu);
    e->pbox->pad[1] = (unsigned int)(e->t_host >> 32);
    CA_CUDA(cudaMemsetAsync(e->d_arrive, 0, 2 * sizeof(unsigned int), e->stream));
    CA_CUDA(cudaMemcpyAsync(e->d_go, e->pbox->pad, sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));  // pinned source
    PersistArgs pa{};
    pa.box = e->pbox; pa.ring = e->d_ring; pa.X = t0.X; pa.H = t0.H; pa.Ypart = e->d_pY; pa.st = e->d_st; pa.par_dev = e->d_ppar;
    pa.twM = t0.tw; pa.tw2M = t0.tw + e->B; pa.go = e->d_go; pa.arrive = e->d_arrive; pa.t0 = e->t_host;
    pa.n_in = e->n_in; pa.n_out = e->n_out; pa.nv = e->nv; pa.Lring = t0.Lring; pa.P = t0.P; pa.ring_len = e->ring_len; pa.ring_out = e->ring_out;
    pa.k_off = e->k_off; pa.raw_wet = (e->cfg.flags & CA_FLAG_RAW_WET) ? 1u : 0u; pa.gen = e->p_gen; pa.n_items_alloc = e->n_inst * e->n_in;
    pa.vp = voice_pool(e);
    pa.Ysum = e->d_pYsum; pa.stamp = e->p_stamp ? 1u : 0u;
    void *kargs[] = {&pa};
    CA_CUDA(cudaLaunchCooperativeKernel((const void *)persist_pick(e->R, e->n_out), dim3(e->p_ctas), dim3(kPersistThreads), kargs, 0, e->stream));
    e->p_running = true;
    e->launches += 1;
    return CA_OK;
}

// one period through the mailbox; in / out: host memory of any kind
int process_persistent(ca_engine *e, const float *in, float *out)
{
    PersistBox *b = e->pbox;
    drain_params(e);
    e->par_dirty = false;
    memcpy(b->par, e->par.data(), e->n_in * sizeof(InParamDev));
    memcpy(b->in, in, (size_t)e->n_in * e->B * sizeof(float));
    if (e->p_running && b->exited == e->p_gen) {  // the kernel left (idle): collect it
        CA_CUDA(cudaStreamSynchronize(e->stream));
        e->p_running = false;
    }
    if (!e->p_running) {
        const int rc = persist_launch(e);
        if (rc) return rc;
    }
    const unsigned long long want = e->t_host + 1ull;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    b->seq_in = want;
    const double t0 = now_us();
    for (unsigned spins = 0; b->seq_out < want; spins++) {
        if ((spins & 0xfff) == 0xfff) {
            if (b->exited == e->p_gen) {  // it left between our check and the signal: relaunch, the request stands
                CA_CUDA(cudaStreamSynchronize(e->stream));
                e->p_running = false;
                const int rc = persist_launch(e);
                if (rc) return rc;
                b->seq_in = want;
            }
            if (now_us() - t0 > 4e6) { g_last_error = "persistent kernel does not answer"; return CA_ERR_CUDA; }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(out, b->out, (size_t)e->n_out * e->B * sizeof(float));
    if (b->err) { g_last_error = "persistent kernel: grid barrier timed out"; return CA_ERR_CUDA; }
    e->t_host = want;
    return CA_OK;
}

void record_wall(ca_engine *e, double us)
{
    e->wall[e->periods % e->wall.size()] = (float)us;
    e->periods++;
    e->wall_sum += us;
    if (us > e->wall_max) e->wall_max = us;
    if (e->deadline_us > 0 && us > e->deadline_us) e->xruns++;
}

}  // namespace

extern "C" {

int ca_api_version(void) { return CA_API_VERSION; }

const char *ca_strerror(int code)
{
    switch (code) {
    case CA_OK: return "ok";
    case CA_ERR_INVALID: return "invalid argument";
    case CA_ERR_CUDA: return "CUDA failure";
    case CA_ERR_NOMEM: return "out of memory";
    case CA_ERR_STATE: return "invalid state (IR slot not loaded?)";
    case CA_ERR_UNSUPPORTED: return "unsupported configuration";
    case CA_ERR_PERIOD: return "nframes does not match the configured period";
    default: return "unknown error";
    }
}

const char *ca_last_error_string(void) { return g_last_error.c_str(); }

void ca_config_init(ca_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->period = 256;
    cfg->n_instances = 1;
    cfg->n_in = cfg->n_out = 2;
    cfg->max_ir_frames = 512 * 256 - 1024;  // CONV_DEFAULT_FFTSIZE - default nframes, conv.h:10-12,63
    cfg->n_ir_slots = 2;
    cfg->sample_rate = 48000.f;
}

// Non-uniform partitioning a la Gardner with one period of slack per tier: tier j+1's block is
// `growth` times tier j's, and tier j carries just enough partitions for tier j+1 to start at an
// offset >= its own block size.
int ca_config_auto_tiers(ca_config *cfg, uint32_t growth, uint32_t max_block)
{
    if (!cfg || !is_pow2(cfg->period) || !cfg->max_ir_frames) return CA_ERR_INVALID;
    if (!max_block) max_block = 16384;
    // growth 8: fewest tiers, least FFT work.  For batches at period 256 growth 4 (256 x 4 | 1024 x 3 | 4096 x 3 | 16384)
    // streams 258 KB per instance-period instead of 319 KB and is 4 % faster device-resident (1 077 -> 1 034 us at
    // 16 128 instances) but no faster through ca_process (1 158 vs 1 157 us) and its short work items stream at 0.93 of
    // the HBM peak instead of 0.99: not the default.
    if (!growth) growth = 8;
    if (!is_pow2(growth) || growth < 2 || !is_pow2(max_block)) return CA_ERR_INVALID;
    uint32_t n = 0, S = cfg->period, off = 0;
    memset(cfg->tier_block, 0, sizeof(cfg->tier_block));
    memset(cfg->tier_parts, 0, sizeof(cfg->tier_parts));
    while (n < CA_MAX_TIERS) {
        uint32_t next = S * growth;
        if (next < 256) next = 256;  // long tiers use the CTA-level FFT: block >= 256
        if (next > max_block && S < max_block) next = max_block;
        const bool last = (n + 1 == CA_MAX_TIERS) || next > max_block || off + next + ((cfg->flags & CA_FLAG_ASYNC_TIERS) ? cfg->period : 0u) >= cfg->max_ir_frames;
        cfg->tier_block[n] = S;
        if (last) { cfg->tier_parts[n] = 0; n++; break; }
        // reach offset >= next (the tier's result is due the period after its block closes), one period
        // more with CA_FLAG_ASYNC_TIERS (due two periods later: the tier runs beside the next period)
        const uint32_t slack = (cfg->flags & CA_FLAG_ASYNC_TIERS) ? cfg->period : 0u;
        const uint32_t parts = (next + slack - off + S - 1) / S;
        cfg->tier_parts[n] = parts;
        off += parts * S;
        S = next;
        n++;
    }
    cfg->n_tiers = n;
    return CA_OK;
}

int ca_destroy(ca_engine *e)
{
    if (!e) return CA_OK;
    cudaSetDevice(e->device);
    persist_stop(e);
    if (e->s_mac) cudaStreamSynchronize(e->s_mac);
    if (e->s_def) cudaStreamSynchronize(e->s_def);
    for (auto &st : e->s_tier) if (st) cudaStreamSynchronize(st);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->s_out) cudaStreamSynchronize(e->s_out);
    drop_graphs(e);
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto &row : e->tev) for (auto &ev : row) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->upload_done) if (ev) cudaEventDestroy(ev);
    if (e->out_ready) cudaEventDestroy(e->out_ready);
    for (auto &row : e->io_ev) for (auto &ev : row) if (ev) cudaEventDestroy(ev);
    for (auto &st : e->s_tier) if (st) cudaStreamDestroy(st);
    if (e->fork_ev) cudaEventDestroy(e->fork_ev);
    for (auto &ev : e->join_ev) if (ev) cudaEventDestroy(ev);
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    if (e->s_mac) cudaStreamDestroy(e->s_mac);
    if (e->s_def) cudaStreamDestroy(e->s_def);
    if (e->ev_period) cudaEventDestroy(e->ev_period);
    for (auto &ev : e->ev_def) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->pf_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->pm_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->ptf_ev) if (ev) cudaEventDestroy(ev);
    for (auto &row : e->ptm_ev) for (auto &ev : row) if (ev) cudaEventDestroy(ev);
    if (e->pm_tail) cudaEventDestroy(e->pm_tail);
    for (auto &ev : e->ptinv_ev) if (ev) cudaEventDestroy(ev);
    for (auto &t : e->tiers) { cudaFree(t.Ypart); cudaFree(t.Ypart2); cudaFree(t.tw); cudaFree(t.workctr); }
    cudaFree(e->d_arena); cudaFree(e->d_ring); cudaFree(e->d_acc);
    cudaFree(e->d_in); cudaFree(e->d_out); cudaFree(e->d_par); cudaFree(e->d_st); cudaFree(e->d_ctl); cudaFree(e->d_rowtw); cudaFree(e->d_vpool);
    cudaFree(e->d_irsum); cudaFree(e->d_qdelta); cudaFree(e->d_qrun); cudaFree(e->d_qring);
    cudaFree(e->d_go); cudaFree(e->d_arrive); cudaFree(e->d_ppar); cudaFree(e->d_pY); cudaFree(e->d_pYsum);
    if (e->pbox) cudaFreeHost(e->pbox);
    cudaFreeHost(e->h_in); cudaFreeHost(e->h_out); cudaFreeHost(e->h_upload[0]); cudaFreeHost(e->h_upload[1]);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->g_mac || e->g_fft) {
        decltype(&cuGreenCtxDestroy) pDestroy = nullptr;
        if (drv("cuGreenCtxDestroy", &pDestroy)) {
            if (e->g_mac) pDestroy(e->g_mac);
            if (e->g_fft) pDestroy(e->g_fft);
        }
    }
    delete e;
    return CA_OK;
}

static int plan_tiers(const ca_config *cfg, ca_engine *e)
{
    const uint32_t B = cfg->period;
    const uint32_t nt = cfg->n_tiers > 1 ? cfg->n_tiers : 1;
    if (nt > CA_MAX_TIERS) { g_last_error = "too many tiers"; return CA_ERR_INVALID; }
    e->tiers.assign(nt, Tier{});
    if (nt == 1) {
        Tier &t = e->tiers[0];
        const uint32_t P_total = (cfg->max_ir_frames + B - 1) / B;
        t.S = B; t.m = 1; t.off = 0;
        if (cfg->part_count) {
            if (cfg->part_begin + cfg->part_count > P_total) { g_last_error = "partition shard exceeds the IR"; return CA_ERR_INVALID; }
            t.P = cfg->part_count; e->k_off = cfg->part_begin;
        } else { t.P = P_total; e->k_off = 0; }
        t.Lring = e->k_off + t.P;
        return CA_OK;
    }
    if (cfg->part_count) { g_last_error = "partition-range shards and tiers cannot be combined"; return CA_ERR_UNSUPPORTED; }
    uint32_t off = 0;
    for (uint32_t j = 0; j < nt; j++) {
        Tier &t = e->tiers[j];
        t.S = cfg->tier_block[j];
        if (j == 0 && t.S != B) { g_last_error = "tier 0 block must equal the period"; return CA_ERR_INVALID; }
        if (!is_pow2(t.S) || (j > 0 && (t.S <= e->tiers[j - 1].S || t.S < 256 || t.S > 16384))) { g_last_error = "tier blocks must be increasing powers of two, 256..16384 above tier 0"; return CA_ERR_INVALID; }
        if (j > 0 && off < t.S) { g_last_error = "tier offset must be >= its block size (previous tiers too short)"; return CA_ERR_INVALID; }
        if (j > 0 && (cfg->flags & CA_FLAG_ASYNC_TIERS) && off < t.S + B) { g_last_error = "CA_FLAG_ASYNC_TIERS: tier offset must be >= block + period"; return CA_ERR_INVALID; }
        if (off >= cfg->max_ir_frames) { e->tiers.resize(j); break; }
        t.m = t.S / B; t.off = off;
        const uint32_t rest = (cfg->max_ir_frames - off + t.S - 1) / t.S;
        t.P = (j + 1 == nt || !cfg->tier_parts[j]) ? rest : std::min(cfg->tier_parts[j], rest);
        t.Lring = t.P;
        off += t.P * t.S;
    }
    if (off < cfg->max_ir_frames) { g_last_error = "tiers do not cover max_ir_frames"; return CA_ERR_INVALID; }
    return CA_OK;
}

static int create_impl(const ca_config *cfg, ca_engine *e)
{
    e->cfg = *cfg;
    e->device = cfg->device;
    CA_CUDA(cudaSetDevice(cfg->device));
    e->B = cfg->period; e->R = cfg->period / 32;
    e->n_inst = e->n_active = cfg->n_instances; e->n_in = cfg->n_in; e->n_out = cfg->n_out;
    e->nv = cfg->max_voices ? cfg->max_voices : 2u;
    if (cfg->io_chunks) e->io_chunks = std::min<uint32_t>(kIoChunks, cfg->io_chunks);
    if (const char *c = getenv("CA_IO_CHUNKS")) e->io_chunks = std::max(1, std::min<int>(kIoChunks, atoi(c)));
    int rc = plan_tiers(cfg, e);
    if (rc) return rc;
    if (cfg->flags & CA_FLAG_REF_QUIRKS) {
        const uint32_t N = cfg->ref_fft_size;
        if (e->n_in != 2 || e->n_out != 2 || !N || N % e->B || N < 2 * e->B || cfg->part_begin || cfg->part_count ||
            (cfg->flags & (CA_FLAG_RAW_WET | CA_FLAG_PERSISTENT)) || (cfg->schedule & CA_SCHED_PIPELINED) || cfg->sm_split) {
            g_last_error = "CA_FLAG_REF_QUIRKS: true stereo (2 in, 2 out), ref_fft_size a multiple of the period, whole IR, no raw-wet / persistent / pipelined schedule";
            return CA_ERR_UNSUPPORTED;
        }
        e->quirks = true;
    }
    e->fft = fft_pick((int)e->R);
    int variant = -1;  // -1: per tier by rows per CTA (measured: short row lists want many small CTAs per SM)
    if (const char *v = getenv("CA_MAC_VARIANT")) variant = atoi(v);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);

    // FFT lanes above the MAC lane: when a (memory-bound) MAC CTA retires, a waiting FFT CTA gets its slot
    int prio_lo = 0, prio_hi = 0;
    CA_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (getenv("CA_PIPE_PRIO") && atoi(getenv("CA_PIPE_PRIO")) == 0) prio_hi = prio_lo;
    const uint32_t sm_split = getenv("CA_SM_SPLIT") ? (uint32_t)atoi(getenv("CA_SM_SPLIT")) : cfg->sm_split;
    if (sm_split) {
        rc = setup_sm_split(e, sm_split, prio_lo, prio_hi);
        if (rc) return rc;
    } else
    CA_CUDA(cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_hi));
    for (auto &ev : e->ev) CA_CUDA(cudaEventCreate(&ev));
    for (auto &row : e->tev) for (auto &ev : row) CA_CUDA(cudaEventCreate(&ev));
    for (auto &ev : e->upload_done) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CA_CUDA(cudaEventCreateWithFlags(&e->out_ready, cudaEventDisableTiming));
    CA_CUDA(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    CA_CUDA(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    if (!sm_split) for (auto &st : e->s_tier) CA_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_hi));
    CA_CUDA(cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming));
    for (auto &ev : e->join_ev) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &row : e->io_ev) for (auto &ev : row) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (!sm_split) CA_CUDA(cudaStreamCreateWithPriority(&e->s_mac, cudaStreamNonBlocking, prio_lo));
    for (auto &ev : e->pf_ev) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : e->pm_ev) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : e->ptf_ev) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &row : e->ptm_ev) for (auto &ev : row) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CA_CUDA(cudaEventCreateWithFlags(&e->pm_tail, cudaEventDisableTiming));
    for (auto &ev : e->ptinv_ev) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->pipe_mode = ((cfg->schedule & CA_SCHED_PIPELINED) || sm_split) ? 1 : 0;
    if (const char *pm = getenv("CA_PIPELINE")) e->pipe_mode = atoi(pm) ? 1 : 0;
    if (e->quirks) e->pipe_mode = 0;
    e->async_tiers = (cfg->flags & CA_FLAG_ASYNC_TIERS) && e->tiers.size() > 1 && !(cfg->flags & CA_FLAG_PROFILE);
    CA_CUDA(cudaStreamCreateWithPriority(&e->s_def, cudaStreamNonBlocking, prio_lo));
    CA_CUDA(cudaEventCreateWithFlags(&e->ev_period, cudaEventDisableTiming));
    for (auto &ev : e->ev_def) CA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->pdl = !(cfg->flags & CA_FLAG_GRAPH) && !(cfg->schedule & CA_SCHED_NO_PDL);
    if (const char *pd = getenv("CA_PDL")) e->pdl = atoi(pd) != 0;
    if (const char *tr = getenv("CA_PIPE_TRACE")) e->trace_at = (uint64_t)atoll(tr);

    const size_t n_items = (size_t)e->n_inst * e->n_in;
    // voice pool: one home entry per (instance, input) + shared entries for the second (third ...) voice of
    // inputs that are cross-fading.  Small engines get every voice resident (n_items * (nv - 1)); batches share
    // n_items / 8 (at least 64), or what ca_config.voice_pool asks for.
    {
        const uint64_t all = (uint64_t)n_items * (e->nv - 1u);
        uint64_t extra = cfg->voice_pool ? cfg->voice_pool : std::max<uint64_t>(64, n_items / 8);
        e->n_extra = (uint32_t)std::min<uint64_t>(all, extra);
        e->n_voices = (uint32_t)n_items + e->n_extra;
    }
    uint32_t s_max = e->B, reach = e->B;
    e->arena_bytes = 0;
    const bool legacy_fft = (cfg->flags & CA_FLAG_LEGACY_FFT) != 0;
    e->rows0 = !legacy_fft && e->B == 256;
    e->rows16 = !(cfg->schedule & CA_SCHED_ROWS8) && !(getenv("CA_ROWS16") && getenv("CA_ROWS16")[0] == '0');
    for (size_t j = 0; j < e->tiers.size(); j++) {
        Tier &t = e->tiers[j];
        t.s_log = t.S >= 256 ? ilog2(t.S / 256) : 0;
        t.M1 = t.S / 256;
        t.rows_mode = (j == 0 || legacy_fft || t.S < 256) ? 0 : (t.M1 <= 16 ? 1 : 2);
        // bin tile of the long tiers: 512 complex (4 KB arrays) halves the CTA count per byte (+6 % measured)
        const uint32_t bt_max = getenv("CA_MAC_BT") ? (uint32_t)atoi(getenv("CA_MAC_BT")) : 512u;
        t.bt = std::min<uint32_t>(t.S, j == 0 ? 256u : bt_max);
        t.tiles = t.S / t.bt;
        // rows per CTA before splitting: long lists (uniform, P in the hundreds) stream best with 96 KB /
        // 2 CTAs per SM, short ones (tiers: 14..22 rows) with 4 CTAs per SM (the register limit) of 3 x 12 KB
        // stages (measured r01, K = 4096: tier-0 MAC 83 us with 96 KB, 78 with 4 x 12 KB, 74.5 with 3 x 12 KB)
        // Short row lists (the tiers of a batch: 6..22 rows per work item) stream best with ONE 48 KB stage per CTA and
        // 4 CTAs per SM: the overlap comes from the other CTAs of the SM, and nothing is paid per stage hand-off.
        // Measured at 16 128 instances, us per period (tier 0 / long tiers; 8-7-11 partitions): 3 x 12 KB ring 275 / 556,
        // 2 x 12 KB 264 / 561, 2 x 24 KB 255 / 550, 1 x 48 KB 249 / 553 (= 6.3 / 6.4 TB/s), 1 x 36 KB 256 / 555, 1 x 72 KB
        // 366 / 555, 1 x 96 KB 312 / 560; one-row stages (6 KB x 2..6) 295-342 / 586.
        const int tier_variant = variant >= 0 ? variant : (t.P * e->n_in <= 128 ? 13 : 1);
        t.mac = mac_pick((int)t.bt, (int)e->n_out, tier_variant);
        CA_CUDA(cudaFuncSetAttribute((const void *)t.mac.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.mac.smem));
        {
            // persistent schedule: resident CTA slots of the device (CA_MAC_PERSIST=0 off, =1 always when
            // n_split == 1, and forces n_split = 1; CA_MAC_SLOTS=n overrides the slot count -- tests force several items per CTA)
            const char *pe = getenv("CA_MAC_PERSIST");
            const bool p_off = pe ? pe[0] == '0' : (cfg->schedule & CA_SCHED_MAC_PER_ITEM) != 0;
            t.p_slots = 0; t.p_force = pe ? pe[0] == '1' : (cfg->schedule & CA_SCHED_MAC_PERSISTENT) != 0;
            if (!p_off) {
                CA_CUDA(cudaFuncSetAttribute((const void *)t.mac.pfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.mac.psmem));
                int per_sm = 0;
                CA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)t.mac.pfn, kMacThreads, t.mac.psmem));
                if (const char *cs = getenv("CA_MAC_CTAS")) per_sm = std::min(per_sm, std::max(1, atoi(cs)));  // resident MAC CTAs per SM
                t.p_slots = (uint32_t)std::max(1, per_sm) * (e->sm_mac ? e->sm_mac : (uint32_t)sms);
                if (const char *sl = getenv("CA_MAC_SLOTS")) t.p_slots = (uint32_t)std::max(1, atoi(sl));
            }
        }
        // split of the row list per instance: enough CTAs to cover the SMs when few instances run
        // (latency schedule), 1 when the batch alone fills the machine.
        uint32_t split = j == 0 ? cfg->mac_split : 0;
        if (t.p_force) split = 1;  // CA_MAC_PERSIST=1 (tests): every tier on the persistent schedule
        if (!split) {
            // few instances: cut the row list so that ~2 CTAs per SM each stream >= 256 KB
            const uint64_t ctas = (uint64_t)e->n_inst * t.tiles;
            const uint64_t bytes = (uint64_t)8 * t.S * t.P * (e->n_in * e->n_out + e->n_in) * e->n_inst;
            const uint64_t want = std::min<uint64_t>(2 * (uint64_t)sms, std::max<uint64_t>(1, bytes / (256u << 10)));
            split = ctas >= want ? 1u : (uint32_t)std::min<uint64_t>(256, (want + ctas - 1) / ctas);
        }
        const uint32_t rows = t.P * e->n_in;  // steady state: one voice per input
        t.n_split = std::max<uint32_t>(1, std::min(split, std::max<uint32_t>(1, rows / (uint32_t)t.mac.kc)));
        t.h_bytes = (size_t)cfg->n_ir_slots * e->n_out * t.P * t.S * sizeof(float2);
        t.x_bytes = (size_t)e->n_voices * t.Lring * t.S * sizeof(float2);
        e->arena_bytes += t.h_bytes + t.x_bytes;
        s_max = std::max(s_max, t.S);
        reach = std::max(reach, t.off + t.S);
    }
    {
        const char *fz = getenv("CA_FUSE");
        if (!fz && (cfg->schedule & CA_SCHED_FUSED_TIER0)) fz = "1";
        if (!fz && (cfg->schedule & CA_SCHED_NO_FUSED_TIER0)) fz = "0";
        fused_fn fn = e->n_out == 1 ? e->fft.fused1 : e->fft.fused2;
        e->fused = e->tiers.size() > 1 && fn && e->tiers[0].n_split == 1 && e->tiers[0].tiles == 1 && e->n_in * e->nv <= 4 &&
                   !e->quirks && (fz ? fz[0] == '1' : e->n_inst <= 16);  // measured: -2 us p50 for one instance, but 178 vs 132 us at 4096 instances
                                                            // (2 CTAs/SM cannot hide the per-instance FFT latency chain)
        if (e->fused) {
            e->fused_smem = e->n_out == 1 ? (e->R == 8 ? FusedCfg<8, 1>::SMEM_BYTES : e->R == 4 ? FusedCfg<4, 1>::SMEM_BYTES : e->R == 2 ? FusedCfg<2, 1>::SMEM_BYTES : FusedCfg<1, 1>::SMEM_BYTES)
                                          : (e->R == 8 ? FusedCfg<8, 2>::SMEM_BYTES : e->R == 4 ? FusedCfg<4, 2>::SMEM_BYTES : e->R == 2 ? FusedCfg<2, 2>::SMEM_BYTES : FusedCfg<1, 2>::SMEM_BYTES);
            CA_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->fused_smem));
        }
    }
    if (e->tiers.size() > 1) {
        CA_CUDA(cudaFuncSetAttribute((const void *)k_tier_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(s_max * sizeof(float2))));
        CA_CUDA(cudaFuncSetAttribute((const void *)k_tier_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(s_max * sizeof(float2))));
        CA_CUDA(cudaFuncSetAttribute((const void *)k_tier_ir, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(s_max * sizeof(float2))));
    }
    // time-domain ring per voice: predelay reach + the longest overlap-save window + one period: with
    // CA_FLAG_ASYNC_TIERS a tier forward of t_end reads [t_end*B - 2S, t_end*B) WHILE the forward of period
    // t_end clears [t_end*B + 8192, +B) and scatter-adds up to t_end*B + 8191 + B; those must not alias
    e->ring_len = 1;
    while (e->ring_len < kMaxPredelay + 2 * s_max + e->B) e->ring_len <<= 1;
    uint32_t total_periods = 0;
    for (auto &t : e->tiers) total_periods = std::max(total_periods, (t.off + t.P * t.S) / e->B + 2 * t.m);
    e->ring_out = std::max(total_periods + e->k_off, e->ring_len / e->B) + 2;
    e->acc_len = 1;
    while (e->acc_len < reach + e->B) e->acc_len <<= 1;

    CA_CUDA(cudaMalloc(&e->d_arena, e->arena_bytes));
    CA_CUDA(cudaMemsetAsync(e->d_arena, 0, e->arena_bytes, e->stream));
    e->device_bytes = e->arena_bytes;
    {
        unsigned char *p = e->d_arena;
        for (auto &t : e->tiers) { t.H = reinterpret_cast<float2 *>(p); p += t.h_bytes; }
        for (auto &t : e->tiers) { t.X = reinterpret_cast<float2 *>(p); p += t.x_bytes; }
    }
    for (auto &t : e->tiers) {
        // long tiers are phase-staggered: at most ceil(n_inst / m) instances fire per period
        const size_t yp_inst = t.m > 1 ? (e->n_inst + t.m - 1) / t.m : e->n_inst;
        const size_t yp_bytes = yp_inst * t.n_split * e->n_out * t.S * sizeof(float2);
        CA_CUDA(cudaMalloc(&t.Ypart, yp_bytes));
        CA_CUDA(cudaMemsetAsync(t.Ypart, 0, yp_bytes, e->stream));
        e->device_bytes += yp_bytes;
        if (t.m > 1 && e->pipe_mode != 0) {  // pipelined batch schedule: the long tiers' partial sums live one period longer
            CA_CUDA(cudaMalloc(&t.Ypart2, yp_bytes));
            CA_CUDA(cudaMemsetAsync(t.Ypart2, 0, yp_bytes, e->stream));
            e->device_bytes += yp_bytes;
        }
        CA_CUDA(cudaMalloc(&t.workctr, 4 * sizeof(uint32_t)));
        CA_CUDA(cudaMemsetAsync(t.workctr, 0, 4 * sizeof(uint32_t), e->stream));
        // twiddles, fp64 -> fp32: [W_S^n, n < S | W_2S^k, k < S]
        std::vector<float2> tw(2 * (size_t)t.S);
        for (uint32_t n = 0; n < t.S; n++) {
            const double a = -2.0 * M_PI * (double)n / (double)t.S, b = -M_PI * (double)n / (double)t.S;
            tw[n] = make_float2((float)cos(a), (float)sin(a));
            tw[t.S + n] = make_float2((float)cos(b), (float)sin(b));
        }
        CA_CUDA(cudaMalloc(&t.tw, tw.size() * sizeof(float2)));
        CA_CUDA(cudaMemcpyAsync(t.tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, e->stream));  // pageable: staged before return
    }
    const size_t ring_bytes = (size_t)e->n_voices * e->ring_len * sizeof(float);
    {
        const size_t words = std::max<size_t>(1, (e->n_extra + 31u) / 32u);
        CA_CUDA(cudaMalloc(&e->d_vpool, words * sizeof(uint32_t)));
        CA_CUDA(cudaMemsetAsync(e->d_vpool, 0, words * sizeof(uint32_t), e->stream));
    }
    CA_CUDA(cudaMalloc(&e->d_ring, ring_bytes));
    CA_CUDA(cudaMemsetAsync(e->d_ring, 0, ring_bytes, e->stream));
    size_t acc_bytes = 0;
    if (e->tiers.size() > 1) {
        acc_bytes = (size_t)e->n_inst * e->n_out * e->acc_len * sizeof(float);
        CA_CUDA(cudaMalloc(&e->d_acc, acc_bytes));
        CA_CUDA(cudaMemsetAsync(e->d_acc, 0, acc_bytes, e->stream));
    }
    const size_t in_bytes = n_items * e->B * sizeof(float), out_bytes = (size_t)e->n_inst * e->n_out * e->B * sizeof(float);
    CA_CUDA(cudaMalloc(&e->d_in, in_bytes));
    CA_CUDA(cudaMalloc(&e->d_out, out_bytes));
    CA_CUDA(cudaMalloc(&e->d_par, n_items * sizeof(InParamDev)));
    CA_CUDA(cudaMalloc(&e->d_st, 2 * n_items * sizeof(ItemState)));
    CA_CUDA(cudaMemsetAsync(e->d_st, 0, 2 * n_items * sizeof(ItemState), e->stream));
    CA_CUDA(cudaMalloc(&e->d_ctl, sizeof(Ctl)));
    CA_CUDA(cudaMemsetAsync(e->d_ctl, 0, sizeof(Ctl), e->stream));
    {
        // twiddles of the 256-point row FFT, fp64 -> fp32: [W_256^n | W_512^k | W_256^(l q) at q * 16 + l]
        std::vector<float2> tw(768);
        for (int n = 0; n < 256; n++) {
            const double a = -2.0 * M_PI * n / 256.0, b = -2.0 * M_PI * n / 512.0;
            const double c = -2.0 * M_PI * (((n >> 4) * (n & 15)) % 256) / 256.0;  // between the radix-16 stages (fft_rows16.cuh)
            tw[n] = make_float2((float)cos(a), (float)sin(a));
            tw[256 + n] = make_float2((float)cos(b), (float)sin(b));
            tw[512 + n] = make_float2((float)cos(c), (float)sin(c));
        }
        CA_CUDA(cudaMalloc(&e->d_rowtw, tw.size() * sizeof(float2)));
        CA_CUDA(cudaMemcpyAsync(e->d_rowtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, e->stream));
        CA_CUDA(cudaStreamSynchronize(e->stream));  // pageable source: staged before `tw` goes out of scope
    }
    if (e->quirks) {
        e->q_kr = cfg->ref_fft_size / e->B + kMaxPredelay / e->B + 4;
        e->q_len = 1;
        while (e->q_len < kMaxPredelay + 2 * e->B) e->q_len <<= 1;
        const size_t nd = (size_t)e->n_inst * e->q_kr * 4, nq = (size_t)e->n_inst * 2 * e->q_len;
        CA_CUDA(cudaMalloc(&e->d_irsum, (size_t)cfg->n_ir_slots * 4 * sizeof(double)));
        CA_CUDA(cudaMalloc(&e->d_qdelta, nd * sizeof(double)));
        CA_CUDA(cudaMalloc(&e->d_qrun, (size_t)e->n_inst * 4 * sizeof(double)));
        CA_CUDA(cudaMalloc(&e->d_qring, nq * sizeof(float)));
        CA_CUDA(cudaMemsetAsync(e->d_irsum, 0, (size_t)cfg->n_ir_slots * 4 * sizeof(double), e->stream));
        CA_CUDA(cudaMemsetAsync(e->d_qdelta, 0, nd * sizeof(double), e->stream));
        CA_CUDA(cudaMemsetAsync(e->d_qrun, 0, (size_t)e->n_inst * 4 * sizeof(double), e->stream));
        CA_CUDA(cudaMemsetAsync(e->d_qring, 0, nq * sizeof(float), e->stream));
    }
    CA_CUDA(cudaMallocHost(&e->h_in, in_bytes));
    CA_CUDA(cudaMallocHost(&e->h_out, out_bytes));
    for (auto &u : e->h_upload) CA_CUDA(cudaMallocHost(&u, n_items * sizeof(InParamDev)));
    e->device_bytes += ring_bytes + acc_bytes + in_bytes + out_bytes + n_items * (sizeof(InParamDev) + 2 * sizeof(ItemState));

    // parameter defaults == Convolution::CC::value defaults (conv.h:42-50)
    e->par.assign(n_items, InParamDev{});
    e->user.resize(n_items);
    e->par_queue.resize(4 * n_items);
    for (size_t i = 0; i < n_items; i++) {
        ca_params u{};
        u.select = 0; u.predelay = 0; u.speed = 100; u.vsteps = -1;
        u.dry = 0.5f; u.wet = 0.5f; u.panDry = 0.f; u.panWet = 0.f; u.level = 1.0f;
        e->user.set(i, u);
        fill_dev_param(e->par[i], u);
    }
    e->ir_loaded.reset(new std::atomic<uint8_t>[cfg->n_ir_slots]);
    for (uint32_t i = 0; i < cfg->n_ir_slots; i++) e->ir_loaded[i].store(0, std::memory_order_relaxed);
    e->wall.assign(1u << 16, 0.f);
    e->deadline_us = cfg->sample_rate > 0 ? 1e6 * (double)e->B / (double)cfg->sample_rate : 0.0;

    if (cfg->flags & CA_FLAG_L2_PERSIST) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, cfg->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, cfg->device);
        if (max_persist > 0 && max_window > 0) {
            const size_t want = std::min<size_t>(e->arena_bytes, (size_t)max_persist);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = e->d_arena;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(e->arena_bytes, (size_t)max_window);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)av.accessPolicyWindow.num_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
        }
    }
    if (cfg->flags & CA_FLAG_PERSISTENT) {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device);
        if (e->n_inst != 1 || e->tiers.size() != 1 || e->B > 256 || !coop || !persist_pick(e->R, e->n_out) ||
            (cfg->flags & (CA_FLAG_GRAPH | CA_FLAG_PROFILE | CA_FLAG_ASYNC_TIERS))) {
            g_last_error = "CA_FLAG_PERSISTENT: one instance, uniform partitioning, period <= 256, no graph / profile / async-tier flags";
            return CA_ERR_UNSUPPORTED;
        }
        e->persistent = true;
        const uint32_t rows = e->tiers[0].P * e->n_in;
        e->p_ctas = (uint32_t)std::max(1, std::min<int>(sms - 4, (int)(rows / 12)));  // <= 8-row batches x 1..2 per CTA; a few SMs stay free
        CA_CUDA(cudaHostAlloc(&e->pbox, sizeof(PersistBox), cudaHostAllocMapped));
        memset((void *)e->pbox, 0, sizeof(PersistBox));
        CA_CUDA(cudaMalloc(&e->d_go, sizeof(unsigned long long)));
        CA_CUDA(cudaMalloc(&e->d_arrive, 2 * sizeof(unsigned int)));
        CA_CUDA(cudaMalloc(&e->d_pYsum, (size_t)e->n_out * e->B * sizeof(float2)));
        e->p_stamp = getenv("CA_PERSIST_STAMPS") != nullptr;  // diagnostics: phase timestamps (ca_persist_stamps)
        CA_CUDA(cudaMalloc(&e->d_ppar, 2 * sizeof(InParamDev)));
        CA_CUDA(cudaMemsetAsync(e->d_ppar, 0, 2 * sizeof(InParamDev), e->stream));
        CA_CUDA(cudaMalloc(&e->d_pY, (size_t)e->p_ctas * e->n_out * e->B * sizeof(float2)));
    }
    CA_CUDA(cudaStreamSynchronize(e->stream));
    return prewarm_graphs(e);
}

int ca_create(const ca_config *cfg, ca_engine **out)
{
    if (!cfg || !out) return CA_ERR_INVALID;
    *out = nullptr;
    if (cfg->struct_size != sizeof(ca_config)) { g_last_error = "ca_config.struct_size mismatch"; return CA_ERR_INVALID; }
    if (!is_pow2(cfg->period) || cfg->period < 32 || cfg->period > 1024) { g_last_error = "period must be a power of two in [32, 1024]"; return CA_ERR_INVALID; }
    if (cfg->n_in < 1 || cfg->n_in > 2 || cfg->n_out < 1 || cfg->n_out > 2) { g_last_error = "n_in / n_out must be 1 or 2"; return CA_ERR_INVALID; }
    if (cfg->max_voices > (uint32_t)kMaxVoices) { g_last_error = "max_voices must be <= 4"; return CA_ERR_INVALID; }
    if (!cfg->n_instances || !cfg->max_ir_frames || !cfg->n_ir_slots) { g_last_error = "n_instances, max_ir_frames, n_ir_slots must be > 0"; return CA_ERR_INVALID; }
    // item indices (instance x input / output) and per-launch instance counts are 32-bit: far above what 180 GB hold at any geometry worth running
    if (cfg->n_instances > (1u << 20)) { g_last_error = "n_instances must be <= 1048576"; return CA_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
        (void)cudaGetLastError();
        g_last_error = "no usable CUDA device (this engine has no CPU fallback)";
        return CA_ERR_CUDA;
    }
    ca_engine *e = new (std::nothrow) ca_engine();
    if (!e) return CA_ERR_NOMEM;
    int rc = create_impl(cfg, e);
    if (rc) { std::string keep = g_last_error; ca_destroy(e); g_last_error = keep; (void)cudaGetLastError(); return rc; }
    *out = e;
    return CA_OK;
}

static int load_ir_dev(ca_engine *e, uint32_t slot, const float *d_left, const float *d_right, uint32_t frames, uint32_t stride, bool foreign)
{
    if (!e || !d_left || slot >= e->cfg.n_ir_slots) return CA_ERR_INVALID;
    if (e->n_out == 2 && !d_right) return CA_ERR_INVALID;
    CA_CUDA(cudaSetDevice(e->device));
    // caller-produced device data (e.g. WavFile::buffer, filled by copies/kernels on the legacy or
    // another stream) must be complete before this engine's non-blocking stream reads it;
    // the reference's prepare() does the same (conv.cu:237)
    {
        const int rc = persist_stop(e);  // the resident kernel reads the IR bank: it is relaunched by the next period
        if (rc) return rc;
    }
    if (foreign) CA_CUDA(cudaDeviceSynchronize());
    {
        const int rc = drain_all(e);  // lane M / the asynchronous tiers may still read the IR bank
        if (rc) return rc;
    }
    frames = std::min(frames, e->cfg.max_ir_frames);  // truncation like conv.cu:239
    if (e->quirks) k_ir_sums<<<1, 256, 0, e->stream>>>(d_left, d_right, frames, stride, e->d_irsum + (size_t)slot * 4);
    for (size_t j = 0; j < e->tiers.size(); j++) {
        const Tier &t = e->tiers[j];
        float2 *H = t.H + (size_t)slot * e->n_out * t.P * t.S;
        const float scale = 1.0f / (2.0f * (float)t.S);  // both FFT normalisations live in H
        if (j == 0) {
            IrArgs a{};
            a.stride = stride;
            a.h[0] = d_left; a.h[1] = d_right ? d_right : d_left;
            a.H = H; a.twM = t.tw; a.tw2M = t.tw + t.S;
            a.frames = frames; a.P = t.P; a.n_out = e->n_out; a.frame_off = e->k_off * e->B; a.scale = scale;
            const uint32_t items = e->n_out * t.P;
            e->fft.ir<<<(items + kFwdWarps - 1) / kFwdWarps, kFwdWarps * 32, 0, e->stream>>>(a);
        } else {
            TierIrArgs a{};
            a.stride = stride;
            a.h[0] = d_left; a.h[1] = d_right ? d_right : d_left;
            a.H = H; a.twM = t.tw; a.tw2M = t.tw + t.S;
            a.frames = frames; a.P = t.P; a.n_out = e->n_out; a.frame_off = t.off; a.S = t.S; a.s_log = t.s_log; a.scale = scale;
            k_tier_ir<<<e->n_out * t.P, std::min<uint32_t>(kTierThreads, std::max<uint32_t>(tier_min(), t.S / tier_div())), t.S * sizeof(float2), e->stream>>>(a);
        }
        e->launches += 1;
    }
    CA_CUDA(cudaGetLastError());
    CA_CUDA(cudaStreamSynchronize(e->stream));
    e->ir_loaded[slot].store(1, std::memory_order_release);
    return CA_OK;
}

int ca_load_ir_device(ca_engine *e, uint32_t slot, const float *d_left, const float *d_right, uint32_t frames)
{
    return load_ir_dev(e, slot, d_left, d_right, frames, 1, true);
}

int ca_load_ir_interleaved_device(ca_engine *e, uint32_t slot, const float *d_lr, uint32_t frames)
{
    if (!d_lr) return CA_ERR_INVALID;
    return load_ir_dev(e, slot, d_lr, d_lr + 1, frames, 2, true);
}

int ca_load_ir(ca_engine *e, uint32_t slot, const float *left, const float *right, uint32_t frames)
{
    if (!e || !left || slot >= e->cfg.n_ir_slots || !frames) return CA_ERR_INVALID;
    if (e->n_out == 2 && !right) return CA_ERR_INVALID;
    CA_CUDA(cudaSetDevice(e->device));
    const uint32_t n = std::min(frames, e->cfg.max_ir_frames);
    float *d = nullptr;
    CA_CUDA(cudaMalloc(&d, (size_t)2 * n * sizeof(float)));
    // stream-ordered uploads: a plain cudaMemcpy from pageable memory returns once the data is staged,
    // its DMA runs on the legacy stream and is NOT ordered before kernels of this non-blocking stream
    cudaError_t rc = cudaMemcpyAsync(d, left, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, e->stream);
    if (rc == cudaSuccess && right) rc = cudaMemcpyAsync(d + n, right, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, e->stream);
    int r = CA_OK;
    if (rc != cudaSuccess) { g_last_error = cudaGetErrorString(rc); r = CA_ERR_CUDA; }
    else r = load_ir_dev(e, slot, d, right ? d + n : nullptr, n, 1, false);
    cudaFree(d);
    return r;
}

int ca_set_params(ca_engine *e, uint32_t instance, uint32_t input, const ca_params *p)
{
    if (!e || !p || instance >= e->n_inst || input >= e->n_in) return CA_ERR_INVALID;
    if (p->select >= e->cfg.n_ir_slots || p->predelay >= CA_MAX_PREDELAY) return CA_ERR_INVALID;
    if (!e->ir_loaded[p->select].load(std::memory_order_acquire)) return CA_ERR_STATE;  // the reference would dereference nullptr (conv.cu:340)
    ParamCmd c;
    c.item = instance * e->n_in + input; c.kind = 0; c.p = *p;
    if (!e->par_queue.push(c)) { g_last_error = "parameter queue full (no ca_process call is draining it)"; return CA_ERR_STATE; }
    ca_params shown = *p;
    shown.vsteps = -1;
    e->user.set(c.item, shown);
    return CA_OK;
}

int ca_get_params(ca_engine *e, uint32_t instance, uint32_t input, ca_params *p)
{
    if (!e || !p || instance >= e->n_inst || input >= e->n_in) return CA_ERR_INVALID;
    e->user.get((size_t)instance * e->n_in + input, p);
    return CA_OK;
}

int ca_set_glide(ca_engine *e, uint32_t instance, uint32_t input, float g)
{
    if (!e || instance >= e->n_inst || input >= e->n_in) return CA_ERR_INVALID;
    ParamCmd c;
    c.item = instance * e->n_in + input; c.kind = 1; c.glide = g;
    if (!e->par_queue.push(c)) { g_last_error = "parameter queue full (no ca_process call is draining it)"; return CA_ERR_STATE; }
    return CA_OK;
}

// instances [a0, a1) start like new ones: voices restart at the current period (so every older delay-line block is
// skipped, and a voice's first forward clears its time ring), pending tier output and the reference-compatibility
// terms are dropped; shared cross-fade voices go back to the pool.  The engine must be drained.
static int reset_instances(ca_engine *e, size_t a0, size_t a1)
{
    if (a1 <= a0) return CA_OK;
    const size_t i0 = a0 * e->n_in, cnt = (a1 - a0) * e->n_in, n_alloc = (size_t)e->n_inst * e->n_in, na = a1 - a0;
    k_release_voices<<<(unsigned)((cnt + 127) / 128), 128, 0, e->stream>>>(e->d_st + (e->t_host & 1) * n_alloc + i0, (uint32_t)cnt, voice_pool(e));
    for (int b = 0; b < 2; b++) CA_CUDA(cudaMemsetAsync(e->d_st + b * n_alloc + i0, 0, cnt * sizeof(ItemState), e->stream));
    if (e->d_acc) CA_CUDA(cudaMemsetAsync(e->d_acc + a0 * e->n_out * e->acc_len, 0, na * e->n_out * e->acc_len * sizeof(float), e->stream));
    if (e->quirks) {
        CA_CUDA(cudaMemsetAsync(e->d_qdelta + a0 * e->q_kr * 4, 0, na * e->q_kr * 4 * sizeof(double), e->stream));
        CA_CUDA(cudaMemsetAsync(e->d_qrun + a0 * 4, 0, na * 4 * sizeof(double), e->stream));
        CA_CUDA(cudaMemsetAsync(e->d_qring + a0 * 2 * e->q_len, 0, na * 2 * e->q_len * sizeof(float), e->stream));
    }
    CA_CUDA(cudaStreamSynchronize(e->stream));
    return CA_OK;
}

int ca_set_active(ca_engine *e, uint32_t n)
{
    if (!e || !n || n > e->n_inst) return CA_ERR_INVALID;
    if (n == e->n_active) return CA_OK;
    if (e->persistent) return CA_ERR_UNSUPPORTED;
    CA_BIND(e);
    int rc = drain_all(e);
    if (rc) return rc;
    CA_CUDA(cudaStreamSynchronize(e->stream));
    // instances that were parked keep frozen voice state, delay lines and tier output: a reactivated instance must
    // not replay pre-deactivation audio as a tail
    rc = reset_instances(e, e->n_active, n);
    if (rc) return rc;
    e->n_active = n;
    return prewarm_graphs(e);  // not a real-time call: rebuild the graphs for the new batch size now
}

int ca_reset(ca_engine *e)
{
    if (!e) return CA_ERR_INVALID;
    if (e->persistent) return CA_ERR_UNSUPPORTED;
    CA_CUDA(cudaSetDevice(e->device));
    const int rc = drain_all(e);
    if (rc) return rc;
    CA_CUDA(cudaStreamSynchronize(e->stream));
    return reset_instances(e, 0, e->n_active);
}

int ca_process_device(ca_engine *e, const float *d_in, float *d_out, uint32_t nframes)
{
    if (!e || !d_in || !d_out) return CA_ERR_INVALID;
    if (nframes != e->B) return CA_ERR_PERIOD;
    if (e->persistent) { g_last_error = "CA_FLAG_PERSISTENT engines are driven through ca_process"; return CA_ERR_UNSUPPORTED; }
    CA_BIND(e);
    const double t0 = now_us();
    int rc;
    if (use_pipeline(e)) {
        rc = run_pipelined(e, d_in, d_out, 1, nullptr, nullptr);
        if (rc) return rc;
    } else {
        rc = run_period(e, d_in, d_out);
        if (rc) return rc;
        rc = run_deferred(e);
        if (rc) return rc;
    }
    record_wall(e, now_us() - t0);
    return CA_OK;
}

int ca_process(ca_engine *e, const float *in, float *out, uint32_t nframes)
{
    if (!e || !in || !out) return CA_ERR_INVALID;  // the reference silently returns on null ports (conv.cu:297)
    if (nframes != e->B) return CA_ERR_PERIOD;
    CA_BIND(e);
    const double t0 = now_us();
    if (e->persistent) {
        const int rc = process_persistent(e, in, out);
        if (rc) return rc;
        record_wall(e, now_us() - t0);
        return CA_OK;
    }
    const size_t in_stride = (size_t)e->n_in * e->B, out_stride = (size_t)e->n_out * e->B;  // floats per instance
    const size_t in_bytes = e->n_active * in_stride * sizeof(float), out_bytes = e->n_active * out_stride * sizeof(float);
    const float *src = in;
    if (!is_pinned(e, in)) { memcpy(e->h_in, in, in_bytes); src = e->h_in; }
    float *dst = is_pinned(e, out) ? out : e->h_out;
    const bool graph = (e->cfg.flags & CA_FLAG_GRAPH) != 0, profile = (e->cfg.flags & CA_FLAG_PROFILE) != 0;
    // Large batches: cut the instances into chunks and pipeline H2D | kernels | D2H on three streams so
    // the PCIe copies (2 KB per instance and direction) hide behind the kernels of the other chunks.
    const uint32_t chunks = (graph || profile || e->n_active < 512 || e->async_tiers) ? 1u : std::min<uint32_t>(e->io_chunks, e->n_active / 256);
    int rc = CA_OK;
    if (use_pipeline(e)) {
        rc = run_pipelined(e, e->d_in, e->d_out, std::max(1u, chunks), src, dst);
        if (rc) return rc;
        CA_CUDA(cudaEventSynchronize(e->out_ready));
        if (dst != out) memcpy(out, e->h_out, out_bytes);
        record_wall(e, now_us() - t0);
        return CA_OK;
    }
    if (chunks <= 1) {
        CA_CUDA(cudaMemcpyAsync(e->d_in, src, in_bytes, cudaMemcpyHostToDevice, e->stream));
        rc = run_period(e, e->d_in, e->d_out);
        if (rc) return rc;
        CA_CUDA(cudaMemcpyAsync(dst, e->d_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
        CA_CUDA(cudaEventRecord(e->out_ready, e->stream));
    } else {
        rc = flush_params(e);
        if (rc) return rc;
        // CA_IO_TRACE=n: device timeline of call n (development): when each chunk's upload, kernels and download end
        static const long trace_at = getenv("CA_IO_TRACE") ? atol(getenv("CA_IO_TRACE")) : -1;
        const bool trace = trace_at >= 0 && (long)e->t_host == trace_at;
        static cudaEvent_t tr_base, tr_ev[3][kIoChunks], tr_tiers, tr_prev;
        static bool tr_init = false;
        if (trace_at >= 0 && !tr_init) {
            cudaEventCreate(&tr_base); cudaEventCreate(&tr_tiers); cudaEventCreate(&tr_prev);
            for (auto &row : tr_ev) for (auto &ev : row) cudaEventCreate(&ev);
            tr_init = true;
        }
        if (trace) { cudaEventRecord(tr_base, e->s_in); cudaEventRecord(tr_prev, e->stream); }
        // d_in / d_out are free: the previous call returned only after its D2H (hence every kernel of its
        // period pipeline) had completed; the deferred tiers still running on e->stream touch neither.
        for (uint32_t c = 0; c < chunks; c++) {
            const uint32_t i0 = (uint32_t)((uint64_t)e->n_active * c / chunks), i1 = (uint32_t)((uint64_t)e->n_active * (c + 1) / chunks);
            CA_CUDA(cudaMemcpyAsync(e->d_in + i0 * in_stride, src + i0 * in_stride, (i1 - i0) * in_stride * sizeof(float), cudaMemcpyHostToDevice, e->s_in));
            CA_CUDA(cudaEventRecord(e->io_ev[0][c], e->s_in));
            if (trace) cudaEventRecord(tr_ev[0][c], e->s_in);
        }
        // (Alternating chunks on two streams, so that one chunk's kernels fill the ramp and tail of the other's, was
        // measured and dropped: 1 149 -> 1 204 us with 3 chunks, 1 135 -> 1 159 us with 4 -- the uploads and the deferred
        // tiers slow down by more than the tier-0 kernels gain.)
        for (uint32_t c = 0; c < chunks; c++) {
            const uint32_t i0 = (uint32_t)((uint64_t)e->n_active * c / chunks), i1 = (uint32_t)((uint64_t)e->n_active * (c + 1) / chunks);
            CA_CUDA(cudaStreamWaitEvent(e->stream, e->io_ev[0][c], 0));
            rc = launch_period(e, e->d_in, e->d_out, false, i0, i1, c + 1 == chunks);
            if (rc) return rc;
            CA_CUDA(cudaEventRecord(e->io_ev[1][c], e->stream));
            if (trace) cudaEventRecord(tr_ev[1][c], e->stream);
            CA_CUDA(cudaStreamWaitEvent(e->s_out, e->io_ev[1][c], 0));
            CA_CUDA(cudaMemcpyAsync(dst + i0 * out_stride, e->d_out + i0 * out_stride, (i1 - i0) * out_stride * sizeof(float), cudaMemcpyDeviceToHost, e->s_out));
            if (trace) cudaEventRecord(tr_ev[2][c], e->s_out);
        }
        e->launches += e->fused ? chunks + 1 : 3 * chunks;
        CA_CUDA(cudaEventRecord(e->out_ready, e->s_out));
        if (trace) {
            const double t_enq = now_us() - t0;
            rc = run_deferred(e);
            if (rc) return rc;
            const double t_enq2 = now_us() - t0;
            cudaEventRecord(tr_tiers, e->stream);
            CA_CUDA(cudaEventSynchronize(e->out_ready));
            const double t_ret = now_us() - t0;
            cudaEventSynchronize(tr_tiers);
            auto el = [&](cudaEvent_t ev) { float ms = 0; cudaEventElapsedTime(&ms, tr_base, ev); return 1e3 * ms; };
            fprintf(stderr, "[io trace] call %ld: host enqueue tier0 %.0f us, + tiers %.0f us, returns at %.0f us | previous tiers end %.0f\n", trace_at, t_enq, t_enq2, t_ret, el(tr_prev));
            for (uint32_t c = 0; c < chunks; c++) fprintf(stderr, "[io trace]   chunk %u: upload done %.0f, kernels done %.0f, download done %.0f\n", c, el(tr_ev[0][c]), el(tr_ev[1][c]), el(tr_ev[2][c]));
            fprintf(stderr, "[io trace]   this call's tiers end %.0f\n", el(tr_tiers));
            if (dst != out) memcpy(out, e->h_out, out_bytes);
            record_wall(e, now_us() - t0);
            return CA_OK;
        }
    }
    rc = run_deferred(e);  // long tiers keep the GPU busy while the host already has its output
    if (rc) return rc;
    CA_CUDA(cudaEventSynchronize(e->out_ready));
    if (dst != out) memcpy(out, e->h_out, out_bytes);
    record_wall(e, now_us() - t0);
    return CA_OK;
}

int ca_sync(ca_engine *e)
{
    if (!e) return CA_ERR_INVALID;
    if (e->persistent) return CA_OK;  // ca_process is synchronous; the resident kernel keeps the stream busy by design
    CA_BIND(e);
    const int rc = drain_all(e);
    if (rc) return rc;
    CA_CUDA(cudaStreamSynchronize(e->stream));
    return CA_OK;
}

void *ca_stream(ca_engine *e) { return e ? (void *)e->stream : nullptr; }

int ca_get_stats(ca_engine *e, ca_stats *s)
{
    if (!e || !s) return CA_ERR_INVALID;
    memset(s, 0, sizeof(*s));
    s->periods = e->periods; s->xruns = e->xruns;
    const size_t n = (size_t)std::min<uint64_t>(e->periods, e->wall.size());
    if (n) {
        std::vector<float> v(e->wall.begin(), e->wall.begin() + n);
        std::sort(v.begin(), v.end());
        s->p50_us = v[n / 2];
        s->p99_us = v[std::min(n - 1, (size_t)std::ceil(0.99 * (double)n))];
        s->max_us = e->wall_max;
        s->mean_us = e->wall_sum / (double)e->periods;
    }
    if (e->prof_n) {
        s->fwd_us = e->prof_us[0] / (double)e->prof_n;
        s->mac_us = e->prof_us[1] / (double)e->prof_n;
        s->inv_us = e->prof_us[2] / (double)e->prof_n;
        s->tiers_us = e->prof_us[3] / (double)e->prof_n;
        s->tier_fwd_us = e->tier_prof_us[0] / (double)e->prof_n;
        s->tier_mac_us = e->tier_prof_us[1] / (double)e->prof_n;
        s->tier_inv_us = e->tier_prof_us[2] / (double)e->prof_n;
        s->total_us = s->fwd_us + s->mac_us + s->inv_us + s->tiers_us;
    }
    s->gpu_launches = e->launches;
    // the MAC streams every IR partition spectrum and every FDL slot once per firing
    const Tier &t0 = e->tiers[0];
    s->mac_bytes = (uint64_t)8 * t0.S * t0.P * (uint64_t)(e->n_in * e->n_out + e->n_in) * e->n_active;
    double amort = 0;
    for (auto &t : e->tiers) amort += 8.0 * t.S * t.P * (double)(e->n_in * e->n_out + e->n_in) / (double)t.m;
    s->mac_bytes_amortized = (uint64_t)(amort * e->n_active);
    s->partitions = t0.P; s->mac_split = t0.n_split; s->device_bytes = e->device_bytes;
    s->n_tiers = (uint32_t)e->tiers.size();
    s->tier0_fused = e->fused ? 1u : 0u;
    for (size_t j = 0; j < e->tiers.size() && j < CA_MAX_TIERS; j++) { s->tier_block[j] = e->tiers[j].S; s->tier_parts[j] = e->tiers[j].P; s->tier_offset[j] = e->tiers[j].off; }
    return CA_OK;
}

int ca_reset_stats(ca_engine *e)
{
    if (!e) return CA_ERR_INVALID;
    e->periods = e->xruns = 0; e->wall_sum = e->wall_max = 0;
    for (auto &p : e->prof_us) p = 0;
    for (auto &p : e->tier_prof_us) p = 0;
    e->prof_n = 0;
    return CA_OK;
}

// ---- diagnostics: read-only bandwidth sweep (L2-resident vs HBM), the denominator for the
//      single-instance (L2-resident) MAC that MEASURED_PEAKS.json does not contain ----
namespace {
__global__ void __launch_bounds__(256) k_read_sweep(const float4 *__restrict__ p, size_t n4, float *sink)
{
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
        acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
        acc.z += a.z + b.z + c.z + d.z; acc.w += a.w + b.w + c.w + d.w;
    }
    for (; i < n4; i += stride) { const float4 a = p[i]; acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w; }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;  // keep the loads alive
}
}  // namespace

int ca_measure_read_gbs(int device, size_t bytes, int iters, double *gbs)
{
    if (!gbs || bytes < 4096 || iters < 1) return CA_ERR_INVALID;
    CA_CUDA(cudaSetDevice(device));
    float4 *buf = nullptr;
    float *sink = nullptr;
    CA_CUDA(cudaMalloc(&buf, bytes));
    CA_CUDA(cudaMalloc(&sink, sizeof(float)));
    CA_CUDA(cudaMemset(buf, 0, bytes));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaEvent_t e0, e1;
    CA_CUDA(cudaEventCreate(&e0));
    CA_CUDA(cudaEventCreate(&e1));
    const size_t n4 = bytes / sizeof(float4);
    for (int w = 0; w < 3; w++) k_read_sweep<<<sms * 8, 256>>>(buf, n4, sink);
    CA_CUDA(cudaEventRecord(e0));
    for (int it = 0; it < iters; it++) k_read_sweep<<<sms * 8, 256>>>(buf, n4, sink);
    CA_CUDA(cudaEventRecord(e1));
    CA_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *gbs = (double)bytes * iters / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    return CA_OK;
}

int ca_persist_stamps(ca_engine *e, uint64_t stamps_ns[8])
{
    if (!e || !stamps_ns || !e->persistent || !e->pbox) return CA_ERR_INVALID;
    for (int i = 0; i < 8; i++) stamps_ns[i] = e->pbox->stamps[i];
    return CA_OK;
}

int ca_host_alloc(void **p, size_t bytes)
{
    if (!p) return CA_ERR_INVALID;
    CA_CUDA(cudaMallocHost(p, bytes));
    return CA_OK;
}

int ca_host_free(void *p)
{
    CA_CUDA(cudaFreeHost(p));
    return CA_OK;
}

}  // extern "C"

#include "group.cuh"
