// kernels_rows16.cuh -- the period's FFT kernels on the 16 x 16 row FFT (fft_rows16.cuh): two rows per warp.
//
// Same work, same global layouts and same arguments as the kernels of kernels_rows.cuh (which stay as the A/B
// reference, CA_SCHED_ROWS8, and for 256-point tier blocks); what changes is the transform and the mapping of
// rows to lanes:
//   k_fwd0_x2 / k_inv0_x2            tier 0 at B = 256: one HALF-warp per (instance, input | output)
//   k_tfwd_x2 / k_tinv_x2<M1>        long tier of M1 x 256 points, 2 <= M1 <= 16: one CTA of M1 / 2 warps per
//                                    transform; a warp owns a row and its split partner (r, M1 - r)
//   k_trows_fwd_x2 / k_trows_inv_x2  M1 = 32, 64: the row launches between / after the column launches
//                                    (k_tcols_*, unchanged), eight row pairs per CTA
// Replaces cufftExecC2C + f_unpackC22R + f_pack2R2C (conv.cu:35-73, 367, 405-408) like the kernels it mirrors.
#pragma once
#include "fft_rows16.cuh"
#include "kernels_rows.cuh"

namespace ca {

#ifndef CA_X2_MINB
#define CA_X2_MINB 3
#endif
constexpr int kX2MinB = CA_X2_MINB;  // resident CTAs per SM the tier-0 / row kernels are compiled for (register bound)
constexpr int kX2Warps = 8;
constexpr int kX2Threads = kX2Warps * 32;
constexpr int kX2Rows = 2 * kX2Warps;
constexpr uint32_t kX2Smem = kX2Rows * kR16Slots * sizeof(float2);  // 36 864 B

// item_step_warp_from for a half-warp per item: lane l == 0 of the half owns the side effects
__device__ __forceinline__ ItemState item_step_half_from(const ItemState &old, ItemState *st, uint32_t n_items_alloc, uint32_t item, const InParamDev &p,
                                                         unsigned long long t, int nv, uint32_t ring_out, const VoicePool &vp, int lane, bool valid)
{
    ItemState s = step_item_state(old, p, t, nv, ring_out);
    const bool lead = (lane & 15) == 0 && valid;
    if (lead) voice_storage_update(s, old.active, item, nv, vp);
    const int src = lane & 16;
    s.active = __shfl_sync(kFull, s.active, src);
    s.fresh = __shfl_sync(kFull, s.fresh, src);
#pragma unroll
    for (int v = 0; v < kMaxVoices; v++) s.pool[v] = __shfl_sync(kFull, s.pool[v], src);
    if (lead) st[((t + 1ull) & 1ull) * n_items_alloc + item] = s;
    return s;
}

// ------------------------------------------------------------------------------------------
// tier 0, B = 256: forward.  One half-warp per (instance, input); semantics identical to k_forward<8>.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kX2Threads, kX2MinB) k_fwd0_x2(const FwdArgs a)
{
    pdl_trigger();
    pdl_wait();
    constexpr int B = 256;
    extern __shared__ __align__(16) float2 sm[];
    __shared__ R16Tables tb;
    __shared__ float s_vc[kX2Rows][kMaxVoices];
    __shared__ uint32_t s_ve[kX2Rows][kMaxVoices];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, l = lane & 15;
    const int rowi = 2 * warp + h;
    const uint32_t w = (blockIdx.x * kX2Warps + warp) * 2 + h;
    const bool valid = w < a.n_items;
    const uint32_t item = a.item0 + (valid ? w : 0u);  // an idle half shadows item 0's loads and stores nothing
    const unsigned long long t = a.t_host_p1 ? a.t_host_p1 - 1ull : a.ctl->t;
    // parameter / state loads are in flight while the twiddle tables are staged
    const InParamDev p = a.par[item];
    const uint32_t pd = a.par[(item / a.n_in) * a.n_in].predelay;  // input 0's, conv.cu:412,415
    const ItemState old = a.st[(t & 1ull) * a.n_items_alloc + item];
    {
        // the window sits behind two dependent round trips (state -> storage entry -> ring): pull the blocks of the
        // item's HOME entry (where its voice lives outside cross-fades) and the input block into L2 meanwhile
        const uint32_t mask0 = a.ring_len - 1;
        const uint32_t base0 = (uint32_t)((t * (unsigned long long)B) & mask0);
        const float *home = a.ring + (size_t)item * a.ring_len;
        prefetch_l2(home + (l < 8 ? ((base0 - B) & mask0) + 32 * l : base0 + 32 * (l - 8)));
        if (l < 8) prefetch_l2(a.in + (size_t)item * B + 32 * l);
    }
    r16_tables_init(tb, a.rowtw);
    if (!__any_sync(kFull, valid)) return;
    if (w == 0 && l == 0) a.ctl->t_next = t + 1ull;
    const ItemState s = item_step_half_from(old, a.st, a.n_items_alloc, item, p, t, (int)a.nv, a.ring_out, a.vp, lane, valid);
    if (l < kMaxVoices) {
        float cv = 0.f;
        uint32_t ev = 0u;
#pragma unroll
        for (int q = 0; q < kMaxVoices; q++) { cv = (q == l) ? s.c[q] : cv; ev = (q == l) ? s.pool[q] : ev; }
        s_vc[rowi][l] = cv;
        s_ve[rowi][l] = ev;
    }
    const uint32_t active = valid ? s.active : 0u, fresh = s.fresh;
    const float level = p.level;
    __syncwarp();

    float2 *S = sm + rowi * kR16Slots;
    const uint32_t mask = a.ring_len - 1;
    const float *x = a.in + (size_t)item * B;
    const uint32_t base = (uint32_t)((t * (unsigned long long)B) & mask);
    const uint32_t prev = (base - B) & mask;
    const uint32_t slot = (a.Lring - 1u) - (uint32_t)((t + 1ull) % a.Lring);  // the FDL ring runs backwards
    const R16Pair pr = r16_pair(true, lane, h);

#pragma unroll 1
    for (uint32_t v = 0; v < a.nv; v++) {
        const bool act = ((active >> v) & 1u) != 0;
        if (!__any_sync(kFull, act)) continue;  // warp-uniform: both halves take part in the transform's exchanges
        const float gain = s_vc[rowi][v] * level;
        const uint32_t entry = act ? s_ve[rowi][v] - 1u : 0u;
        float *ring = a.ring + (size_t)entry * a.ring_len;
        if (act && ((fresh >> v) & 1u)) {  // (re)allocated voice: its time-domain history belongs to another IR
            for (uint32_t n = 4 * l; n < a.ring_len; n += 64) *reinterpret_cast<float4 *>(ring + n) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        float2 z[16];
        if (act) {
#pragma unroll
            for (int j = 0; j < 4; j++)  // clear the block that becomes reachable by the predelay scatter in this period
                *reinterpret_cast<float4 *>(ring + ((base + kMaxPredelay + 4 * l + 64 * j) & mask)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (__any_sync(kFull, act && pd != 0)) {
            __syncwarp();  // the scatter may reach into the block cleared above
            if (act && pd != 0) {
                // predelay ring: the whole response of this block is delayed by pd samples (conv.cu:97)
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int n = l + 16 * j;
                    ring[(base + pd + n) & mask] += gain * __ldg(&x[n]);
                }
            }
            __syncwarp();
        }
        // window z[n] = (w[2n], w[2n+1]), w = [previous block | current block]; lane holds z[l + 16 b]
        if (act) {
#pragma unroll
            for (int b = 0; b < 8; b++) {
                z[b] = *reinterpret_cast<const float2 *>(ring + prev + 2 * (l + 16 * b));
                z[8 + b] = *reinterpret_cast<const float2 *>(ring + base + 2 * (l + 16 * b));
            }
            if (pd == 0) {
                float2 xi[8];
#pragma unroll
                for (int b = 0; b < 8; b++) xi[b] = *reinterpret_cast<const float2 *>(x + 2 * (l + 16 * b));
#pragma unroll
                for (int b = 0; b < 8; b++) {  // what earlier, delayed blocks left in the current block + this period's input
                    z[8 + b].x = fmaf(gain, xi[b].x, z[8 + b].x);
                    z[8 + b].y = fmaf(gain, xi[b].y, z[8 + b].y);
                    *reinterpret_cast<float2 *>(ring + base + 2 * (l + 16 * b)) = z[8 + b];
                }
            }
        } else {
#pragma unroll
            for (int b = 0; b < 16; b++) z[b] = make_float2(0.f, 0.f);
        }
        fft256x2<false>(z, S, tb, l);
        float2 *dst = a.X + ((size_t)entry * a.Lring + slot) * B + l;
        r16_split_fwd<false>(z, pr, make_float2(1.f, 0.f), tb, l, [&](int q, float2 X) { if (act) dst[16 * q] = X; });
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// tier 0, B = 256: inverse + overlap discard + output ring + clamp + dry mix.  One half-warp per
// (instance, output); semantics identical to k_inverse<8, true>.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kX2Threads, kX2MinB) k_inv0_x2(const InvArgs a)
{
    pdl_trigger();
    pdl_wait();
    constexpr int B = 256;
    extern __shared__ __align__(16) float2 sm[];
    __shared__ R16Tables tb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, l = lane & 15;
    const uint32_t local0 = (blockIdx.x * kX2Warps + warp) * 2 + h;
    const bool valid = local0 < a.n_items;
    const uint32_t local = valid ? local0 : 0u;  // an idle half shadows item 0's loads and stores nothing
    const uint32_t item = a.item0 + local;
    const uint32_t inst = item / a.n_out, o = item % a.n_out;
    const unsigned long long t = a.t_host_p1 ? a.t_host_p1 - 1ull : a.ctl->t_next - 1ull;
    float2 *S = sm + (2 * warp + h) * kR16Slots;
    if (a.n_wait) {  // ca_group root: the peers' spectra arrive over NVLink (see InvArgs)
        group_wait(a.wait_flags, a.n_wait, t + 1ull, a.gerr, lane);
        __syncwarp();
    }
    // partial spectra of the MAC's row-range splits, fixed order; lane holds bins l + 16 b
    float2 y[16];
    {
        const float2 *src = a.Ypart + (((size_t)inst * a.n_split) * a.n_out + o) * B + l;
        const size_t stride = (size_t)a.n_out * B;
#pragma unroll
        for (int b = 0; b < 16; b++) y[b] = src[16 * b];
        for (uint32_t sp = 1; sp < a.n_split; sp++) {
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const float2 r = src[sp * stride + 16 * b];
                y[b].x += r.x; y[b].y += r.y;
            }
        }
    }
    // what the epilogue reads (dry input, output ring of the long tiers) is pulled into L2 while the transform runs:
    // no register can be spared for it.  Kept time samples [B, 2B) = z[128, 256): register p >= 8 of lane l holds
    // output samples 2 (l + 16 (p - 8)), + 1
    const bool raw = a.raw_wet != 0;
    float dg[2] = {0.f, 0.f};
    if (!raw) {  // dry gain per input: dry * panDry * level, conv.cu:418-427
        const InParamDev p0 = a.par[inst * a.n_in];
        const InParamDev p1 = a.par[inst * a.n_in + (a.n_in - 1)];
        dg[0] = p0.dry * pan_gain(p0.panDry, (int)o, (int)a.n_out) * p0.level;
        dg[1] = a.n_in > 1 ? p1.dry * pan_gain(p1.panDry, (int)o, (int)a.n_out) * p1.level : 0.f;
    }
    const float *x0 = a.in + ((size_t)inst * a.n_in) * B + 2 * l;
    const float *x1 = x0 + (a.n_in > 1 ? B : 0);
    float *accp = a.accring ? a.accring + (size_t)item * a.acc_len + (uint32_t)((t * (unsigned long long)B) & (a.acc_len - 1)) + 2 * l : nullptr;
    if (l < 8) {  // 8 lines of 128 B each
        prefetch_l2(x0 - 2 * l + 32 * l);
        if (accp) prefetch_l2(accp - 2 * l + 32 * l);
    } else if (a.n_in > 1) {
        prefetch_l2(x1 - 2 * l + 32 * (l - 8));
    }
    r16_tables_init(tb, a.rowtw);  // after every global load of the prologue has been issued
    if (!__any_sync(kFull, valid)) return;
    const R16Pair pr = r16_pair(true, lane, h);
    r16_split_inv<false>(y, pr, make_float2(1.f, 0.f), tb, l);
    fft256x2<true>(y, S, tb, l);
    if (!valid) return;
    float *dst = a.out + ((size_t)inst * a.n_out + o) * B + 2 * l;
    float2 xa[8], xb[8], accv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        xa[j] = *reinterpret_cast<const float2 *>(x0 + 32 * j);
        xb[j] = *reinterpret_cast<const float2 *>(x1 + 32 * j);
        accv[j] = accp ? *reinterpret_cast<const float2 *>(accp + 32 * j) : make_float2(0.f, 0.f);
    }
    auto clampf = [raw](float w) { return raw ? w : fminf(fmaxf(w, -1.0f), 1.0f); };  // conv.cu:98
#pragma unroll
    for (int j = 0; j < 8; j++) {
        // wet = this period's tier-0 block + what the long (deferred) tiers left for it
        float2 out = make_float2(clampf(y[8 + j].x + accv[j].x), clampf(y[8 + j].y + accv[j].y));
        out.x = fmaf(dg[0], xa[j].x, fmaf(dg[1], xb[j].x, out.x));
        out.y = fmaf(dg[0], xa[j].y, fmaf(dg[1], xb[j].y, out.y));
        *reinterpret_cast<float2 *>(dst + 32 * j) = out;
        if (accp) *reinterpret_cast<float2 *>(accp + 32 * j) = make_float2(0.f, 0.f);  // consumed
    }
    if (a.advance && local0 == 0 && l == 0) { a.ctl->t = t + 1ull; a.ctl->t_def[(t + 1ull) & 1ull] = t + 1ull; }
}

// ------------------------------------------------------------------------------------------
// long tiers, M = S = M1 x 256 complex points of a 2S-sample real window
// ------------------------------------------------------------------------------------------
// rows of row pair `pi` (pi = 0: rows 0 and M1 / 2, both self-paired; else pi and M1 - pi): the row of half h,
// and the half that holds its split partner
template <int M1>
__device__ __forceinline__ int x2_row(int pi, int h) { return pi == 0 ? (h ? M1 / 2 : 0) : (h ? M1 - pi : pi); }
__device__ __forceinline__ int x2_partner_half(int pi, int h) { return pi == 0 ? h : (h ^ 1); }

// The window of a forward tier transform sits behind two dependent round trips (period count -> voice state -> storage
// entry -> ring).  Voice 0's CTA pulls the window of the item's HOME entry (where its voice lives outside cross-fades)
// into L2 while the state is on its way; `tend` must not need a load (host-driven launches).
__device__ __forceinline__ void tier_prefetch_home(const TierFwdArgs &a, uint32_t v, uint32_t i, uint32_t z, uint32_t tid, uint32_t nthreads)
{
    if (v != 0 || !a.tend_host) return;
    const uint32_t item = (a.inst0 + z * a.inst_stride) * a.n_in + i;
    const uint32_t mask = a.ring_len - 1;
    const uint32_t start = (uint32_t)((a.tend_host * (unsigned long long)a.B - 2ull * a.S) & mask);
    const float *home = a.ring + (size_t)item * a.ring_len;
    for (uint32_t line = tid; line < a.S / 16; line += nthreads) prefetch_l2(home + ((start + 32u * line) & mask));  // 2 S floats = S / 16 lines of 128 B
}

// One CTA of M1 / 2 warps per (voice, input, firing instance): window -> column DFTs -> rows -> split -> FDL slot
template <int M1>
__global__ void __launch_bounds__(M1 * 16) k_tfwd_x2(const TierFwdArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];  // [M1][kR16Slots]
    __shared__ R16Tables tb;
    tier_prefetch_home(a, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, M1 * 16);
    const TierCommon c = tier_fwd_common(a, blockIdx.x, blockIdx.y, blockIdx.z);
    if (!c.active) return;  // uniform for the CTA
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, h = lane >> 4, l = lane & 15;
    const uint32_t mask = a.ring_len - 1;
    const float *ring = a.ring + (size_t)c.w * a.ring_len;
    const uint32_t start = (uint32_t)((c.tend * (unsigned long long)a.B - 2ull * a.S) & mask);
    // columns: A[k1][n2] = sum_n1 z[256 n1 + n2] W_M1^(n1 k1)
    for (int n2 = threadIdx.x; n2 < 256; n2 += M1 * 16) {
        float2 col[M1];
#pragma unroll
        for (int n1 = 0; n1 < M1; n1++) col[n1] = *reinterpret_cast<const float2 *>(ring + ((start + 2u * (256u * n1 + n2)) & mask));
        dft_reg<M1, false>(col);
#pragma unroll
        for (int k1 = 0; k1 < M1; k1++) sm[k1 * kR16Slots + n2] = col[k1];
    }
    const int r = x2_row<M1>(wi, h);
    const float2 cr = r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f);
    r16_tables_init(tb, a.rowtw);  // staged after the window loads were issued; its barrier also closes the column phase
    // rows: X[k1 + M1 k2] = sum_n2 A[k1][n2] W_M^(n2 k1) W_256^(n2 k2)
    float2 *reg = sm + r * kR16Slots;
    float2 v[16];
#pragma unroll
    for (int b = 0; b < 16; b++) v[b] = reg[l + 16 * b];
    if (r) {
#pragma unroll
        for (int b = 0; b < 16; b++) v[b] = cmul(v[b], __ldg(&a.twM[(l + 16 * b) * r]));
    }
    __syncwarp();
    fft256x2<false>(v, reg, tb, l);
    const R16Pair pr = r16_pair(r == 0, lane, x2_partner_half(wi, h));
    float2 *dst = a.X + ((size_t)c.w * a.Lring + c.slot) * a.S + (brev_s((uint32_t)r, a.s_log) << 8) + l;
    r16_split_fwd<true>(v, pr, cr, tb, l, [&](int q, float2 X) { dst[16 * q] = X; });
}

// One CTA per (output, firing instance): partial sums -> split -> rows -> column DFTs -> += output ring at +off
template <int M1>
__global__ void __launch_bounds__(M1 * 16) k_tinv_x2(const TierInvArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];
    __shared__ R16Tables tb;
    const uint32_t z = blockIdx.y, o = blockIdx.x;
    const uint32_t inst = a.inst0 + z * a.inst_stride;
    const uint32_t item = inst * a.n_out + o;
    const unsigned long long tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, 0u);
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, h = lane >> 4, l = lane & 15;
    {   // the output-ring block this transform adds into is read at the very end: pull it into L2 now
        const uint32_t amask0 = a.acc_len - 1;
        const uint32_t p0 = (uint32_t)((tend * (unsigned long long)a.B - a.S + a.off) & amask0);
        const float *acc0 = a.accring + (size_t)item * a.acc_len;
        for (uint32_t line = threadIdx.x; line < a.S / 32; line += M1 * 16) prefetch_l2(acc0 + ((p0 + 32u * line) & amask0));
    }
    const int r = x2_row<M1>(wi, h);
    float2 *reg = sm + r * kR16Slots;
    float2 y[16];
    {
        const float2 *src = a.Ypart + (((size_t)z * a.n_split) * a.n_out + o) * a.S + (brev_s((uint32_t)r, a.s_log) << 8) + l;
        const size_t stride = (size_t)a.n_out * a.S;
#pragma unroll
        for (int b = 0; b < 16; b++) y[b] = src[16 * b];
        for (uint32_t sp = 1; sp < a.n_split; sp++) {
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const float2 u = src[sp * stride + 16 * b];
                y[b].x += u.x; y[b].y += u.y;
            }
        }
    }
    const float2 cr = r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f);
    r16_tables_init(tb, a.rowtw);  // after the partial-sum loads
    const R16Pair pr = r16_pair(r == 0, lane, x2_partner_half(wi, h));
    r16_split_inv<true>(y, pr, cr, tb, l);
    fft256x2<true>(y, reg, tb, l);
    __syncwarp();  // the exchange region now takes the row in natural order for the column phase
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const int n2 = l + 16 * p;
        reg[n2] = r ? cmulc(y[p], __ldg(&a.twM[n2 * r])) : y[p];
    }
    __syncthreads();
    // columns: z[256 n1 + n2] = sum_k1 A'[k1][n2] W_M1^(-n1 k1); keep the second half of the time samples
    // (overlap discard): they belong to output times [t_end*B - S + off, t_end*B + off)
    const uint32_t amask = a.acc_len - 1;
    const uint32_t pos0 = (uint32_t)((tend * (unsigned long long)a.B - a.S + a.off) & amask);
    float *acc = a.accring + (size_t)item * a.acc_len;
    for (int n2 = threadIdx.x; n2 < 256; n2 += M1 * 16) {
        float2 col[M1];
#pragma unroll
        for (int k1 = 0; k1 < M1; k1++) col[k1] = sm[k1 * kR16Slots + n2];
        dft_reg<M1, true>(col);
        float2 q[M1 / 2];
#pragma unroll
        for (int hh = 0; hh < M1 / 2; hh++) q[hh] = *reinterpret_cast<const float2 *>(acc + ((pos0 + 2u * (256u * hh + n2)) & amask));
#pragma unroll
        for (int hh = 0; hh < M1 / 2; hh++) {
            q[hh].x += col[M1 / 2 + hh].x; q[hh].y += col[M1 / 2 + hh].y;
            *reinterpret_cast<float2 *>(acc + ((pos0 + 2u * (256u * hh + n2)) & amask)) = q[hh];
        }
    }
}

// ---- M1 = 32, 64: the row launches (columns: k_tcols_fwd / k_tcols_inv of kernels_rows.cuh) ----------------
// rows, forward, in place in the delay-line slot: grid (voice * (M1/16) + row group, input, firing instance)
template <int M1>
__global__ void __launch_bounds__(kX2Threads, kX2MinB) k_trows_fwd_x2(const TierFwdArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];
    __shared__ R16Tables tb;
    const TierCommon c = tier_fwd_common(a, blockIdx.x / (M1 / 16), blockIdx.y, blockIdx.z);
    if (!c.active) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, l = lane & 15;
    const int pi = (int)(blockIdx.x % (M1 / 16)) * kX2Warps + warp;
    const int r = x2_row<M1>(pi, h);
    float2 *reg = sm + (2 * warp + h) * kR16Slots;
    float2 *slot = a.X + ((size_t)c.w * a.Lring + c.slot) * a.S + (brev_s((uint32_t)r, a.s_log) << 8) + l;
    float2 v[16];
#pragma unroll
    for (int b = 0; b < 16; b++) v[b] = slot[16 * b];
    const float2 cr = r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f);
    r16_tables_init(tb, a.rowtw);  // the row is in flight meanwhile
    if (r) {
#pragma unroll
        for (int b = 0; b < 16; b++) v[b] = cmul(v[b], __ldg(&a.twM[(l + 16 * b) * r]));
    }
    fft256x2<false>(v, reg, tb, l);
    const R16Pair pr = r16_pair(r == 0, lane, x2_partner_half(pi, h));
    r16_split_fwd<true>(v, pr, cr, tb, l, [&](int q, float2 X) { slot[16 * q] = X; });
}

// rows, inverse, in place in split 0 of the partial sums: grid (output * (M1/16) + row group, firing instance)
template <int M1>
__global__ void __launch_bounds__(kX2Threads, kX2MinB) k_trows_inv_x2(const TierInvArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];
    __shared__ R16Tables tb;
    const uint32_t z = blockIdx.y, o = blockIdx.x / (M1 / 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, l = lane & 15;
    const int pi = (int)(blockIdx.x % (M1 / 16)) * kX2Warps + warp;
    const int r = x2_row<M1>(pi, h);
    float2 *reg = sm + (2 * warp + h) * kR16Slots;
    float2 *y0 = const_cast<float2 *>(a.Ypart) + (((size_t)z * a.n_split) * a.n_out + o) * a.S + (brev_s((uint32_t)r, a.s_log) << 8) + l;  // engine-owned scratch
    float2 y[16];
    {
        const size_t stride = (size_t)a.n_out * a.S;
#pragma unroll
        for (int b = 0; b < 16; b++) y[b] = y0[16 * b];
        for (uint32_t sp = 1; sp < a.n_split; sp++) {
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const float2 u = y0[sp * stride + 16 * b];
                y[b].x += u.x; y[b].y += u.y;
            }
        }
    }
    const float2 cr = r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f);
    r16_tables_init(tb, a.rowtw);  // after the partial-sum loads
    const R16Pair pr = r16_pair(r == 0, lane, x2_partner_half(pi, h));
    r16_split_inv<true>(y, pr, cr, tb, l);
    fft256x2<true>(y, reg, tb, l);
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const int n2 = l + 16 * p;
        y0[16 * p] = r ? cmulc(y[p], __ldg(&a.twM[n2 * r])) : y[p];
    }
}

}  // namespace ca
