// kernels_rows.cuh -- the period's FFT kernels on the row-FFT family (fft_rows.cuh).
//
// Same work, same global layouts and same arguments as k_forward<8> / k_inverse<8, true> /
// k_tier_forward / k_tier_inverse in kernels.cuh (which stay as the general path for periods other than
// 256 and as the A/B reference, CA_FLAG_LEGACY_FFT); what changes is the transform:
//   k_fwd0_rows / k_inv0_rows        tier 0 at B = 256: one warp per (instance, input | output)
//   k_tfwd_fused / k_tinv_fused<M1>  long tier of M1 x 256 points, M1 <= 16: one CTA of M1 warps per
//                                    transform, M1-point column DFTs in registers, M1 row FFTs
//   k_tcols_* / k_trows_*<M1>        M1 = 32, 64: column DFTs and row FFTs as two launches of small CTAs,
//                                    the intermediate lives IN PLACE in the delay-line slot / partial sum
// Replaces cufftExecC2C + f_unpackC22R + f_pack2R2C (conv.cu:35-73, 367, 405-408) like the kernels it mirrors.
#pragma once
#include "fft_rows.cuh"

namespace ca {

constexpr int kRowsWarps = 8;
constexpr int kRowsThreads = kRowsWarps * 32;
constexpr uint32_t kRowsSmem = kRowsWarps * kRowSlots * sizeof(float2);  // 18 432 B

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t brev_s(uint32_t r, uint32_t s_log) { return s_log ? (__brev(r) >> (32 - s_log)) : 0u; }

// ------------------------------------------------------------------------------------------
// tier 0, B = 256: forward.  One warp per (instance, input); semantics identical to k_forward<8>.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsThreads, 4) k_fwd0_rows(const FwdArgs a)
{
    pdl_trigger();
    pdl_wait();
    constexpr int B = 256;
    extern __shared__ __align__(16) float2 sm[];
    __shared__ RowTables tb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * kRowsWarps + warp;
    const bool valid = w < a.n_items;
    const uint32_t item = a.item0 + (valid ? w : 0u);
    const unsigned long long t = a.t_host_p1 ? a.t_host_p1 - 1ull : a.ctl->t;
    // parameter / state loads are in flight while the twiddle tables are staged (one L2 round trip, not two)
    const InParamDev p = a.par[item];
    const uint32_t pd = a.par[(item / a.n_in) * a.n_in].predelay;  // input 0's, conv.cu:412,415
    const ItemState old = a.st[(t & 1ull) * a.n_items_alloc + item];
    rows_tables_init(tb, a.rowtw);
    if (!valid) return;
    if (w == 0 && lane == 0) a.ctl->t_next = t + 1ull;
    const ItemState s = item_step_warp_from(old, a.st, a.n_items_alloc, item, p, t, (int)a.nv, a.ring_out, a.vp, lane);
    // per-voice coefficient and storage entry wait in shared memory: keeping eight more values alive across the
    // transform spills at 64 registers, re-reading them from the stored state would be an L2 round trip per voice
    __shared__ float s_vc[kRowsWarps][kMaxVoices];
    __shared__ uint32_t s_ve[kRowsWarps][kMaxVoices];
    if (lane < kMaxVoices) {
        float cv = 0.f;
        uint32_t ev = 0u;
#pragma unroll
        for (int q = 0; q < kMaxVoices; q++) { cv = (q == lane) ? s.c[q] : cv; ev = (q == lane) ? s.pool[q] : ev; }
        s_vc[warp][lane] = cv;
        s_ve[warp][lane] = ev;
    }
    const uint32_t active = s.active, fresh = s.fresh;
    __syncwarp();

    float2 *row = sm + warp * kRowSlots;
    const uint32_t mask = a.ring_len - 1;
    const float *x = a.in + (size_t)item * B;
    const uint32_t base = (uint32_t)((t * (unsigned long long)B) & mask);
    const uint32_t prev = (base - B) & mask;
    const uint32_t slot = (a.Lring - 1u) - (uint32_t)((t + 1ull) % a.Lring);  // the FDL ring runs backwards

#pragma unroll 1
    for (uint32_t v = 0; v < a.nv; v++) {
        if (!((active >> v) & 1u)) continue;
        const float gain = s_vc[warp][v] * p.level;
        const uint32_t entry = s_ve[warp][v] - 1u;
        float *ring = a.ring + (size_t)entry * a.ring_len;
        if ((fresh >> v) & 1u) {  // (re)allocated voice: its time-domain history belongs to another IR
            for (uint32_t n = 4 * lane; n < a.ring_len; n += 128) *reinterpret_cast<float4 *>(ring + n) = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; j++)  // clear the block that becomes reachable by the predelay scatter in this period
            ring[(base + kMaxPredelay + lane + 32 * j) & mask] = 0.f;

        // window z[n] = (w[2n], w[2n+1]), w = [previous block | current block]; lane holds z[lane + 32 b]
        float2 z[8];
        if (pd == 0) {
            float2 xi[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                z[b] = *reinterpret_cast<const float2 *>(ring + prev + 2 * (lane + 32 * b));
                z[4 + b] = *reinterpret_cast<const float2 *>(ring + base + 2 * (lane + 32 * b));
                xi[b] = *reinterpret_cast<const float2 *>(x + 2 * (lane + 32 * b));
            }
#pragma unroll
            for (int b = 0; b < 4; b++) {  // what earlier, delayed blocks left in the current block + this period's input
                z[4 + b].x = fmaf(gain, xi[b].x, z[4 + b].x);
                z[4 + b].y = fmaf(gain, xi[b].y, z[4 + b].y);
                *reinterpret_cast<float2 *>(ring + base + 2 * (lane + 32 * b)) = z[4 + b];
            }
        } else {
            // predelay ring: the whole response of this block is delayed by pd samples (conv.cu:97)
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int n = lane + 32 * j;
                ring[(base + pd + n) & mask] += gain * __ldg(&x[n]);
            }
            __syncwarp();
#pragma unroll
            for (int b = 0; b < 4; b++) {
                z[b] = *reinterpret_cast<const float2 *>(ring + prev + 2 * (lane + 32 * b));
                z[4 + b] = *reinterpret_cast<const float2 *>(ring + base + 2 * (lane + 32 * b));
            }
        }
        row_fft256<false>(z, row, tb, lane, NoPostTw{});
        rows_split<false>(row, row, true, make_float2(1.f, 0.f), tb, lane);
        __syncwarp();
        row_store_global(row, a.X + ((size_t)entry * a.Lring + slot) * B, lane);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// tier 0, B = 256: inverse + overlap discard + output ring + clamp + dry mix.  One warp per
// (instance, output); semantics identical to k_inverse<8, true>.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsThreads, 4) k_inv0_rows(const InvArgs a)
{
    pdl_trigger();
    pdl_wait();
    constexpr int B = 256;
    extern __shared__ __align__(16) float2 sm[];
    __shared__ RowTables tb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t local0 = blockIdx.x * kRowsWarps + warp;
    const bool valid = local0 < a.n_items;
    const uint32_t local = valid ? local0 : 0u;   // idle warps of the last CTA shadow item 0's loads, then leave
    const uint32_t item = a.item0 + local;
    const uint32_t inst = item / a.n_out, o = item % a.n_out;
    const unsigned long long t = a.t_host_p1 ? a.t_host_p1 - 1ull : a.ctl->t_next - 1ull;
    float2 *row = sm + warp * kRowSlots;
    if (a.n_wait) {  // ca_group root: the peers' spectra arrive over NVLink (see InvArgs)
        group_wait(a.wait_flags, a.n_wait, t + 1ull, a.gerr, lane);
        __syncwarp();
    }

    // partial spectra of the MAC's row-range splits, fixed order
    {
        const float4 *src = reinterpret_cast<const float4 *>(a.Ypart + (((size_t)inst * a.n_split) * a.n_out + o) * B);
        const size_t stride4 = (size_t)a.n_out * B / 2;
        float4 q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = src[lane + 32 * i];
        for (uint32_t sp = 1; sp < a.n_split; sp++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 r = src[sp * stride4 + lane + 32 * i];
                q[i].x += r.x; q[i].y += r.y; q[i].z += r.z; q[i].w += r.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) *reinterpret_cast<float4 *>(row + e3(2 * lane + 64 * i)) = q[i];
    }
    // everything the epilogue needs is fetched BEFORE the transform.  Kept time samples [B, 2B) = z[128, 256):
    // lane holds z[128 + 4 lane, +4) = output samples [8 lane, 8 lane + 8)
    const int n0 = 8 * lane;
    const bool raw = a.raw_wet != 0;
    float dg[2] = {0.f, 0.f};
    if (!raw) {  // dry gain per input: dry * panDry * level, conv.cu:418-427
        const InParamDev p0 = a.par[inst * a.n_in];
        const InParamDev p1 = a.par[inst * a.n_in + (a.n_in - 1)];
        dg[0] = p0.dry * pan_gain(p0.panDry, (int)o, (int)a.n_out) * p0.level;
        dg[1] = a.n_in > 1 ? p1.dry * pan_gain(p1.panDry, (int)o, (int)a.n_out) * p1.level : 0.f;
    }
    const float *x0 = a.in + ((size_t)inst * a.n_in) * B + n0;
    const float *x1 = x0 + (a.n_in > 1 ? B : 0);
    float *accp = a.accring ? a.accring + (size_t)item * a.acc_len + (uint32_t)((t * (unsigned long long)B) & (a.acc_len - 1)) + n0 : nullptr;
    float4 xa[2], xb[2], accv[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        xa[j] = *reinterpret_cast<const float4 *>(x0 + 4 * j);
        xb[j] = *reinterpret_cast<const float4 *>(x1 + 4 * j);
        accv[j] = accp ? *reinterpret_cast<const float4 *>(accp + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rows_tables_init(tb, a.rowtw);  // after every global load of the prologue has been issued
    if (!valid) return;
    rows_split<true>(row, row, true, make_float2(1.f, 0.f), tb, lane);
    __syncwarp();
    float2 v[8];
#pragma unroll
    for (int b = 0; b < 8; b++) v[b] = row[e3(lane + 32 * b)];
    __syncwarp();
    row_fft256<true>(v, row, tb, lane, NoPostTw{});
    auto clampf = [raw](float w) { return raw ? w : fminf(fmaxf(w, -1.0f), 1.0f); };  // conv.cu:98
    float *dst = a.out + ((size_t)inst * a.n_out + o) * B + n0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const float4 zz = *reinterpret_cast<const float4 *>(row + e3(128 + 4 * lane + 2 * j));
        // wet = this period's tier-0 block + what the long (deferred) tiers left for it
        float4 y = make_float4(clampf(zz.x + accv[j].x), clampf(zz.y + accv[j].y), clampf(zz.z + accv[j].z), clampf(zz.w + accv[j].w));
        y.x = fmaf(dg[0], xa[j].x, fmaf(dg[1], xb[j].x, y.x));
        y.y = fmaf(dg[0], xa[j].y, fmaf(dg[1], xb[j].y, y.y));
        y.z = fmaf(dg[0], xa[j].z, fmaf(dg[1], xb[j].z, y.z));
        y.w = fmaf(dg[0], xa[j].w, fmaf(dg[1], xb[j].w, y.w));
        *reinterpret_cast<float4 *>(dst + 4 * j) = y;
        if (accp) *reinterpret_cast<float4 *>(accp + 4 * j) = make_float4(0.f, 0.f, 0.f, 0.f);  // consumed
    }
    if (a.advance && local == 0 && lane == 0) { a.ctl->t = t + 1ull; a.ctl->t_def[(t + 1ull) & 1ull] = t + 1ull; }
}

// ------------------------------------------------------------------------------------------
// long tiers, M = S = M1 x 256 complex points of a 2S-sample real window
// ------------------------------------------------------------------------------------------
struct TierCommon {
    uint32_t inst, item, w, slot;
    unsigned long long tend;
    bool active;
};

// firing instance / voice state / delay-line slot of a forward tier CTA (voice v, input i, firing instance z)
__device__ __forceinline__ TierCommon tier_fwd_common(const TierFwdArgs &a, uint32_t v, uint32_t i, uint32_t z)
{
    TierCommon c;
    c.inst = a.inst0 + z * a.inst_stride;
    c.item = c.inst * a.n_in + i;
    c.tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, 0u);
    const ItemState &st = a.st[(c.tend & 1ull) * a.n_items_alloc + c.item];
    c.active = ((st.active >> v) & 1u) != 0;
    c.w = st.pool[v] - 1u;  // voice pool entry (meaningful when active)
    const unsigned long long n_fire = (c.tend + (c.inst & (a.m - 1u))) >> (31 - __clz((int)a.m));  // m is a power of two
    c.slot = (a.Lring - 1u) - (uint32_t)(n_fire % a.Lring);
    return c;
}

// One CTA of max(M1, 1) warps per (voice, input, firing instance): window -> column DFTs -> rows -> split -> FDL slot
template <int M1>
__global__ void __launch_bounds__(M1 * 32) k_tfwd_fused(const TierFwdArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];  // [M1][kRowSlots]
    __shared__ RowTables tb;
    const TierCommon c = tier_fwd_common(a, blockIdx.x, blockIdx.y, blockIdx.z);
    if (!c.active) return;  // uniform for the CTA
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const uint32_t mask = a.ring_len - 1;
    const float *ring = a.ring + (size_t)c.w * a.ring_len;
    const uint32_t start = (uint32_t)((c.tend * (unsigned long long)a.B - 2ull * a.S) & mask);
    // columns: A[k1][n2] = sum_n1 z[256 n1 + n2] W_M1^(n1 k1)
    for (int n2 = threadIdx.x; n2 < 256; n2 += M1 * 32) {
        float2 col[M1];
#pragma unroll
        for (int n1 = 0; n1 < M1; n1++) col[n1] = *reinterpret_cast<const float2 *>(ring + ((start + 2u * (256u * n1 + n2)) & mask));
        dft_reg<M1, false>(col);
#pragma unroll
        for (int k1 = 0; k1 < M1; k1++) sm[k1 * kRowSlots + e3(n2)] = col[k1];
    }
    rows_tables_init(tb, a.rowtw);  // staged after the window loads were issued; its barrier also closes the column phase
    // rows: X[k1 + M1 k2] = sum_n2 A[k1][n2] W_M^(n2 k1) W_256^(n2 k2)
    float2 *row = sm + r * kRowSlots;
    float2 v[8];
#pragma unroll
    for (int b = 0; b < 8; b++) v[b] = row[e3(lane + 32 * b)];
    if (r) {
#pragma unroll
        for (int b = 0; b < 8; b++) v[b] = cmul(v[b], __ldg(&a.twM[(lane + 32 * b) * r]));
    }
    __syncwarp();
    row_fft256<false>(v, row, tb, lane, NoPostTw{});
    __syncthreads();
    const int rp = (M1 - r) % M1;
    rows_split<false>(row, sm + rp * kRowSlots, r == 0, r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f), tb, lane);
    __syncthreads();
    float2 *dst = a.X + ((size_t)c.w * a.Lring + c.slot) * a.S;
    row_store_global(row, dst + (brev_s((uint32_t)r, a.s_log) << 8), lane);
}

// One CTA per (output, firing instance): partial sums -> split -> rows -> column DFTs -> += output ring at +off
template <int M1>
__global__ void __launch_bounds__(M1 * 32) k_tinv_fused(const TierInvArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];
    __shared__ RowTables tb;
    const uint32_t z = blockIdx.y, o = blockIdx.x;
    const uint32_t inst = a.inst0 + z * a.inst_stride;
    const uint32_t item = inst * a.n_out + o;
    const unsigned long long tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, 0u);
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    float2 *row = sm + r * kRowSlots;
    {
        const float4 *src = reinterpret_cast<const float4 *>(a.Ypart + (((size_t)z * a.n_split) * a.n_out + o) * a.S + (brev_s((uint32_t)r, a.s_log) << 8));
        const size_t stride4 = (size_t)a.n_out * a.S / 2;
        float4 q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = src[lane + 32 * i];
        for (uint32_t sp = 1; sp < a.n_split; sp++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 u = src[sp * stride4 + lane + 32 * i];
                q[i].x += u.x; q[i].y += u.y; q[i].z += u.z; q[i].w += u.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) *reinterpret_cast<float4 *>(row + e3(2 * lane + 64 * i)) = q[i];
    }
    rows_tables_init(tb, a.rowtw);  // after the partial-sum loads; its barrier also publishes the rows
    const int rp = (M1 - r) % M1;
    rows_split<true>(row, sm + rp * kRowSlots, r == 0, r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f), tb, lane);
    __syncthreads();
    float2 v[8];
#pragma unroll
    for (int b = 0; b < 8; b++) v[b] = row[e3(lane + 32 * b)];
    __syncwarp();
    const float2 *twM = a.twM;
    auto post = [twM, lane, r](int j, float2 val) { return r ? cmulc(val, __ldg(&twM[(lane + 32 * j) * r])) : val; };
    row_fft256<true>(v, row, tb, lane, post);
    __syncthreads();
    // columns: z[256 n1 + n2] = sum_k1 A'[k1][n2] W_M1^(-n1 k1); keep the second half of the time samples
    // (overlap discard): they belong to output times [t_end*B - S + off, t_end*B + off)
    const uint32_t amask = a.acc_len - 1;
    const uint32_t pos0 = (uint32_t)((tend * (unsigned long long)a.B - a.S + a.off) & amask);
    float *acc = a.accring + (size_t)item * a.acc_len;
    for (int n2 = threadIdx.x; n2 < 256; n2 += M1 * 32) {
        float2 col[M1];
#pragma unroll
        for (int k1 = 0; k1 < M1; k1++) col[k1] = sm[k1 * kRowSlots + e3(n2)];
        dft_reg<M1, true>(col);
        if constexpr (M1 == 1) {
            if (n2 >= 128) {
                float2 *p = reinterpret_cast<float2 *>(acc + ((pos0 + 2u * (n2 - 128)) & amask));
                float2 q = *p;
                q.x += col[0].x; q.y += col[0].y;
                *p = q;
            }
        } else {
            float2 q[M1 / 2];
#pragma unroll
            for (int h = 0; h < M1 / 2; h++) q[h] = *reinterpret_cast<const float2 *>(acc + ((pos0 + 2u * (256u * h + n2)) & amask));
#pragma unroll
            for (int h = 0; h < M1 / 2; h++) {
                q[h].x += col[M1 / 2 + h].x; q[h].y += col[M1 / 2 + h].y;
                *reinterpret_cast<float2 *>(acc + ((pos0 + 2u * (256u * h + n2)) & amask)) = q[h];
            }
        }
    }
}

// ---- M1 = 32, 64: columns and rows as separate launches --------------------------------------------
// row handled by warp `warp` of row-group CTA `rg`: pair p = 4 rg + warp / 2 is (0, M1/2) for p = 0 (both
// self-paired) and (p, M1 - p) otherwise
template <int M1>
__device__ __forceinline__ int rows_pair_row(int rg, int warp)
{
    const int p = 4 * rg + (warp >> 1), second = warp & 1;
    return p == 0 ? (second ? M1 / 2 : 0) : (second ? M1 - p : p);
}

// columns, forward: grid (voice * 8 + column group, input, firing instance), 256 threads = 32 columns x 8
template <int M1>
__global__ void __launch_bounds__(256, 4) k_tcols_fwd(const TierFwdArgs a)
{
    pdl_trigger();
    pdl_wait();
    constexpr int Mb = M1 / 8;
    __shared__ float2 S[Mb * 8 * 32];
    if ((blockIdx.x >> 3) == 0 && a.tend_host && threadIdx.x < 2 * M1) {
        // the window sits behind two dependent round trips (voice state -> storage entry -> ring): pull this CTA's 32
        // columns of the item's HOME entry (where its voice lives outside cross-fades) into L2 meanwhile
        const uint32_t item = (a.inst0 + blockIdx.z * a.inst_stride) * a.n_in + blockIdx.y, mask0 = a.ring_len - 1;
        const uint32_t start0 = (uint32_t)((a.tend_host * (unsigned long long)a.B - 2ull * a.S) & mask0);
        const uint32_t n1 = threadIdx.x >> 1, half = threadIdx.x & 1;   // 64 floats (two lines) per row of the column block
        prefetch_l2(a.ring + (size_t)item * a.ring_len + ((start0 + 2u * (256u * n1 + 32u * (blockIdx.x & 7)) + 32u * half) & mask0));
    }
    const TierCommon c = tier_fwd_common(a, blockIdx.x >> 3, blockIdx.y, blockIdx.z);
    if (!c.active) return;
    const int cc = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int n2 = (blockIdx.x & 7) * 32 + cc;
    const uint32_t mask = a.ring_len - 1;
    const float *ring = a.ring + (size_t)c.w * a.ring_len;
    const uint32_t start = (uint32_t)((c.tend * (unsigned long long)a.B - 2ull * a.S) & mask);
    // n1 = j + 8 b: radix Mb over b, twiddle W_M1^(j q)
    float2 y[Mb];
#pragma unroll
    for (int b = 0; b < Mb; b++) y[b] = *reinterpret_cast<const float2 *>(ring + ((start + 2u * (256u * (j + 8 * b) + n2)) & mask));
    dft_reg<Mb, false>(y);
#pragma unroll
    for (int q = 0; q < Mb; q++) S[(q * 8 + j) * 32 + cc] = q ? cmul(y[q], __ldg(&a.twM[256 * j * q])) : y[0];  // W_M1^(jq) = W_M^(256 jq)
    __syncthreads();
    if (j < Mb) {
        const int q = j;
        float2 u[8];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) u[jj] = S[(q * 8 + jj) * 32 + cc];
        dft_reg<8, false>(u);
        float2 *dst = a.X + ((size_t)c.w * a.Lring + c.slot) * a.S;
#pragma unroll
        for (int rr = 0; rr < 8; rr++) dst[(brev_s((uint32_t)(q + Mb * rr), a.s_log) << 8) + n2] = u[rr];  // k1 = q + Mb r
    }
}

// rows, forward, in place in the delay-line slot: grid (voice * (M1/8) + row group, input, firing instance)
template <int M1>
__global__ void __launch_bounds__(kRowsThreads, 4) k_trows_fwd(const TierFwdArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];
    __shared__ RowTables tb;
    const TierCommon c = tier_fwd_common(a, blockIdx.x / (M1 / 8), blockIdx.y, blockIdx.z);
    if (!c.active) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = rows_pair_row<M1>((int)(blockIdx.x % (M1 / 8)), warp);
    float2 *row = sm + warp * kRowSlots;
    float2 *slot = a.X + ((size_t)c.w * a.Lring + c.slot) * a.S + (brev_s((uint32_t)r, a.s_log) << 8);
    float2 v[8], tw[8];
#pragma unroll
    for (int b = 0; b < 8; b++) { v[b] = slot[lane + 32 * b]; tw[b] = __ldg(&a.twM[(lane + 32 * b) * r]); }
    rows_tables_init(tb, a.rowtw);  // the row and its four-step twiddles are in flight meanwhile
    if (r) {
#pragma unroll
        for (int b = 0; b < 8; b++) v[b] = cmul(v[b], tw[b]);
    }
    row_fft256<false>(v, row, tb, lane, NoPostTw{});
    __syncthreads();
    const int rp = (M1 - r) % M1;
    rows_split<false>(row, rp == r ? row : sm + (warp ^ 1) * kRowSlots, r == 0, r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f), tb, lane);
    __syncthreads();
    row_store_global(row, slot, lane);
}

// rows, inverse, in place in split 0 of the partial sums: grid (output * (M1/8) + row group, firing instance)
template <int M1>
__global__ void __launch_bounds__(kRowsThreads, 4) k_trows_inv(const TierInvArgs a)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float2 sm[];
    __shared__ RowTables tb;
    const uint32_t z = blockIdx.y, o = blockIdx.x / (M1 / 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = rows_pair_row<M1>((int)(blockIdx.x % (M1 / 8)), warp);
    float2 *row = sm + warp * kRowSlots;
    float2 *y0 = const_cast<float2 *>(a.Ypart) + (((size_t)z * a.n_split) * a.n_out + o) * a.S + (brev_s((uint32_t)r, a.s_log) << 8);  // engine-owned scratch
    {
        const float4 *src = reinterpret_cast<const float4 *>(y0);
        const size_t stride4 = (size_t)a.n_out * a.S / 2;
        float4 q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = src[lane + 32 * i];
        for (uint32_t sp = 1; sp < a.n_split; sp++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 u = src[sp * stride4 + lane + 32 * i];
                q[i].x += u.x; q[i].y += u.y; q[i].z += u.z; q[i].w += u.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) *reinterpret_cast<float4 *>(row + e3(2 * lane + 64 * i)) = q[i];
    }
    rows_tables_init(tb, a.rowtw);  // after the partial-sum loads; its barrier also publishes the rows
    const int rp = (M1 - r) % M1;
    rows_split<true>(row, rp == r ? row : sm + (warp ^ 1) * kRowSlots, r == 0, r ? __ldg(&a.tw2M[r]) : make_float2(1.f, 0.f), tb, lane);
    __syncthreads();
    float2 v[8];
#pragma unroll
    for (int b = 0; b < 8; b++) v[b] = row[e3(lane + 32 * b)];
    __syncwarp();
    const float2 *twM = a.twM;
    auto post = [twM, lane, r](int j, float2 val) { return r ? cmulc(val, __ldg(&twM[(lane + 32 * j) * r])) : val; };
    row_fft256<true>(v, row, tb, lane, post);
    row_store_global(row, y0, lane);
}

// columns, inverse: grid (output * 8 + column group, firing instance); reads split 0 of the partial sums
template <int M1>
__global__ void __launch_bounds__(256, 4) k_tcols_inv(const TierInvArgs a)
{
    pdl_trigger();
    pdl_wait();
    constexpr int Mb = M1 / 8;
    __shared__ float2 S[Mb * 8 * 32];
    const uint32_t z = blockIdx.y, o = blockIdx.x >> 3;
    const uint32_t inst = a.inst0 + z * a.inst_stride;
    const uint32_t item = inst * a.n_out + o;
    const unsigned long long tend = ctl_tend(a.ctl, a.tend_host, a.t_sel, 0u);
    const int cc = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int n2 = (blockIdx.x & 7) * 32 + cc;
    if (threadIdx.x < M1) {  // the output-ring samples these 32 columns add into (M1 / 2 kept rows of 64 floats): into L2 now
        const uint32_t amask0 = a.acc_len - 1;
        const uint32_t p0 = (uint32_t)((tend * (unsigned long long)a.B - a.S + a.off) & amask0);
        const uint32_t hrow = threadIdx.x >> 1, half = threadIdx.x & 1;
        prefetch_l2(a.accring + (size_t)item * a.acc_len + ((p0 + 2u * (256u * hrow + 32u * (blockIdx.x & 7)) + 32u * half) & amask0));
    }
    const float2 *src = a.Ypart + (((size_t)z * a.n_split) * a.n_out + o) * a.S;
    // k1 = j + 8 b: radix Mb over b, twiddle conj W_M1^(j q)
    float2 y[Mb];
#pragma unroll
    for (int b = 0; b < Mb; b++) y[b] = src[(brev_s((uint32_t)(j + 8 * b), a.s_log) << 8) + n2];
    dft_reg<Mb, true>(y);
#pragma unroll
    for (int q = 0; q < Mb; q++) S[(q * 8 + j) * 32 + cc] = q ? cmulc(y[q], __ldg(&a.twM[256 * j * q])) : y[0];
    __syncthreads();
    if (j < Mb) {
        const int q = j;
        float2 u[8];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) u[jj] = S[(q * 8 + jj) * 32 + cc];
        dft_reg<8, true>(u);
        // n1 = q + Mb r; the kept half n1 >= M1/2 is r >= 4
        const uint32_t amask = a.acc_len - 1;
        const uint32_t pos0 = (uint32_t)((tend * (unsigned long long)a.B - a.S + a.off) & amask);
        float *acc = a.accring + (size_t)item * a.acc_len;
        float2 old[4];
#pragma unroll
        for (int rr = 0; rr < 4; rr++) old[rr] = *reinterpret_cast<const float2 *>(acc + ((pos0 + 2u * (256u * (q + Mb * rr)) + 2u * n2) & amask));
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            old[rr].x += u[4 + rr].x; old[rr].y += u[4 + rr].y;
            *reinterpret_cast<float2 *>(acc + ((pos0 + 2u * (256u * (q + Mb * rr)) + 2u * n2) & amask)) = old[rr];
        }
    }
}

}  // namespace ca
