// fft_rows16.cuh -- 256-point complex row FFT as 16 x 16: TWO rows per warp (one per half-warp), 16 points
// per lane, for sm_100a.
//
// Why a third layout beside fft_rows.cuh (256 = 8 x 8 x 4, one row per warp): ncu of the round-2 row kernels
// showed them latency-bound on SHARED memory as much as on global memory (short-scoreboard stalls = long-
// scoreboard stalls, IPC 1.05-1.45 of 4): a row went through three shared-memory exchanges plus a fourth round
// trip for the real-FFT split, ~584 warp instructions of which only 286 are floating point.  Here
//   * a row is two radix-16 register DFTs (compile-time twiddles, fft_warp.cuh) around ONE exchange through a
//     2.3 KB shared-memory region private to the half-warp (16 x 8-byte stores, 8 x 16-byte loads per lane, all
//     bank-conflict free: [q][l] with a row pitch of 18 float2),
//   * the real-FFT split pairs bin k with bin M - k through register SHUFFLES: after the transform lane l holds
//     bins l + 16 p, the partner of (l, p) is (16 - l, 15 - p) in the same row (row 0 of a transform) or
//     (15 - l, 15 - p) in the partner row, which the OTHER half-warp of the same warp holds,
//   * rows enter and leave through 8-byte global accesses of 128 contiguous bytes per half-warp.
// 388 warp instructions per row instead of 584, one shared-memory round trip instead of four; measured on a bare
// load -> FFT -> split -> store kernel over 32 768 rows: 23.7 us against 35.4 us (HBM floor of that traffic: 21 us).
// Global layouts are those of fft_rows.cuh / fft_cta.cuh (position order), so the families stay interchangeable.
// Index math modelled in numpy (tests/_rows16_fft_model.py, tests/test_fft_model.py).
#pragma once
#include "fft_rows.cuh"

namespace ca {

constexpr int kR16Pitch = 18;               // float2 per exchange row: 8-byte writes [q][l] and 16-byte reads [l][2i] conflict-free
constexpr int kR16Slots = 16 * kR16Pitch;   // 288 float2 per row region (>= 256: a region also holds a row in natural order)
static_assert(kR16Slots == kRowSlots, "row regions of both families have the same size");

struct R16Tables {
    float2 w16[256];   // [q * 16 + l] = W_256^(l q): the twiddle between the two radix-16 stages, lane-contiguous
    float2 w512[256];  // W_512^k   (real-FFT split of a row)
};

// g: [W_256^n | W_512^k | W_256^(l q) at q * 16 + l] in global memory; all threads; ends with __syncthreads()
__device__ __forceinline__ void r16_tables_init(R16Tables &t, const float2 *__restrict__ g)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        t.w16[i] = __ldg(&g[512 + i]);
        t.w512[i] = __ldg(&g[256 + i]);
    }
    __syncthreads();
}

// In : v[b] = x[l + 16 b]   (l = lane & 15; the two half-warps hold two independent rows)
// Out: v[p] = X[l + 16 p]   unnormalised forward / inverse DFT
// S: this HALF-warp's exchange region (kR16Slots float2).  All 32 lanes must call.  The region may be reused by
// the caller after a __syncwarp().
template <bool INV>
__device__ __forceinline__ void fft256x2(float2 (&v)[16], float2 *S, const R16Tables &tb, int l)
{
    // n = l + 16 b, k = q + 16 p:  A[l][q] = sum_b x[l + 16 b] W_16^(b q);  X[q + 16 p] = sum_l A[l][q] W_256^(l q) W_16^(l p)
    dft_reg<16, INV>(v);
#pragma unroll
    for (int q = 0; q < 16; q++) S[q * kR16Pitch + l] = q ? tmul<INV>(v[q], tb.w16[q * 16 + l]) : v[0];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) {  // lane l now plays q: it gathers A[0..15][l]
        const float4 t = *reinterpret_cast<const float4 *>(S + l * kR16Pitch + 2 * i);
        v[2 * i] = make_float2(t.x, t.y);
        v[2 * i + 1] = make_float2(t.z, t.w);
    }
    dft_reg<16, INV>(v);
}

// ---- real-FFT split through shuffles ----------------------------------------------------------------
// After fft256x2 lane (h, l) register p holds bin k2 = l + 16 p of its row.  Row r (k1 = r) of an M = M1 x 256
// transform pairs with row (M1 - r) % M1: bin (r, k2) with (M1 - r, 255 - k2), or (0, 256 - k2) inside row 0.
//   kind A (row 0, and the only row of a 256-point transform): partner = lane ((16 - l) & 15) of the SAME half,
//           register 15 - p; lane 0 pairs inside itself: register (16 - p) & 15
//   kind B: partner = lane 15 - l of half `ph` (the other half-warp; the own one for the self-paired row M1 / 2),
//           register 15 - p
struct R16Pair {
    int src;     // source lane of the shuffles
    bool own;    // kind A, l == 0: the partner is one of the lane's own registers
    bool dcny;   // this lane's register 0 is position 0 = (DC, Nyquist)
};
__device__ __forceinline__ R16Pair r16_pair(bool kindA, int lane, int partner_half)
{
    const int l = lane & 15;
    R16Pair p;
    p.src = kindA ? ((lane & 16) | ((16 - l) & 15)) : ((partner_half << 4) | (15 - l));
    p.own = kindA && l == 0;
    p.dcny = p.own;
    return p;
}

// p must be a compile-time constant after unrolling (register indices)
__device__ __forceinline__ float2 r16_partner(const float2 (&v)[16], const R16Pair &pr, int p)
{
    float2 q;
    q.x = __shfl_sync(kFull, v[15 - p].x, pr.src);
    q.y = __shfl_sync(kFull, v[15 - p].y, pr.src);
    return pr.own ? v[(16 - p) & 15] : q;
}

// forward: Z = FFT_M(z) -> packed real-FFT bins; each result is handed to out(p, X[k2 = l + 16 p]).
// HAS_CR: W_2M^(k1 + M1 k2) = cr * W_512^k2 with cr = W_2M^k1 (long tiers); else the row is a whole 256-point transform.
template <bool HAS_CR, class Out>
__device__ __forceinline__ void r16_split_fwd(const float2 (&v)[16], const R16Pair &pr, float2 cr, const R16Tables &tb, int l, const Out &out)
{
    float2 part[16];
#pragma unroll
    for (int p = 0; p < 16; p++) part[p] = r16_partner(v, pr, p);
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const float2 t = tb.w512[l + 16 * p];
        const float2 w = HAS_CR ? cmul(cr, t) : t;
        float2 x = r2c_bin(v[p], part[p], w);
        if (p == 0 && pr.dcny) x = make_float2(v[0].x + v[0].y, v[0].x - v[0].y);
        out(p, x);
    }
}

// inverse, in place: packed bins Y -> Z with IFFT_M(Z)[n] = 2 M (y[2n] + j y[2n+1])
template <bool HAS_CR>
__device__ __forceinline__ void r16_split_inv(float2 (&v)[16], const R16Pair &pr, float2 cr, const R16Tables &tb, int l)
{
    float2 part[16];
#pragma unroll
    for (int p = 0; p < 16; p++) part[p] = r16_partner(v, pr, p);
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const float2 t = tb.w512[l + 16 * p];
        const float2 w = HAS_CR ? cmul(cr, t) : t;
        float2 z = c2r_bin(v[p], part[p], w);
        if (p == 0 && pr.dcny) z = make_float2(v[0].x + v[0].y, v[0].x - v[0].y);
        v[p] = z;
    }
}

}  // namespace ca
