// fft_cta.cuh -- CTA-level complex FFT of M = 256 * 2^s points (s = 0..6) in shared memory, for the
// long tiers of the non-uniform partitioning (block sizes 256 .. 16384).
//
//   forward : radix-4 DIF passes in shared memory (two fused radix-2 stages each, plus one radix-2
//             pass when s is odd) bring the span from M down to 512, then one 256-point warp FFT
//             (fft_warp.cuh, R = 8) per contiguous 256-element block.
//   inverse : the transposed flow graph (warp inverse FFTs, then DIT passes).
// After `forward` bin k lives at position zpos(k) = (bitrev_s(k mod 2^s) << 8) | (k >> s); `inverse`
// expects its input in that order and leaves time samples in natural order.  The long tiers keep
// their spectra in this POSITION order in global memory as well (FDL, IR bank, partial sums): the
// MAC is element-wise over bins, so any consistent order works, position 0 is still bin 0 (DC /
// Nyquist), and no bit-reversed gather is ever needed.
// Shared-memory indices go through swz(): an XOR swizzle that makes the warp FFT's 64-byte-strided
// float4 accesses conflict-free while leaving unit-stride accesses conflict-free.
// Twiddles come from two 128-entry shared-memory tables per transform (W^n = hi[n >> 7] * lo[n & 127],
// filled from the fp64-accurate global table): no dependent global loads inside the passes.
// Layout checked against numpy in tests/test_fft_model.py (cta_fwd / cta_inv / zpos).
#pragma once
#include "fft_warp.cuh"

namespace ca {

__device__ __forceinline__ int zpos(int k, int s)
{
    if (s == 0) return k;
    const int low = k & ((1 << s) - 1);
    const int b = (int)(__brev((unsigned)low) >> (32 - s));
    return (b << 8) | (k >> s);
}

// float2 index -> swizzled float2 index (bits 2:1 ^= bits 5:4; pairs (2i, 2i+1) stay adjacent)
__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) & 3) << 1); }

struct CtaTw {
    float2 hi[128], lo[128];    // W_M^(128 a), W_M^b
    float2 hi2[128], lo2[128];  // W_2M^(128 a), W_2M^b   (real-FFT split)
};

// all threads; twM = W_M^n (n < M), tw2M = W_2M^k (k < M); ends with __syncthreads()
__device__ __forceinline__ void cta_tw_init(CtaTw &t, int M, const float2 *__restrict__ twM, const float2 *__restrict__ tw2M)
{
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        const int a = (128 * i < M) ? 128 * i : 0;
        t.hi[i] = __ldg(&twM[a]);
        t.lo[i] = __ldg(&twM[i]);
        t.hi2[i] = __ldg(&tw2M[a]);
        t.lo2[i] = __ldg(&tw2M[i]);
    }
    __syncthreads();
}
__device__ __forceinline__ float2 tw_m(const CtaTw &t, int n) { return cmul(t.hi[n >> 7], t.lo[n & 127]); }
__device__ __forceinline__ float2 tw_2m(const CtaTw &t, int k) { return cmul(t.hi2[k >> 7], t.lo2[k & 127]); }

__device__ __forceinline__ float2 mul_mj(float2 v) { return make_float2(v.y, -v.x); }  // v * (-j)
__device__ __forceinline__ float2 mul_pj(float2 v) { return make_float2(-v.y, v.x); }  // v * (+j)

// two fused radix-2 DIF stages (st, st+1): span N = M >> st, quarter q = N / 4
__device__ __forceinline__ void dif_radix4_pass(float2 *sm, int M, int st, const CtaTw &t)
{
    const int lq = 31 - __clz(M) - st - 2;
    const int q = 1 << lq;
#pragma unroll 4
    for (int j = threadIdx.x; j < M / 4; j += blockDim.x) {
        const int pos = j & (q - 1);
        const int base = ((j >> lq) << (lq + 2)) | pos;
        const int i0 = swz(base), i1 = swz(base + q), i2 = swz(base + 2 * q), i3 = swz(base + 3 * q);
        const float2 a0 = sm[i0], a1 = sm[i1], a2 = sm[i2], a3 = sm[i3];
        const float2 wA = tw_m(t, pos << st);  // W_N^pos ; W_N^(pos+q) = -j W_N^pos ; W_(N/2)^pos = (W_N^pos)^2
        const float2 wB = cmul(wA, wA);
        const float2 b0 = cadd(a0, a2), b2 = cmul(csub(a0, a2), wA);
        const float2 b1 = cadd(a1, a3), b3 = mul_mj(cmul(csub(a1, a3), wA));
        sm[i0] = cadd(b0, b1);
        sm[i1] = cmul(csub(b0, b1), wB);
        sm[i2] = cadd(b2, b3);
        sm[i3] = cmul(csub(b2, b3), wB);
    }
    __syncthreads();
}

__device__ __forceinline__ void dit_radix4_pass(float2 *sm, int M, int st, const CtaTw &t)
{
    const int lq = 31 - __clz(M) - st - 2;
    const int q = 1 << lq;
#pragma unroll 4
    for (int j = threadIdx.x; j < M / 4; j += blockDim.x) {
        const int pos = j & (q - 1);
        const int base = ((j >> lq) << (lq + 2)) | pos;
        const int i0 = swz(base), i1 = swz(base + q), i2 = swz(base + 2 * q), i3 = swz(base + 3 * q);
        const float2 c0 = sm[i0], c1 = sm[i1], c2 = sm[i2], c3 = sm[i3];
        const float2 wA = tw_m(t, pos << st);
        const float2 wB = cmul(wA, wA);
        const float2 t1 = cmulc(c1, wB), t3 = cmulc(c3, wB);
        const float2 b0 = cadd(c0, t1), b1 = csub(c0, t1), b2 = cadd(c2, t3), b3 = csub(c2, t3);
        const float2 u2 = cmulc(b2, wA), u3 = mul_pj(cmulc(b3, wA));
        sm[i0] = cadd(b0, u2);
        sm[i2] = csub(b0, u2);
        sm[i1] = cadd(b1, u3);
        sm[i3] = csub(b1, u3);
    }
    __syncthreads();
}

__device__ __forceinline__ void radix2_pass(float2 *sm, int M, int st, const CtaTw &t, bool inverse)
{
    const int lh = 31 - __clz(M) - 1 - st;
    const int half = 1 << lh;
#pragma unroll 4
    for (int j = threadIdx.x; j < M / 2; j += blockDim.x) {
        const int pos = j & (half - 1);
        const int n0 = ((j >> lh) << (lh + 1)) | pos;
        const int i0 = swz(n0), i1 = swz(n0 + half);
        const float2 w = tw_m(t, pos << st);  // W_{2 half}^pos
        if (!inverse) {
            const float2 a = sm[i0], b = sm[i1];
            sm[i0] = cadd(a, b);
            sm[i1] = cmul(csub(a, b), w);
        } else {
            const float2 a = sm[i0], b = cmulc(sm[i1], w);
            sm[i0] = cadd(a, b);
            sm[i1] = csub(a, b);
        }
    }
    __syncthreads();
}

// twM: W_M^n, n < M.  All threads of the CTA must call; ends with __syncthreads().
__device__ __forceinline__ void cta_fft_forward(float2 *sm, int M, int s, const CtaTw &t, const float2 *__restrict__ twM)
{
    int st = 0;
    for (; s - st >= 2; st += 2) dif_radix4_pass(sm, M, st, t);
    if (s - st == 1) radix2_pass(sm, M, st, t, false);
    WarpFft<8> f;
    f.init(twM, M >> 8);
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int blk = warp; blk < (1 << s); blk += nwarps) {
        float2 *base = sm + (blk << 8);
        float2 v[8];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const float4 q = *reinterpret_cast<const float4 *>(base + swz(8 * f.lane + 2 * b));
            v[2 * b] = make_float2(q.x, q.y);
            v[2 * b + 1] = make_float2(q.z, q.w);
        }
        f.forward(v);
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 8; d++) base[swz(f.c + 32 * d)] = v[d];
    }
    __syncthreads();
}

__device__ __forceinline__ void cta_fft_inverse(float2 *sm, int M, int s, const CtaTw &t, const float2 *__restrict__ twM)
{
    WarpFft<8> f;
    f.init(twM, M >> 8);
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int blk = warp; blk < (1 << s); blk += nwarps) {
        float2 *base = sm + (blk << 8);
        float2 v[8];
#pragma unroll
        for (int d = 0; d < 8; d++) v[d] = base[swz(f.c + 32 * d)];
        f.inverse(v);
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 4; b++)
            *reinterpret_cast<float4 *>(base + swz(8 * f.lane + 2 * b)) = make_float4(v[2 * b].x, v[2 * b].y, v[2 * b + 1].x, v[2 * b + 1].y);
    }
    __syncthreads();
    int st = s;
    if (s & 1) { st = s - 1; radix2_pass(sm, M, st, t, true); }
    for (st -= 2; st >= 0; st -= 2) dit_radix4_pass(sm, M, st, t);
}

// real-FFT split of one bin (see WarpFft::split_r2c): z = Z[k], zp = Z[M-k], w = W_2M^k
__device__ __forceinline__ float2 r2c_bin(float2 z, float2 zp, float2 w)
{
    const float2 e = make_float2(0.5f * (z.x + zp.x), 0.5f * (z.y - zp.y));
    const float2 m = cmul(make_float2(z.x - zp.x, z.y + zp.y), w);
    return make_float2(e.x + 0.5f * m.y, e.y - 0.5f * m.x);
}
__device__ __forceinline__ float2 c2r_bin(float2 y, float2 yp, float2 w)
{
    const float2 s = make_float2(y.x + yp.x, y.y - yp.y);
    const float2 m = cmulc(make_float2(y.x - yp.x, y.y + yp.y), w);
    return make_float2(s.x - m.y, s.y + m.x);
}

// The split pairs bin k with bin M - k.  In position order, k = (q << s) | r sits in block
// b = bitrev_s(r) at offset q, its partner in block b' = bitrev_s((2^s - r) mod 2^s) at offset
// 255 - q (r != 0) or 256 - q (r == 0).  Iterating q fastest makes BOTH accesses unit-stride in shared
// memory (one ascending, one descending): no bank conflicts.  Every pair is visited once.
template <bool INVERSE>
__device__ __forceinline__ void cta_split(float2 *sm, int M, int s, const CtaTw &t)
{
    const int nblk = 1 << s;
#pragma unroll 2
    for (int idx = threadIdx.x; idx < M / 2 + 128; idx += blockDim.x) {
        // enumerate (r, q): r = 0 owns q in [0, 128]; every r in [1, nblk/2) owns q in [0, 256);
        // r = nblk/2 (self-paired block) owns q in [0, 128)
        int r, q;
        if (idx <= 128) { r = 0; q = idx; }
        else {
            const int rest = idx - 129;
            r = 1 + (rest >> 8); q = rest & 255;
            if (r > nblk / 2 || (r == nblk / 2 && q >= 128)) continue;
            if (nblk == 1) continue;
        }
        const int k = (q << s) | r;
        if (k == 0) {
            const float2 z = sm[0];
            sm[0] = make_float2(z.x + z.y, z.x - z.y);  // (DC, Nyquist) both ways
            continue;
        }
        const int rp = (nblk - r) & (nblk - 1);
        const int b0 = s ? (int)(__brev((unsigned)r) >> (32 - s)) : 0, b1 = s ? (int)(__brev((unsigned)rp) >> (32 - s)) : 0;
        const int qp = r ? 255 - q : 256 - q;
        const int p0 = swz((b0 << 8) | q), p1 = swz((b1 << 8) | qp);
        const float2 z = sm[p0], zp = sm[p1];
        const float2 w = tw_2m(t, k), wp = make_float2(-w.x, w.y);  // W_2M^(M-k) = -conj(W_2M^k)
        if (!INVERSE) {
            sm[p0] = r2c_bin(z, zp, w);
            if (p1 != p0) sm[p1] = r2c_bin(zp, z, wp);
        } else {
            sm[p0] = c2r_bin(z, zp, w);
            if (p1 != p0) sm[p1] = c2r_bin(zp, z, wp);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void cta_split_r2c(float2 *sm, int M, int s, const CtaTw &t) { cta_split<false>(sm, M, s, t); }
__device__ __forceinline__ void cta_split_c2r(float2 *sm, int M, int s, const CtaTw &t) { cta_split<true>(sm, M, s, t); }

}  // namespace ca
