// fft_cta.cuh -- CTA-level complex FFT of M = 256 * 2^s points (s = 0..6) in shared memory, for the
// long tiers of the non-uniform partitioning (block sizes 256 .. 16384).
//
//   forward : s radix-2 DIF stages in shared memory (span M -> 512), then one 256-point warp FFT
//             (fft_warp.cuh, R = 8) per contiguous 256-element block.
//   inverse : the transposed flow graph (warp inverse FFTs, then radix-2 DIT stages).
// After `forward` bin k lives at shared index zpos(k) = (bitrev_s(k mod 2^s) << 8) | (k >> s);
// `inverse` expects its input in that layout and leaves time samples in natural order.
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

// twM: W_M^n, n < M.  All threads of the CTA must call; ends with __syncthreads().
__device__ __forceinline__ void cta_fft_forward(float2 *sm, int M, int s, const float2 *__restrict__ twM)
{
    const int tid = threadIdx.x, nthreads = blockDim.x;
    for (int st = 0; st < s; st++) {
        const int lh = 31 - __clz(M) - 1 - st;  // log2(half)
        const int half = 1 << lh;
        for (int j = tid; j < M / 2; j += nthreads) {
            const int pos = j & (half - 1);
            const int i0 = ((j >> lh) << (lh + 1)) | pos, i1 = i0 + half;
            const float2 a = sm[i0], b = sm[i1];
            const float2 w = __ldg(&twM[pos << st]);  // W_{2 half}^pos
            sm[i0] = cadd(a, b);
            sm[i1] = cmul(csub(a, b), w);
        }
        __syncthreads();
    }
    WarpFft<8> f;
    f.init(twM, M >> 8);
    const int warp = tid >> 5, nwarps = nthreads >> 5;
    for (int blk = warp; blk < (1 << s); blk += nwarps) {
        float2 *base = sm + (blk << 8);
        float2 v[8];
#pragma unroll
        for (int b = 0; b < 8; b++) v[b] = base[8 * f.lane + b];
        f.forward(v);
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 8; d++) base[f.c + 32 * d] = v[d];
    }
    __syncthreads();
}

__device__ __forceinline__ void cta_fft_inverse(float2 *sm, int M, int s, const float2 *__restrict__ twM)
{
    const int tid = threadIdx.x, nthreads = blockDim.x;
    WarpFft<8> f;
    f.init(twM, M >> 8);
    const int warp = tid >> 5, nwarps = nthreads >> 5;
    for (int blk = warp; blk < (1 << s); blk += nwarps) {
        float2 *base = sm + (blk << 8);
        float2 v[8];
#pragma unroll
        for (int d = 0; d < 8; d++) v[d] = base[f.c + 32 * d];
        f.inverse(v);
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 8; b++) base[8 * f.lane + b] = v[b];
    }
    __syncthreads();
    for (int st = s - 1; st >= 0; st--) {
        const int lh = 31 - __clz(M) - 1 - st;
        const int half = 1 << lh;
        for (int j = tid; j < M / 2; j += nthreads) {
            const int pos = j & (half - 1);
            const int i0 = ((j >> lh) << (lh + 1)) | pos, i1 = i0 + half;
            const float2 a = sm[i0];
            const float2 b = cmulc(sm[i1], __ldg(&twM[pos << st]));
            sm[i0] = cadd(a, b);
            sm[i1] = csub(a, b);
        }
        __syncthreads();
    }
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

// in place over the zpos layout: Z (M-point FFT of z[n] = w[2n] + j w[2n+1]) -> packed real spectrum
__device__ __forceinline__ void cta_split_r2c(float2 *sm, int M, int s, const float2 *__restrict__ tw2M)
{
    for (int k = threadIdx.x; k <= M / 2; k += blockDim.x) {
        if (k == 0) {
            const float2 z = sm[0];
            sm[0] = make_float2(z.x + z.y, z.x - z.y);  // (DC, Nyquist)
        } else {
            const int p0 = zpos(k, s), p1 = zpos(M - k, s);
            const float2 z = sm[p0], zp = sm[p1];
            sm[p0] = r2c_bin(z, zp, __ldg(&tw2M[k]));
            if (p1 != p0) sm[p1] = r2c_bin(zp, z, __ldg(&tw2M[M - k]));
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void cta_split_c2r(float2 *sm, int M, int s, const float2 *__restrict__ tw2M)
{
    for (int k = threadIdx.x; k <= M / 2; k += blockDim.x) {
        if (k == 0) {
            const float2 y = sm[0];
            sm[0] = make_float2(y.x + y.y, y.x - y.y);
        } else {
            const int p0 = zpos(k, s), p1 = zpos(M - k, s);
            const float2 y = sm[p0], yp = sm[p1];
            sm[p0] = c2r_bin(y, yp, __ldg(&tw2M[k]));
            if (p1 != p0) sm[p1] = c2r_bin(yp, y, __ldg(&tw2M[M - k]));
        }
    }
    __syncthreads();
}

}  // namespace ca
