// fft_rows.cuh -- 256-point complex row FFT (one warp per row, 8 points per lane) and the four-step
// decomposition M = M1 x 256 built on it, for sm_100a.
//
// Why a second FFT family beside fft_warp.cuh / fft_cta.cuh: ncu of round 1 showed the FFT kernels of a
// batched period at ~1.1 K warp instructions per 256-point transform and 125-128 registers per thread
// (5 shuffle stages x 8 registers, 16..20 warps per SM), and the 16 K-point tier transform holding 128 KB of
// shared memory per CTA -- too heavy to run BESIDE the memory-bound MAC.  Here:
//   * a row is 256 = 8 x 8 x 4: two radix-8 DFTs and one radix-4 pair in registers (compile-time twiddles,
//     fft_warp.cuh), three exchanges through a 2.3 KB shared-memory region private to the warp
//     (__syncwarp only, every access bank-conflict free), twiddles from a 4 KB shared table;
//   * a long-tier transform of M = M1 x 256 points is M1-point column DFTs (registers, or two register
//     stages for M1 = 32 / 64) followed by M1 independent row FFTs, so a 16 K-point transform spreads over
//     40 small CTAs instead of one 128 KB CTA, and everything fits 64 registers per thread;
//   * spectra keep the POSITION order of fft_cta.cuh (bin k = k1 + M1 k2 at (bitrev(k1) << 8) | k2), so IR
//     spectra, delay lines and partial sums are interchangeable between the two families.
// Index math modelled in numpy (tests/_rows_fft_model.py, tests/test_fft_model.py), including the
// bank-conflict count of every shared-memory access pattern.
#pragma once
#include "fft_cta.cuh"

namespace ca {

constexpr int kRowSlots = 288;  // float2 slots of one row region: 8 x (32 + 4) = 4 x (64 + 4) + 16 = 256 + 8 x 4

struct RowTables {
    float2 w256[256];  // W_256^n
    float2 w512[256];  // W_512^k   (real-FFT split of a row)
};

// g: [W_256^n, n < 256 | W_512^k, k < 256] in global memory; all threads; ends with __syncthreads()
__device__ __forceinline__ void rows_tables_init(RowTables &t, const float2 *__restrict__ g)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        t.w256[i] = __ldg(&g[i]);
        t.w512[i] = __ldg(&g[256 + i]);
    }
    __syncthreads();
}

// natural order with 4 slots of padding per 32 (conflict-free for unit-stride AND for the stage-3 writes)
__device__ __forceinline__ int e3(int k) { return k + ((k >> 5) << 2); }

template <bool INV>
__device__ __forceinline__ float2 tmul(float2 v, float2 w) { return INV ? cmulc(v, w) : cmul(v, w); }

struct NoPostTw {
    __device__ __forceinline__ float2 operator()(int, float2 v) const { return v; }
};

// In : v[b] = x[lane + 32 b]  (registers)
// Out: X[k] (unnormalised forward / inverse DFT) in natural order at row[e3(k)]; element n2 = lane + 32 j is
//      passed through post(j, value) on its way out (the four-step twiddle of the inverse direction).
// `row` is private to the calling warp; all 32 lanes must call.
template <bool INV, class Post>
__device__ __forceinline__ void row_fft256(float2 (&v)[8], float2 *row, const RowTables &tb, int lane, const Post &post)
{
    // stage 1: radix 8 over b, twiddle W_256^(lane q)
    dft_reg<8, INV>(v);
#pragma unroll
    for (int q = 0; q < 8; q++) row[q * 36 + lane] = q ? tmul<INV>(v[q], tb.w256[lane * q]) : v[0];
    __syncwarp();
    // stage 2: thread (q, l0) = (lane >> 2, lane & 3): radix 8 over l1, twiddle W_32^(l0 r0)
    const int q2 = lane >> 2, l0 = lane & 3;
#pragma unroll
    for (int l1 = 0; l1 < 8; l1++) v[l1] = row[q2 * 36 + l0 + 4 * l1];
    __syncwarp();
    dft_reg<8, INV>(v);
#pragma unroll
    for (int r0 = 0; r0 < 8; r0++) row[l0 * 68 + r0 * 8 + q2] = r0 ? tmul<INV>(v[r0], tb.w256[8 * l0 * r0]) : v[0];
    __syncwarp();
    // stage 3: groups g = lane and lane + 32 (g = 8 r0 + q): radix 4 over l0; X[g + 64 r1]
    float2 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = row[i * 68 + lane]; b[i] = row[i * 68 + lane + 32]; }
    __syncwarp();
    dft_reg<4, INV>(a);
    dft_reg<4, INV>(b);
#pragma unroll
    for (int r1 = 0; r1 < 4; r1++) {
        row[lane + 72 * r1] = post(2 * r1, a[r1]);           // e3(lane + 64 r1)
        row[lane + 36 + 72 * r1] = post(2 * r1 + 1, b[r1]);  // e3(lane + 32 + 64 r1)
    }
    __syncwarp();
}

// ---- real-FFT split in position order -------------------------------------------------------------
// Row r (k1 = r) of an M = M1 x 256 transform pairs with row rp = (M1 - r) % M1: bin (r, k2) with
// (rp, 255 - k2), or (0, 256 - k2) inside row 0.  The warp that owns row r handles k2 in [0, 128) of its own
// row plus their partners, so every pair is visited exactly once by exactly one lane (in place).
// cr = W_2M^r (1 for row 0); w512 = W_512^k2; W_2M^(k1 + M1 k2) = cr * W_512^k2.
template <bool INV>
__device__ __forceinline__ void rows_split(float2 *own, float2 *partner, bool row0, float2 cr, const RowTables &tb, int lane)
{
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const int k2 = lane + 32 * d;
        if (row0 && k2 == 0) {
            // position 0 = (DC, Nyquist) both ways; bin 128 of row 0 is its own partner
            const float2 z0 = own[0];
            own[0] = make_float2(z0.x + z0.y, z0.x - z0.y);
            const float2 zn = own[e3(128)], wn = tb.w512[128];
            own[e3(128)] = INV ? c2r_bin(zn, zn, wn) : r2c_bin(zn, zn, wn);
            continue;
        }
        const int p0 = e3(k2), p1 = e3(row0 ? 256 - k2 : 255 - k2);
        const float2 z = own[p0], zp = partner[p1];
        const float2 w = row0 ? tb.w512[k2] : cmul(cr, tb.w512[k2]);
        const float2 wp = make_float2(-w.x, w.y);  // W_2M^(M - k) = -conj(W_2M^k)
        if (!INV) {
            own[p0] = r2c_bin(z, zp, w);
            partner[p1] = r2c_bin(zp, z, wp);
        } else {
            own[p0] = c2r_bin(z, zp, w);
            partner[p1] = c2r_bin(zp, z, wp);
        }
    }
}

// row region (natural order, e3 layout) <-> 2 KB of global memory, float4 per lane (bins 2 l + 64 i, +1)
__device__ __forceinline__ void row_store_global(const float2 *row, float2 *dst, int lane)
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int k = 2 * lane + 64 * i;
        reinterpret_cast<float4 *>(dst)[lane + 32 * i] = *reinterpret_cast<const float4 *>(row + e3(k));
    }
}

}  // namespace ca
