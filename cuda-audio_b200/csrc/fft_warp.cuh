// fft_warp.cuh -- one-warp complex FFT of M = 32*R points (R = 1..32) for sm_100a.
//
// Replaces the reference's cufftExecC2C calls (src/conv.cu:243,367,405,407) and its
// two-for-one split kernel f_unpackC22R (src/conv.cu:47-73) for the partitioned engine:
// a real FFT of size 2B (B = period) is an M = B point complex FFT of z[n] = w[2n] + j w[2n+1]
// followed (R2C) or preceded (C2R) by a split pass.
//
// Decomposition (four-step, M = 32 * R):
//   time layout     lane a, register b  <->  n = R*a + b      (R contiguous complex per lane:
//                                                              float4 global loads/stores)
//   spectral layout lane l, register d  <->  k = brev5(l) + 32*d
//   forward : 32-point DIF across lanes with __shfl_xor butterflies -> twiddle W_M^(b*c)
//             -> R-point radix-2/4/8 DFT in registers (compile-time twiddles)
//   inverse : the transpose: R-point DFT in registers -> conj twiddle -> 32-point DIT
//             across lanes.
// Index math validated against numpy in tests/test_fft_model.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ca {

// ---- compile-time sin/cos (Taylor, fp64) so every register-DFT twiddle is an immediate ----
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double cx_sin(double x)
{
    double term = x, sum = x;
    for (int i = 1; i < 24; i++) { term *= -x * x / double((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double cx_cos(double x)
{
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < 24; i++) { term *= -x * x / double((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

// multiply by W_N^K (forward, exp(-2 pi i K/N)) or its conjugate (INV), with the trivial
// cases folded at compile time (this is what makes the radix-2 recursion below cost the
// same as hand-written radix-4 / radix-8 butterflies).
template <int N, int K, bool INV>
__device__ __forceinline__ float2 mul_tw(float2 v)
{
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) return v;
    else if constexpr (2 * k == N) return make_float2(-v.x, -v.y);
    else if constexpr (4 * k == N) return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);       // -j / +j
    else if constexpr (4 * k == 3 * N) return INV ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
    else if constexpr (8 * k == N) {  // (1 - j)/sqrt2  (conj: (1 + j)/sqrt2)
        constexpr float h = 0.70710678118654752440f;
        return INV ? make_float2(h * (v.x - v.y), h * (v.x + v.y)) : make_float2(h * (v.x + v.y), h * (v.y - v.x));
    } else if constexpr (8 * k == 3 * N) {  // (-1 - j)/sqrt2 (conj: (-1 + j)/sqrt2)
        constexpr float h = 0.70710678118654752440f;
        return INV ? make_float2(h * (-v.x - v.y), h * (v.x - v.y)) : make_float2(h * (v.y - v.x), h * (-v.x - v.y));
    } else {
        constexpr float c = (float)cx_cos(2.0 * kPi * double(k) / double(N));
        constexpr float s = (float)cx_sin(2.0 * kPi * double(k) / double(N));  // W = c - j s (fwd)
        return INV ? make_float2(v.x * c - v.y * s, v.x * s + v.y * c) : make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
    }
}

// ---- R-point DFT in registers: recursive radix-2 DIT, natural order in and out ------------
template <int N, int S, bool INV>
struct RegDft {
    template <int K>
    static __device__ __forceinline__ void combine(const float2 (&e)[N / 2], const float2 (&o)[N / 2], float2 *out)
    {
        if constexpr (K < N / 2) {
            float2 t = mul_tw<N, K, INV>(o[K]);
            out[K] = cadd(e[K], t);
            out[K + N / 2] = csub(e[K], t);
            combine<K + 1>(e, o, out);
        }
    }
    static __device__ __forceinline__ void run(const float2 *in, float2 *out)
    {
        float2 e[N / 2], o[N / 2];
        RegDft<N / 2, 2 * S, INV>::run(in, e);
        RegDft<N / 2, 2 * S, INV>::run(in + S, o);
        combine<0>(e, o, out);
    }
};
template <int S, bool INV>
struct RegDft<1, S, INV> {
    static __device__ __forceinline__ void run(const float2 *in, float2 *out) { out[0] = in[0]; }
};

template <int R, bool INV>
__device__ __forceinline__ void dft_reg(float2 (&v)[R])
{
    if constexpr (R > 1) {
        float2 out[R];
        RegDft<R, 1, INV>::run(v, out);
#pragma unroll
        for (int i = 0; i < R; i++) v[i] = out[i];
    }
}

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int brev5(int x) { return (int)(__brev((unsigned)x) >> 27); }

template <int R>
struct WarpFft {
    static constexpr int M = 32 * R;
    int lane;    // 0..31
    int c;       // brev5(lane): this lane's spectral residue (k = c + 32 d)
    int plane;   // lane holding the conjugate partner residue (32 - c) & 31
    float2 lw[5];   // W_{2*half}^{lane & (half-1)} for half = 16 >> s
    float2 twb[R];  // W_M^(b*c)

    // twM: exp(-2 pi i n / (M * stride)), n < M * stride (global memory, fp64-accurate values rounded
    // to fp32); stride > 1 lets a larger transform's table serve its 32R-point sub-transforms.
    __device__ __forceinline__ void init(const float2 *__restrict__ twM, int stride = 1)
    {
        lane = threadIdx.x & 31;
        c = brev5(lane);
        plane = brev5((32 - c) & 31);
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const int half = 16 >> s;
            const int j = lane & (half - 1);
            lw[s] = __ldg(&twM[(j << s) * R * stride]);  // W_32^(j * 16/half) = W_M^(j * (1<<s) * R)
        }
#pragma unroll
        for (int b = 0; b < R; b++) twb[b] = __ldg(&twM[b * c * stride]);
    }

    // time layout -> spectral layout, unnormalised forward DFT
    __device__ __forceinline__ void forward(float2 (&v)[R]) const
    {
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const int half = 16 >> s;
            const bool up = (lane & half) != 0;
#pragma unroll
            for (int b = 0; b < R; b++) {
                float2 p;
                p.x = __shfl_xor_sync(kFull, v[b].x, half);
                p.y = __shfl_xor_sync(kFull, v[b].y, half);
                float2 sum = cadd(v[b], p), dif = csub(p, v[b]);
                if (s < 4) dif = cmul(dif, lw[s]);
                v[b] = up ? dif : sum;
            }
        }
#pragma unroll
        for (int b = 1; b < R; b++) v[b] = cmul(v[b], twb[b]);
        dft_reg<R, false>(v);
    }

    // spectral layout -> time layout, unnormalised inverse DFT
    __device__ __forceinline__ void inverse(float2 (&v)[R]) const
    {
        dft_reg<R, true>(v);
#pragma unroll
        for (int b = 1; b < R; b++) v[b] = cmulc(v[b], twb[b]);
#pragma unroll
        for (int s = 4; s >= 0; s--) {
            const int half = 16 >> s;
            const bool up = (lane & half) != 0;
#pragma unroll
            for (int b = 0; b < R; b++) {
                float2 vt = v[b];
                if (s < 4) { float2 t = cmulc(vt, lw[s]); vt = up ? t : vt; }
                float2 p;
                p.x = __shfl_xor_sync(kFull, vt.x, half);
                p.y = __shfl_xor_sync(kFull, vt.y, half);
                v[b] = up ? csub(p, vt) : cadd(vt, p);
            }
        }
    }

    // partner[d] = V[(M - k) mod M] for k = c + 32 d, from a spectral-layout array
    __device__ __forceinline__ void partner(const float2 (&v)[R], float2 (&p)[R]) const
    {
#pragma unroll
        for (int d = 0; d < R; d++) {
            float2 q;
            q.x = __shfl_sync(kFull, v[R - 1 - d].x, plane);
            q.y = __shfl_sync(kFull, v[R - 1 - d].y, plane);
            const float2 own = v[(R - d) % R];
            p[d] = (c == 0) ? own : q;
        }
    }

    // real FFT split.  In: Z = FFT_M(z) in spectral layout.  Out: X[k], k = c + 32 d, of the
    // 2M-point real FFT; bin 0 is packed as (X[0], X[M]) = (DC, Nyquist).
    // tw2M: exp(-2 pi i k / (2M)), k < M.
    __device__ __forceinline__ void split_r2c(float2 (&v)[R], const float2 *__restrict__ tw2M) const
    {
        float2 p[R];
        partner(v, p);
#pragma unroll
        for (int d = 0; d < R; d++) {
            const int k = c + 32 * d;
            const float2 w = __ldg(&tw2M[k]);
            const float2 z = v[d], zp = p[d];
            const float2 e = make_float2(0.5f * (z.x + zp.x), 0.5f * (z.y - zp.y));
            const float2 dd = make_float2(z.x - zp.x, z.y + zp.y);
            const float2 m = cmul(dd, w);
            float2 x = make_float2(e.x + 0.5f * m.y, e.y - 0.5f * m.x);
            if (d == 0 && c == 0) x = make_float2(z.x + z.y, z.x - z.y);
            v[d] = x;
        }
    }

    // inverse split.  In: packed spectrum Y[k] (k = c + 32 d) of a real 2M signal.
    // Out: Z with IFFT_M(Z)[n] = M * (y[2n] + j y[2n+1]) * 2  (the 1/(2M) lives in the IR spectra).
    __device__ __forceinline__ void split_c2r(float2 (&v)[R], const float2 *__restrict__ tw2M) const
    {
        float2 p[R];
        partner(v, p);
#pragma unroll
        for (int d = 0; d < R; d++) {
            const int k = c + 32 * d;
            const float2 w = __ldg(&tw2M[k]);
            const float2 y = v[d], yp = p[d];
            const float2 s = make_float2(y.x + yp.x, y.y - yp.y);
            const float2 dd = make_float2(y.x - yp.x, y.y + yp.y);
            const float2 m = cmulc(dd, w);
            float2 z = make_float2(s.x - m.y, s.y + m.x);
            if (d == 0 && c == 0) z = make_float2(y.x + y.y, y.x - y.y);
            v[d] = z;
        }
    }
};

}  // namespace ca
