// kernels_quirks.cuh -- CA_FLAG_REF_QUIRKS: reproduce, on top of the exact convolution, what the reference's
// two-for-one FFT split does to the DC and Nyquist bins, so that the engine matches the reference on
// UNCONSTRAINED impulse responses too (without the flag parity holds for DC/Nyquist-free IRs, the protocol of 8c).
//
// What the reference does (conv.cu:47-73, 207-253, 367-409, fftSize = N):
//   * f_unpackC22R treats bin 0 of Z = FFT(a + j b) as `L[0] = Z[0], R[0] = 0` and never writes bin N/2.  The input
//     block is packed as (in1 + j in2), a stereo IR as (hL + j hR): X1[0] = (sum in1, sum in2), X2[0] = 0,
//     H_L[0] = (sum hL, sum hR), H_R[0] = 0, every bin N/2 = 0.
//   * f_pointwiseMultiplyAndScale's real part at bin 0 is therefore sL1 (sum in1 AL1 - sum in2 AR1) for the left
//     output and 0 for the right one, where A = sum over the live IR (glide coefficient x IR sum) and
//     s = pan * level / N; the exact values are sL1 sum in1 AL1 + sL2 sum in2 AL2 and sR1 sum in1 AR1 + sR2 sum in2 AR2.
//     Its imaginary part (also wrong, conv.cu:119-120) only reaches the imaginary part of the time signal, which the
//     reference never plays.  Bin N/2 contributes nothing although it should contribute s a' A' (-1)^n, with
//     a' = sum (-1)^n in, A' = sum over the live IR of c sum (-1)^n h.
//   * One bin of an N-point inverse FFT is a constant (DC) or an alternating (Nyquist) sequence over the WHOLE
//     accumulator: block t adds D + E (-1)^(s - pd) to accumulator samples s in [pd, N) (f_pointwiseAdd,
//     conv.cu:89-100: delayed by predelay, cut at N).
// So: reference output = exact convolution + per-block rank-1 terms, each alive for N - pd samples.  The kernel
// below keeps their running sums per (instance, output): a block's terms join at the first period boundary after
// tB + pd and leave at tB + N (a period boundary); the part of the starting period goes through a small ring.
// It also finishes the block: clamp(wet + correction) + dry mix -- tier 0's inverse kernel runs in raw-wet mode.
// Checked in fp64 against the pinned CPU restatement of conv.cu (tests: test_ref_quirk_model_*) and on the GPU
// against the live reference (tests/test_engine_gpu.py::test_ref_quirks_*).
#pragma once
#include "kernels.cuh"

namespace ca {

struct QuirkArgs {
    const float *in;       // [inst][2][B]
    float *out;            // [inst][2][B]: raw wet block in, finished block out
    const InParamDev *par;
    const ItemState *st;   // [2][n_items_alloc]
    const Ctl *ctl;
    const double *irsum;   // [slot][4]: sum hL, sum hR, sum (-1)^n hL, sum (-1)^n hR
    double *delta;         // [inst][Kr][4]: (D_L, D_R, E_L, E_R) joining (+) / leaving (-) the running sums at period p
    double *run;           // [inst][4]: running sums (D_L, D_R, E_L, E_R), E already signed by (-1)^pd
    float *qring;          // [inst][2][qlen]: terms of blocks that start inside a period
    uint32_t n_items_alloc, nv, B, N, Kr, qlen, inst0, n_inst;
    unsigned long long t_host_p1;
};

constexpr int kQuirkWarps = 4;

__global__ void __launch_bounds__(kQuirkWarps * 32) k_ref_quirks(const QuirkArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t local = blockIdx.x * kQuirkWarps + warp;
    if (local >= a.n_inst) return;  // warp-uniform
    const uint32_t inst = a.inst0 + local;
    const unsigned long long t = a.t_host_p1 ? a.t_host_p1 - 1ull : a.ctl->t - 1ull;  // tier 0's inverse has advanced the counter
    const uint32_t B = a.B;
    const float *x0 = a.in + (size_t)inst * 2 * B, *x1 = x0 + B;
    // block sums: sum x, sum (-1)^n x  (n and the lane have the same parity)
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t n = lane; n < B; n += 32) {
        const float u = x0[n], v = x1[n];
        s[0] += u; s[1] += v;
    }
    s[2] = (lane & 1) ? -s[0] : s[0];
    s[3] = (lane & 1) ? -s[1] : s[1];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int off = 16; off; off >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);

    const InParamDev p0 = a.par[inst * 2], p1 = a.par[inst * 2 + 1];
    const uint32_t pd = p0.predelay;  // input 0's, conv.cu:412,415
    double term[4] = {0.0, 0.0, 0.0, 0.0};  // D_L, D_R, E_L, E_R of block t
    double *run = a.run + (size_t)inst * 4;
    double *delta = a.delta + (size_t)inst * a.Kr * 4;
    const uint32_t per = a.N / B;  // periods a block's terms live
    if (lane == 0) {
        // live IR sums per input: A[i][k] = sum_v c_v irsum[slot_v][k]  (Hlive = sum_v c_v H_v, kernels.cuh: ItemState)
        double A[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const ItemState &st = a.st[((t + 1ull) & 1ull) * a.n_items_alloc + inst * 2 + i];
#pragma unroll
            for (int k = 0; k < 4; k++) A[i][k] = 0.0;
            for (uint32_t v = 0; v < a.nv; v++)
                if ((st.active >> v) & 1u) {
                    const double c = (double)st.c[v];
                    const double *is = a.irsum + (size_t)st.slot[v] * 4;
#pragma unroll
                    for (int k = 0; k < 4; k++) A[i][k] += c * is[k];
                }
        }
        const double invN = 1.0 / (double)a.N;
        const double sL[2] = {(double)(pan_gain(p0.panWet, 0, 2) * p0.level) * invN, (double)(pan_gain(p1.panWet, 0, 2) * p1.level) * invN};
        const double sR[2] = {(double)(pan_gain(p0.panWet, 1, 2) * p0.level) * invN, (double)(pan_gain(p1.panWet, 1, 2) * p1.level) * invN};
        const double a1 = s[0], a2 = s[1], b1 = s[2], b2 = s[3];
        term[0] = -(sL[0] * a2 * A[0][1] + sL[1] * a2 * A[1][0]);
        term[1] = -(sR[0] * a1 * A[0][1] + sR[1] * a2 * A[1][1]);
        term[2] = -(sL[0] * b1 * A[0][2] + sL[1] * b2 * A[1][2]);
        term[3] = -(sR[0] * b1 * A[0][3] + sR[1] * b2 * A[1][3]);
        // terms whose first full period is this one join, terms of block t - N/B leave
        double *d = delta + (size_t)(t % a.Kr) * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) { run[k] += d[k]; d[k] = 0.0; }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) term[k] = __shfl_sync(0xffffffffu, term[k], 0);
    double r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = __shfl_sync(0xffffffffu, lane == 0 ? run[k] : 0.0, 0);

    const uint32_t qmask = a.qlen - 1;
    float *q0 = a.qring + (size_t)inst * 2 * a.qlen, *q1 = q0 + a.qlen;
    if (pd < a.N) {
        // block t's terms cover accumulator samples [tB + pd, tB + N): the rest of the period they start in goes
        // through the ring, from the next boundary on they are part of the running sums until period t + N/B
        const unsigned long long ps = t + pd / B;
        const uint32_t m0 = pd % B;
        const unsigned long long left = (t + per - ps) * (unsigned long long)B;  // samples from ps*B to the end of the block's life
        const uint32_t m_end = left < B ? (uint32_t)left : B;
        for (uint32_t m = m0 + lane; m < m_end; m += 32) {
            const float sg = ((m - m0) & 1u) ? -1.f : 1.f;
            const uint32_t idx = (uint32_t)((ps * B + m) & qmask);
            q0[idx] += (float)term[0] + sg * (float)term[2];
            q1[idx] += (float)term[1] + sg * (float)term[3];
        }
        if (lane == 0 && ps + 1ull < t + per) {
            const double sgn = (pd & 1u) ? -1.0 : 1.0;  // (-1)^(m - pd) = (-1)^m (-1)^pd in later periods (B is even)
            double *dj = delta + (size_t)((ps + 1ull) % a.Kr) * 4, *dl = delta + (size_t)((t + per) % a.Kr) * 4;
            dj[0] += term[0]; dj[1] += term[1]; dj[2] += sgn * term[2]; dj[3] += sgn * term[3];
            dl[0] -= term[0]; dl[1] -= term[1]; dl[2] -= sgn * term[2]; dl[3] -= sgn * term[3];
        }
    }
    __syncwarp();
    // finish the block: clamp(wet + correction) (conv.cu:98), dry mix (conv.cu:126-140, 418-427)
    const float dgL[2] = {p0.dry * pan_gain(p0.panDry, 0, 2) * p0.level, p1.dry * pan_gain(p1.panDry, 0, 2) * p1.level};
    const float dgR[2] = {p0.dry * pan_gain(p0.panDry, 1, 2) * p0.level, p1.dry * pan_gain(p1.panDry, 1, 2) * p1.level};
    float *oL = a.out + (size_t)inst * 2 * B, *oR = oL + B;
    for (uint32_t m = lane; m < B; m += 32) {
        const float sg = (m & 1u) ? -1.f : 1.f;
        const uint32_t idx = (uint32_t)((t * B + m) & qmask);
        const float cL = (float)r[0] + sg * (float)r[2] + q0[idx];
        const float cR = (float)r[1] + sg * (float)r[3] + q1[idx];
        q0[idx] = 0.f; q1[idx] = 0.f;  // consumed
        const float u = x0[m], v = x1[m];
        oL[m] = fminf(fmaxf(oL[m] + cL, -1.f), 1.f) + dgL[0] * u + dgL[1] * v;
        oR[m] = fminf(fmaxf(oR[m] + cR, -1.f), 1.f) + dgR[0] * u + dgR[1] * v;
    }
}

// sums of one stereo IR (time domain, as loaded): [sum hL, sum hR, sum (-1)^n hL, sum (-1)^n hR]
__global__ void __launch_bounds__(256) k_ir_sums(const float *hl, const float *hr, uint32_t frames, uint32_t stride, double *out)
{
    __shared__ double sm[4][256];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (uint32_t n = threadIdx.x; n < frames; n += 256) {
        const double l = hl[(size_t)n * stride], r = hr[(size_t)n * stride];
        const double sg = (n & 1u) ? -1.0 : 1.0;
        s[0] += l; s[1] += r; s[2] += sg * l; s[3] += sg * r;
    }
    for (int k = 0; k < 4; k++) sm[k][threadIdx.x] = s[k];
    __syncthreads();
    for (int w = 128; w; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int k = 0; k < 4; k++) sm[k][threadIdx.x] += sm[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 4) out[threadIdx.x] = sm[threadIdx.x][0];
}

}  // namespace ca
