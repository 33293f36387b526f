/*
 * oracle/refconv.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64 arithmetic) of the reference's convolution
 * hot path, limitz/cuda-audio src/conv.cu.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (cuda-audio_b200/) never links or calls it.
 *
 * It follows the reference *literally*, quirks included, so that it can be
 * pinned against outputs of the real reference (oracle/_ref, run on a B200;
 * committed as tests/golden/ref_*.npz by tests/golden/make_golden.py):
 *
 *   - one fftSize-long C2C FFT per period, overlap-ADD into a residual
 *     accumulator                                  (conv.cu:287-466)
 *   - two-for-one real FFT split whose DC bin is mis-unpacked and whose
 *     Nyquist bin is never written                 (conv.cu:47-73)
 *   - per-period one-pole glide of the live IR spectrum (conv.cu:15-32)
 *   - 3-multiply complex product with the +2*a.y*b.y imaginary error
 *                                                   (conv.cu:102-123)
 *   - clamp(+-1) on the whole accumulator, predelay of input 0 for both
 *     outputs, dry added after the clamp            (conv.cu:89-100,126-140)
 *
 * Third-party arithmetic: the reference calls cuFFT (closed source) for the
 * unnormalised DFTs (conv.cu:190,243,367,405,407).  Its published contract is
 * a plain DFT, restated here as an iterative radix-2 FFT in fp64.  Allocations
 * are zero-filled, matching the zero-malloc shim the oracle/_ref build uses
 * (the reference itself reads uninitialised cudaMalloc memory in bin N/2).
 *
 * Parity status: pinned against oracle/_ref golden vectors (see
 * tests/test_oracle_golden.py).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

#define REF_MAX_PREDELAY 8192 /* CONV_MAX_PREDELAY, conv.h:26-28 */
#define REF_MAX_IR 256

/* ---- plain unnormalised DFT (stand-in for cufftExecC2C) ------------------ */
static void fft_inplace(cplx *a, size_t n, int inverse)
{
    /* bit reversal */
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        double ang = (inverse ? 2.0 : -2.0) * M_PI / (double)len;
        size_t half = len >> 1;
        /* twiddles recomputed per stage from exact angle (no drift) */
        cplx *w = (cplx *)malloc(sizeof(cplx) * half);
        for (size_t k = 0; k < half; k++) w[k] = cos(ang * (double)k) + I * sin(ang * (double)k);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < half; k++) {
                cplx u = a[i + k], v = a[i + k + half] * w[k];
                a[i + k] = u + v;
                a[i + k + half] = u - v;
            }
        free(w);
    }
}

typedef struct {
    /* Convolution::CC::value, conv.h:40-50 */
    size_t select, predelay, speed, vsteps;
    double dry, wet, panDry, panWet, level;
} refconv_cc;

typedef struct {
    size_t N;
    cplx *cin, *cin1, *cin2, *cinFFT;                    /* conv.h:68-69 */
    cplx *irL, *irR;                                     /* `ir` scratch, conv.h:73 */
    cplx *irFFT1L, *irFFT1R, *irFFT2L, *irFFT2R;         /* live IR spectra   */
    cplx *outL, *outR, *resL, *resR;                     /* N + 8192 each     */
    cplx *irbuf[REF_MAX_IR];                             /* _irBuffers, 2N    */
    refconv_cc cc[2];
} refconv;

static cplx *zalloc(size_t n) { return (cplx *)calloc(n, sizeof(cplx)); }

/* Convolution::Convolution, conv.cu:142-195 (buffers zero-filled, see header) */
refconv *refconv_create(size_t fftSize)
{
    refconv *r = (refconv *)calloc(1, sizeof(refconv));
    r->N = fftSize;
    r->cin = zalloc(fftSize); r->cin1 = zalloc(fftSize); r->cin2 = zalloc(fftSize);
    r->cinFFT = zalloc(fftSize);
    r->irL = zalloc(fftSize); r->irR = zalloc(fftSize);
    r->irFFT1L = zalloc(fftSize); r->irFFT1R = zalloc(fftSize);
    r->irFFT2L = zalloc(fftSize); r->irFFT2R = zalloc(fftSize);
    r->outL = zalloc(fftSize + REF_MAX_PREDELAY); r->outR = zalloc(fftSize + REF_MAX_PREDELAY);
    r->resL = zalloc(fftSize + REF_MAX_PREDELAY); r->resR = zalloc(fftSize + REF_MAX_PREDELAY);
    for (int i = 0; i < 2; i++) {
        /* defaults, conv.h:42-50 */
        r->cc[i].select = 0; r->cc[i].predelay = 0; r->cc[i].speed = 100; r->cc[i].vsteps = 0;
        r->cc[i].dry = 0.5; r->cc[i].wet = 0.5; r->cc[i].panDry = 0; r->cc[i].panWet = 0;
        r->cc[i].level = 1.0;
    }
    return r;
}

void refconv_destroy(refconv *r)
{
    if (!r) return;
    cplx *bufs[] = { r->cin, r->cin1, r->cin2, r->cinFFT, r->irL, r->irR, r->irFFT1L, r->irFFT1R,
                     r->irFFT2L, r->irFFT2R, r->outL, r->outR, r->resL, r->resR };
    for (size_t i = 0; i < sizeof(bufs) / sizeof(*bufs); i++) free(bufs[i]);
    for (int i = 0; i < REF_MAX_IR; i++) free(r->irbuf[i]);
    free(r);
}

void refconv_set_cc(refconv *r, int input, const refconv_cc *v) { r->cc[input] = *v; }
void refconv_get_cc(refconv *r, int input, refconv_cc *v) { *v = r->cc[input]; }

/* f_unpackC22R, conv.cu:47-73: two-real-from-one-complex split.
 * s == 0 uses vb = va (not conj(va)) => L[0] = Z[0], R[0] = 0; bin N/2 is
 * never written. */
static void unpackC22R(cplx *L, cplx *R, const cplx *src, size_t N)
{
    for (size_t s = 0; s < N / 2; s++) {
        size_t idxa = s, idxb = N - s;
        cplx va = src[idxa];
        cplx vb = s ? conj(src[idxb]) : va;
        cplx la = 0.5 * (va + vb);
        cplx lb = I * (-0.5 * (va - vb)); /* timesj, conv.cu:12 */
        L[idxa] = la;
        R[idxa] = lb;
        if (s) { L[idxb] = conj(la); R[idxb] = conj(lb); }
    }
}

/* Convolution::prepare, conv.cu:207-253.  left/right = the float2 wav buffer
 * (WavFile::buffer, wav.h:10) split into planes. */
int refconv_prepare(refconv *r, size_t idx, const float *left, const float *right,
                    size_t numFrames, size_t nframes)
{
    if (idx >= REF_MAX_IR) return -1;
    size_t N = r->N;
    free(r->irbuf[idx]);
    cplx *tmp = zalloc(N);
    cplx *buf = zalloc(2 * N);
    size_t n = numFrames < N - nframes ? numFrames : N - nframes; /* conv.cu:239 */
    for (size_t s = 0; s < n; s++) tmp[s] = (double)left[s] + I * (double)right[s];
    fft_inplace(tmp, N, 0);                                       /* conv.cu:243 */
    unpackC22R(buf, buf + N, tmp, N);                             /* conv.cu:246 */
    free(tmp);
    r->irbuf[idx] = buf;
    return 0;
}

/* f_interpolate, conv.cu:15-32 */
static void interpolate(cplx *dst, const cplx *a, const cplx *b, size_t N, size_t steps, double wet)
{
    double div = (double)(steps + 5);
    for (size_t s = 0; s < N / 2; s++) {
        cplx va = a[s];
        cplx vb = b[s] * wet;
        cplx vv = va + (vb - va) / div;
        dst[s] = vv;
        if (s) dst[N - s] = conj(vv);
    }
}

/* f_pointwiseMultiplyAndScale, conv.cu:102-123 (3-multiply product, imaginary
 * part carries the extra +2*a.y*h.y) */
static void mulscale(cplx *r, const cplx *ir1, const cplx *ir2, const cplx *a1, const cplx *a2,
                     size_t n, double scale1, double scale2)
{
    for (size_t s = 0; s < n; s++) {
        double a1x = creal(a1[s]), a1y = cimag(a1[s]), a2x = creal(a2[s]), a2y = cimag(a2[s]);
        double h1x = creal(ir1[s]), h1y = cimag(ir1[s]), h2x = creal(ir2[s]), h2y = cimag(ir2[s]);
        double re1 = a1x * h1x - a1y * h1y;
        double re2 = a2x * h2x - a2y * h2y;
        double im1 = (a1x + a1y) * (h1x + h1y) - re1;
        double im2 = (a2x + a2y) * (h2x + h2y) - re2;
        r[s] = (re1 * scale1 + re2 * scale2) + I * (im1 * scale1 + im2 * scale2);
    }
}

static double clampd(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }

/* f_pointwiseAdd, conv.cu:89-100 */
static void pointwise_add(cplx *r, const cplx *a, const cplx *b, size_t n, size_t predelay)
{
    for (size_t s = 0; s < n; s++) {
        cplx v = a[s];
        if (s >= predelay) v += b[s - predelay];
        r[s] = clampd(creal(v)) + I * clampd(cimag(v));
    }
}

/* Convolution::onProcess, conv.cu:287-466 */
int refconv_process(refconv *r, const float *IN1, const float *IN2, float *L, float *R,
                    size_t nframes)
{
    size_t N = r->N;
    refconv_cc *c0 = &r->cc[0], *c1 = &r->cc[1];
    if (!r->irbuf[c0->select] || !r->irbuf[c1->select]) return -2;

    /* conv.cu:321-328: memset + pack */
    memset(r->cin, 0, sizeof(cplx) * N);
    for (size_t s = 0; s < nframes; s++) r->cin[s] = (double)IN1[s] + I * (double)IN2[s];

    /* conv.cu:339-353: glide live IR spectra toward wet * selected IR */
    interpolate(r->irFFT1L, r->irFFT1L, r->irbuf[c0->select], N, c0->vsteps, c0->wet);
    interpolate(r->irFFT1R, r->irFFT1R, r->irbuf[c0->select] + N, N, c0->vsteps, c0->wet);
    if (c0->vsteps > 0) c0->vsteps--;
    interpolate(r->irFFT2L, r->irFFT2L, r->irbuf[c1->select], N, c1->vsteps, c1->wet);
    interpolate(r->irFFT2R, r->irFFT2R, r->irbuf[c1->select] + N, N, c1->vsteps, c1->wet);
    if (c1->vsteps > 0) c1->vsteps--;

    /* conv.cu:367-371: forward FFT + unpack into cin1 / cin2 */
    memcpy(r->cinFFT, r->cin, sizeof(cplx) * N);
    fft_inplace(r->cinFFT, N, 0);
    unpackC22R(r->cin1, r->cin2, r->cinFFT, N);

    /* conv.cu:386-401 */
    double panL1 = c0->panWet >= 0 ? 1 - c0->panWet : 1;
    double panR1 = c0->panWet <= 0 ? 1 + c0->panWet : 1;
    double panL2 = c1->panWet >= 0 ? 1 - c1->panWet : 1;
    double panR2 = c1->panWet <= 0 ? 1 + c1->panWet : 1;
    double invN = 1.0 / (double)N;
    mulscale(r->outL, r->irFFT1L, r->irFFT2L, r->cin1, r->cin2, N, invN * panL1 * c0->level,
             invN * panL2 * c1->level);
    mulscale(r->outR, r->irFFT1R, r->irFFT2R, r->cin1, r->cin2, N, invN * panR1 * c0->level,
             invN * panR2 * c1->level);

    /* conv.cu:403-408: inverse FFTs into the `ir` scratch */
    memcpy(r->irL, r->outL, sizeof(cplx) * N); fft_inplace(r->irL, N, 1);
    memcpy(r->irR, r->outR, sizeof(cplx) * N); fft_inplace(r->irR, N, 1);

    /* conv.cu:411-415: overlap-add, predelay of input 0 for both, clamp */
    pointwise_add(r->outL, r->resL, r->irL, N, c0->predelay);
    pointwise_add(r->outR, r->resR, r->irR, N, c0->predelay);

    /* conv.cu:418-427 + operators.h:338-342 (float2 += float hits x and y) */
    panL1 = c0->panDry >= 0 ? 1 - c0->panDry : 1;
    panR1 = c0->panDry <= 0 ? 1 + c0->panDry : 1;
    panL2 = c1->panDry >= 0 ? 1 - c1->panDry : 1;
    panR2 = c1->panDry <= 0 ? 1 + c1->panDry : 1;
    double sL1 = c0->dry * panL1 * c0->level, sR1 = c0->dry * panR1 * c0->level;
    double sL2 = c1->dry * panL2 * c1->level, sR2 = c1->dry * panR2 * c1->level;
    for (size_t s = 0; s < nframes; s++) {
        double x = creal(r->cin[s]), y = cimag(r->cin[s]);
        double dl = x * sL1 + y * sL2, dr = x * sR1 + y * sR2;
        r->outL[s] += dl + I * dl;
        r->outR[s] += dr + I * dr;
    }

    /* conv.cu:431-437: real parts out */
    for (size_t s = 0; s < nframes; s++) { L[s] = (float)creal(r->outL[s]); R[s] = (float)creal(r->outR[s]); }

    /* conv.cu:440-451: residual = output + nframes (entries >= N are never
     * written by pointwise_add, i.e. stay zero) */
    size_t cnt = N + REF_MAX_PREDELAY - nframes;
    memmove(r->resL, r->outL + nframes, cnt * sizeof(cplx));
    memmove(r->resR, r->outR + nframes, cnt * sizeof(cplx));
    return 0;
}

/* handleCC, conv.cu:255-276: MIDI CC byte -> parameter.  `which` selects the
 * parameter (the reference compares m2 against the per-parameter CC number):
 * 0 select, 1 predelay, 2 dry, 3 wet, 4 panDry, 5 panWet, 6 level, 7 speed. */
void refconv_handle_cc(refconv_cc *cc, int which, int v, size_t nb)
{
    switch (which) {
    case 0: cc->select = (size_t)v * nb / 0x80; cc->vsteps = cc->speed; break;
    case 1: cc->predelay = (size_t)v * REF_MAX_PREDELAY / 0x80; break;
    case 2: cc->dry = (double)(v / 128.0f); break;
    case 3: cc->wet = (double)(v / 128.0f); break;
    case 4: cc->panDry = (double)(v / 64.0f - 1); break;
    case 5: cc->panWet = (double)(v / 64.0f - 1); break;
    case 6: cc->level = (double)(v / 128.0f); break;
    case 7:
        cc->speed = ((size_t)v * 1024) / 0x80; /* CONV_MAX_SPEED, conv.h:22-24 */
        if (cc->vsteps > cc->speed) cc->vsteps = cc->speed;
        break;
    }
}

/* PCM decode restatement, wav.cu:4-44 (both formats land at HALF scale). */
void refconv_pcm16_to_float(const int16_t *in, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = in[i] / (float)65536; /* wav.cu:13-14 */
}
void refconv_pcm24_to_float(const uint8_t *in, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) { /* wav.cu:24-41 */
        uint32_t v = ((uint32_t)in[3 * i] << 8) | ((uint32_t)in[3 * i + 1] << 16) |
                     ((uint32_t)in[3 * i + 2] << 24);
        int32_t vv = (int32_t)v;
        vv /= 256;
        out[i] = vv / (float)16777216;
    }
}
